"""Prints %globaltimer phase deltas (ns) of single tcgen05 GEMM launches (dense fwd shapes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import _lib
from agents_b200.utils import workspace

dev = torch.device('cuda:0')
_lib.call('b200rl_set_gemm_mode', 1)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
_lib.call('b200rl_tc_debug_buffer', _lib.ptr(dbg))
ws, nb = workspace.get(dev, 64 << 20)
names = ['start', 'after_alloc_sync', 'producers_done', 'mma_issued', 'accum_ready', 'epilogue_done', 'dealloc']
for (M, K, N) in [(256, 6, 512), (256, 512, 64), (1024, 256, 32), (20736, 512, 64), (256, 3136, 512)]:
  x = torch.randn(M, K, device=dev); w = torch.randn(K, N, device=dev); y = torch.empty(M, N, device=dev)
  for rep in range(3):
    dbg.zero_()
    torch.cuda.synchronize()
    _lib.call('b200rl_dense_fwd', _lib.ptr(x), 0, _lib.ptr(w), None, _lib.ptr(y), M, K, N, 0, _lib.ptr(ws), nb, _lib.stream())
    torch.cuda.synchronize()
  d = dbg.view(64, 8).cpu()
  err = (y - x @ w).abs().max().item() / max(1e-9, (x @ w).abs().max().item())
  blk0 = d[0]
  print(f'M={M} K={K} N={N} relerr={err:.2e}')
  print('  blk0 deltas ns:', {names[i]: int(blk0[i] - blk0[0]) for i in range(1, 7)})
  t0 = d[:, 0][d[:, 0] > 0]
  t6 = d[:, 6][d[:, 6] > 0]
  if len(t0):
    print(f'  first-64 CTAs: start spread {int(t0.max()-t0.min())} ns, total {int(t6.max()-t0.min())} ns')
