"""Decodes which (k, n) element the tensor core fetches for an MN-major B tile: with A = I
(K-major, validated) the GEMM output D[m, n] is the value the MMA read for B[k=m, n]."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import _lib
from agents_b200.utils import workspace

dev = torch.device('cuda:0')
_lib.call('b200rl_set_gemm_mode', 1)
ws, nb = workspace.get(dev, 64 << 20)
M, K, N = 128, 32, 128
x = torch.zeros(M, K, device=dev)
x[:K, :K] = torch.eye(K, device=dev)
kk, nn = torch.meshgrid(torch.arange(K, device=dev), torch.arange(N, device=dev), indexing='ij')
w = (kk * 1000 + nn).float()
for var in [0, 2, 4, 6, 1, 3, 5, 8, 9, 12, 13, 17, 25]:
  _lib.call('b200rl_tc_debug_variant', var)
  y = torch.full((M, N), -1.0, device=dev)
  _lib.call('b200rl_dense_fwd', _lib.ptr(x), 0, _lib.ptr(w), None, _lib.ptr(y), M, K, N, 0, _lib.ptr(ws), nb, _lib.stream())
  torch.cuda.synchronize()
  d = y[:K].cpu().numpy()
  exp = w.cpu().numpy()
  ok = np.isclose(d, exp)
  vals = np.rint(d).astype(np.int64)
  nz = int((vals != 0).sum())
  print(f'variant {var:2d}: correct {ok.mean():.4f} nonzero {nz}  row k=1 n=0..9: {[int(v) for v in vals[1, :10]]}  k=5 n=32..37: {[int(v) for v in vals[5, 32:38]]}', flush=True)
_lib.call('b200rl_tc_debug_variant', 0)
