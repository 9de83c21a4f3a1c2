"""Ablation of the tcgen05 GEMM pipeline on the Mnih'15 layers: which stage bounds a K block?

variant 0 = production, 20 = producers skip the shared-memory stores, 21 = MMA thread issues only
the first MMA, 22 = producers skip the global loads, 23 = skip both loads and stores, 24 = 21 + 23 (handshake only),
30 = mbarrier.try_wait instead of test_wait polling, 31 = one lane per producer warp polls, 32 = 30 + 31.
Results of variants != 0 are garbage by construction; only the timings matter.
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import _lib
from agents_b200.networks import layers as L
from agents_b200.networks import q_network
from agents_b200.specs import tensor_spec

dev = torch.device('cuda:0')
obs_spec = tensor_spec.TensorSpec((84, 84, 4), torch.uint8)
act_spec = tensor_spec.BoundedTensorSpec((), torch.int32, 0, 5)
net = q_network.QNetwork(obs_spec, act_spec, preprocessing_layers=L.CastScale(255.),
                         conv_layer_params=((32, 8, 4), (64, 4, 2), (64, 3, 1)), fc_layer_params=(512,), device=dev).set_seed(0)
net.create_variables()
x = torch.randint(0, 256, (256, 84, 84, 4), dtype=torch.uint8, device=dev)
layers = [l for l in net.layers if not isinstance(l, L.CastScale)]
acts = [x]
for l in layers:
  acts.append(l.forward(acts[-1]))
torch.cuda.synchronize()
for var in (0,):
  _lib.call('b200rl_tc_debug_variant', var)
  out = []
  for l, h in zip(layers, acts):
    if type(l).__name__ == 'Flatten':
      continue
    g = torch.cuda.CUDAGraph()
    l.forward(h); torch.cuda.synchronize()
    with torch.cuda.graph(g):
      for _ in range(10):
        l.forward(h)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    out.append(f'{type(l).__name__}{tuple(acts[layers.index(l) + 1].shape[1:])}: {e0.elapsed_time(e1) * 100:.1f} us')
  print("variant", var, " | ".join(out), flush=True)
_lib.call('b200rl_tc_debug_variant', 0)
