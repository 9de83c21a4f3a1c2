"""%globaltimer phase stamps of the first 64 CTAs for the Mnih'15 conv layers (tcgen05 path)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import _lib
from agents_b200.networks import layers as L
from agents_b200.networks import q_network
from agents_b200.specs import tensor_spec

dev = torch.device('cuda:0')
dbg = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
obs_spec = tensor_spec.TensorSpec((84, 84, 4), torch.uint8)
act_spec = tensor_spec.BoundedTensorSpec((), torch.int32, 0, 5)
net = q_network.QNetwork(obs_spec, act_spec, preprocessing_layers=L.CastScale(255.),
                         conv_layer_params=((32, 8, 4), (64, 4, 2), (64, 3, 1)), fc_layer_params=(512,), device=dev).set_seed(0)
net.create_variables()
x = torch.randint(0, 256, (256, 84, 84, 4), dtype=torch.uint8, device=dev)
names = ['start', 'alloc_sync', 'producers_done', 'mma_issued', 'accum_ready', 'epilogue_done', 'dealloc']
layers = [l for l in net.layers if not isinstance(l, L.CastScale)]
for _ in range(2):
  net(x)
torch.cuda.synchronize()
_lib.call('b200rl_tc_debug_buffer', _lib.ptr(dbg))
h = x
for l in layers:
  dbg.zero_(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); y = l.forward(h); e1.record(); torch.cuda.synchronize()
  d = dbg.view(64, 8).cpu()
  if int(d[0, 0]) > 0:
    import numpy as np
    dd = (d[:, 1:7] - d[:, :1]).numpy()
    valid = d[:, 6].numpy() > 0
    med = np.median(dd[valid], axis=0)
    print(type(l).__name__, tuple(y.shape), f'{e0.elapsed_time(e1)*1e3:.1f} us  median phase ns:',
          {names[i + 1]: int(med[i]) for i in range(6)}, 'cta0 start spread', int(d[valid, 0].max() - d[valid, 0].min()))
  else:
    print(type(l).__name__, tuple(y.shape), f'{e0.elapsed_time(e1)*1e3:.1f} us (no tc stamps)')
  h = y
