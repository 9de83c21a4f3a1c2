"""Pipeline timeline of the tc2 GEMM: CTA 0 stamps %globaltimer at the start/end of every step of
each role (b200rl_tc2_trace_buffer).  For the Mnih'15 layers at batch 256 prints, per role, the
mean busy time per step, the mean period between steps and the hop latencies between roles:

  land   loader step end -> converter step start   (cp.async data landing + barrier hop)
  c2m    converter end   -> MMA start
  m2l    MMA commit      -> loader reuses the stage (S steps later)

    python profiles/tc2_trace.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import _lib  # noqa: E402
from agents_b200.networks import layers as L  # noqa: E402

dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(0)
STEPS = 256
buf = torch.zeros(4 * STEPS * 2, dtype=torch.int64, device=dev)


def bind(layer, in_shape):
  layer.build(in_shape)
  ps = [torch.randn(*s, device=dev, generator=gen) * 0.05 for s in layer.param_shapes()]
  gs = [torch.zeros_like(p) for p in ps]
  layer.bind(ps, gs)
  return layer


def trace(name, fn):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  buf.zero_()
  _lib.call('b200rl_tc2_trace_buffer', _lib.ptr(buf))
  fn()
  torch.cuda.synchronize()
  _lib.call('b200rl_tc2_trace_buffer', None)
  t = buf.view(4, STEPS, 2).cpu().numpy().astype(np.int64)
  out = dict(op=name)
  t0 = t[t > 0].min() if (t > 0).any() else 0
  for r, rn in enumerate(['load', 'conv', 'mma', 'epi']):
    valid = (t[r, :, 0] > 0) & (t[r, :, 1] > 0)
    n = int(valid.sum())
    if n == 0:
      continue
    s, e = t[r, :n, 0], t[r, :n, 1]
    out[rn] = dict(steps=n, busy_ns=round(float((e - s).mean()), 1),
                   period_ns=round(float(np.diff(s).mean()), 1) if n > 1 else None,
                   first_start_ns=int(s[0] - t0), last_end_ns=int(e[-1] - t0))
  n = min(int(((t[0, :, 1] > 0) & (t[1, :, 0] > 0) & (t[2, :, 0] > 0)).sum()), STEPS)
  if n > 1:
    out['land_ns'] = round(float((t[1, :n, 0] - t[0, :n, 1]).mean()), 1)
    out['c2m_ns'] = round(float((t[2, :n, 0] - t[1, :n, 1]).mean()), 1)
    out['conv_busy_ns'] = round(float((t[1, :n, 1] - t[1, :n, 0]).mean()), 1)
    out['mma_issue_ns'] = round(float((t[2, :n, 1] - t[2, :n, 0]).mean()), 1)
    # first 8 steps in detail (ns since the first stamp): load start/end, conv start/end, mma start/end
    out['head'] = [[int(t[r, i, w] - t0) for r in range(3) for w in range(2)] for i in range(min(n, 10))]
  print(json.dumps(out), flush=True)


B = 256
x = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=dev, generator=gen)
c1 = bind(L.Conv2D(32, 8, 4, activation='relu'), (84, 84, 4)); c1.pre_divisor = 255.0
h1 = c1.forward(x)
c2 = bind(L.Conv2D(64, 4, 2, activation='relu'), tuple(h1.shape[1:]))
h2 = c2.forward(h1)
c3 = bind(L.Conv2D(64, 3, 1, activation='relu'), tuple(h2.shape[1:]))
h3 = c3.forward(h2)
hf = h3.reshape(B, -1)
f1 = bind(L.Dense(512, activation='relu'), (hf.shape[1],))
h4 = f1.forward(hf)
for name, layer, xin, y in (('conv1', c1, x, h1), ('conv2', c2, h1, h2), ('conv3', c3, h2, h3), ('fc1', f1, hf, h4)):
  dz = torch.randn(y.shape, device=dev, generator=gen)
  trace(name + '.fwd', lambda: layer.forward(xin))
  if xin.dtype != torch.uint8:
    trace(name + '.dX', lambda: layer.backward_parts(xin, dz, True, False))
  trace(name + '.dW', lambda: layer.backward_parts(xin, dz, False, True, accumulate=1))
