"""Cross-checks the oracle's hand-written layer backward against torch CPU autograd, and the
optimiser formulas against torch restatements of TF's documented update rules."""
import numpy as np
import torch

from oracle import nn as onn
from oracle import optim as ooptim

f32 = np.float32


def _torch_forward(layers, x):
  for l in layers:
    k = l['kind']
    if k == 'cast_scale':
      x = x.float() / l['divisor']
    elif k == 'conv':
      w = l['w_t'].permute(3, 2, 0, 1)  # HWIO -> OIHW
      x = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, l['b_t'], stride=l['stride'])
      x = x.permute(0, 2, 3, 1)
    elif k == 'flatten':
      x = x.reshape(x.shape[0], -1)
    elif k == 'dense':
      x = x @ l['w_t'] + l['b_t']
    if l.get('act') == 'relu':
      x = torch.relu(x)
    elif l.get('act') == 'tanh':
      x = torch.tanh(x)
  return x


def test_conv_dense_stack_backward_matches_autograd():
  rng = np.random.RandomState(0)
  layers = [
      dict(kind='cast_scale', divisor=255.0),
      dict(kind='conv', w=(rng.randn(4, 4, 3, 8) * .2).astype(f32), b=rng.randn(8).astype(f32) * .1,
           stride=2, act='relu'),
      dict(kind='conv', w=(rng.randn(3, 3, 8, 6) * .2).astype(f32), b=rng.randn(6).astype(f32) * .1,
           stride=1, act='relu'),
      dict(kind='flatten'),
      dict(kind='dense', w=(rng.randn(6 * 3 * 3, 16) * .2).astype(f32), b=rng.randn(16).astype(f32) * .1,
           act='tanh'),
      dict(kind='dense', w=(rng.randn(16, 5) * .2).astype(f32), b=rng.randn(5).astype(f32) * .1, act=None),
  ]
  x = rng.randint(0, 256, size=(4, 12, 12, 3)).astype(np.uint8)
  net = onn.Sequential(layers)
  y, tape = net.forward(x, keep=True)
  dy = rng.randn(*y.shape).astype(f32)
  grads = net.backward(tape, dy)
  for l in layers:
    if 'w' in l:
      l['w_t'] = torch.tensor(l['w'], requires_grad=True)
      l['b_t'] = torch.tensor(l['b'], requires_grad=True)
  yt = _torch_forward(layers, torch.tensor(x))
  np.testing.assert_allclose(y, yt.detach().numpy(), rtol=2e-5, atol=2e-6)
  (yt * torch.tensor(dy)).sum().backward()
  want = []
  for l in layers:
    if 'w' in l:
      want += [l['w_t'].grad.numpy(), l['b_t'].grad.numpy()]
  assert len(grads) == len(want)
  for g, w in zip(grads, want):
    np.testing.assert_allclose(g, w, rtol=2e-4, atol=2e-5)


def test_adam_matches_closed_form_first_step():
  p = [np.array([1.0, -2.0], f32)]
  g = [np.array([0.5, 0.25], f32)]
  opt = ooptim.AdamTF(lr=1e-3, eps=1e-8)
  opt.apply(p, g)
  # first step: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1 -> step ~= lr * sign(g)
  np.testing.assert_allclose(p[0], [1.0 - 1e-3, -2.0 - 1e-3], rtol=1e-5)
  assert opt.t == 1


def test_rmsprop_first_step():
  p = [np.array([1.0], f32)]
  g = [np.array([2.0], f32)]
  opt = ooptim.RMSPropTF(lr=0.1, decay=0.9, eps=1e-10, ms_init=1.0)
  opt.apply(p, g)
  ms = 1.0 + (4.0 - 1.0) * 0.1
  np.testing.assert_allclose(p[0], [1.0 - 0.1 * 2.0 / np.sqrt(ms + 1e-10)], rtol=1e-6)


def test_clip_by_norm():
  g = np.array([3.0, 4.0], f32)
  np.testing.assert_allclose(ooptim.clip_by_norm(g, 1.0), [0.6, 0.8], rtol=1e-6)
  np.testing.assert_allclose(ooptim.clip_by_norm(g, 10.0), g, rtol=1e-6)


def test_torch_cpu_dqn_oracle_matches_numpy_oracle():
  """The torch-CPU restatement used as bench.py's reference arm computes the same train step as
  the line-by-line numpy oracle (losses within 1e-5 relative over 4 steps)."""
  from oracle import dqn as odqn
  from oracle import dqn_torch
  rng = np.random.RandomState(0)
  A, B = 4, 16
  def layers():
    r = np.random.RandomState(1)
    return [dict(kind='cast_scale', divisor=255.0),
            dict(kind='conv', w=(r.randn(4, 4, 2, 8) * .2).astype(f32), b=np.zeros(8, f32), stride=2, act='relu'),
            dict(kind='flatten'),
            dict(kind='dense', w=(r.randn(8 * 9 * 9, 16) * .1).astype(f32), b=np.zeros(16, f32), act='relu'),
            dict(kind='dense', w=(r.randn(16, A) * .1).astype(f32), b=np.full(A, -.2, f32), act=None)]
  a = odqn.DqnOracle(onn.Sequential(layers()),
                     ooptim.RMSPropTF(2.5e-4, decay=0.95, eps=1e-5, centered=True),
                     gamma=0.99, target_update_period=2)
  b = dqn_torch.DqnTorchOracle(layers(), target_update_period=2)
  for _ in range(4):
    e = dict(observation=rng.randint(0, 256, size=(B, 2, 20, 20, 2)).astype(np.uint8),
             step_type=rng.randint(0, 3, size=(B, 2)).astype(np.int32),
             action=rng.randint(0, A, size=(B, 2)).astype(np.int32),
             reward=rng.rand(B, 2).astype(f32), discount=(rng.rand(B, 2) > .1).astype(f32))
    la, lb = float(a.train(e)['loss']), b.train(e)
    np.testing.assert_allclose(lb, la, rtol=1e-5)
