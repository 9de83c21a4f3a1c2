"""numpy float32 layers with hand-written backward (TEST INFRASTRUCTURE).

Restates what the reference gets from Keras through TensorFlow's CPU kernels:
Dense / Conv2D(VALID, NHWC, HWIO filters) / Flatten as used by
networks/encoding_network.py:224-312, networks/q_network.py:126-135 and the Mnih'15 Atari
network examples/dqn/mnih15/dqn_train_eval_atari.py:104-110 (cast -> /255 -> conv stack).
tests/test_oracle_nn.py cross-checks every backward here against torch CPU autograd.
"""
import numpy as np

f32 = np.float32


def act_fwd(z, act):
  if act == 'relu':
    return np.maximum(z, f32(0))
  if act == 'tanh':
    return np.tanh(z).astype(f32)
  return z


def act_bwd(y, dy, act):
  if act == 'relu':
    return np.where(y > 0, dy, f32(0)).astype(f32)
  if act == 'tanh':
    return (dy * (f32(1) - y * y)).astype(f32)
  return dy


def dense_fwd(x, w, b, act=None):
  z = x.astype(f32) @ w
  if b is not None:
    z = z + b
  return act_fwd(z.astype(f32), act)


def dense_bwd(x, w, dz):
  """dz is the gradient w.r.t. the pre-activation. Returns (dx, dw, db)."""
  return (dz @ w.T).astype(f32), (x.T @ dz).astype(f32), dz.sum(axis=0).astype(f32)


def im2col(x, kh, kw, stride):
  """x [N,H,W,C] -> cols [N*OH*OW, KH*KW*C] with (ky,kx,c) order, c fastest."""
  n, h, w, c = x.shape
  oh = (h - kh) // stride + 1
  ow = (w - kw) // stride + 1
  s0, s1, s2, s3 = x.strides
  view = np.lib.stride_tricks.as_strided(
      x, shape=(n, oh, ow, kh, kw, c),
      strides=(s0, s1 * stride, s2 * stride, s1, s2, s3), writeable=False)
  return view.reshape(n * oh * ow, kh * kw * c), oh, ow


def conv2d_fwd(x, wt, b, stride, act=None):
  """VALID conv, x [N,H,W,C] f32, wt [KH,KW,C,F]."""
  kh, kw, c, f = wt.shape
  cols, oh, ow = im2col(np.ascontiguousarray(x, dtype=f32), kh, kw, stride)
  z = cols @ wt.reshape(kh * kw * c, f)
  if b is not None:
    z = z + b
  return act_fwd(z.astype(f32), act).reshape(x.shape[0], oh, ow, f)


def conv2d_bwd(x, wt, dz, stride, need_dx=True):
  """dz [N,OH,OW,F] gradient w.r.t. pre-activation. Returns (dx or None, dw, db)."""
  kh, kw, c, f = wt.shape
  n, h, w, _ = x.shape
  cols, oh, ow = im2col(np.ascontiguousarray(x, dtype=f32), kh, kw, stride)
  dz2 = dz.reshape(-1, f)
  dw = (cols.T @ dz2).astype(f32).reshape(kh, kw, c, f)
  db = dz2.sum(axis=0).astype(f32)
  dx = None
  if need_dx:
    dcol = (dz2 @ wt.reshape(kh * kw * c, f).T).astype(f32).reshape(n, oh, ow, kh, kw, c)
    dx = np.zeros((n, h, w, c), dtype=f32)
    for ky in range(kh):
      for kx in range(kw):
        dx[:, ky:ky + oh * stride:stride, kx:kx + ow * stride:stride, :] += dcol[:, :, :, ky, kx, :]
  return dx, dw, db


class Sequential(object):
  """A layer stack described by dicts, mirroring agents_b200.networks.sequential.

  layer kinds: {'kind': 'cast_scale', 'divisor': 255.0}, {'kind': 'conv', 'w','b','stride','act'},
  {'kind': 'flatten'}, {'kind': 'dense', 'w','b','act'}.
  """

  def __init__(self, layers):
    self.layers = layers

  def params(self):
    out = []
    for l in self.layers:
      if l['kind'] in ('conv', 'dense'):
        out.append(l['w'])
        if l.get('b') is not None:
          out.append(l['b'])
    return out

  def forward(self, x, keep=False):
    tape = []
    for l in self.layers:
      k = l['kind']
      xin = x
      if k == 'cast_scale':
        x = (x.astype(f32) / f32(l['divisor'])).astype(f32)
      elif k == 'conv':
        x = conv2d_fwd(x, l['w'], l.get('b'), l['stride'], l.get('act'))
      elif k == 'flatten':
        x = x.reshape(x.shape[0], -1)
      elif k == 'dense':
        x = dense_fwd(x, l['w'], l.get('b'), l.get('act'))
      else:
        raise ValueError(k)
      tape.append((xin, x))
    return (x, tape) if keep else x

  def backward(self, tape, dy):
    """Returns gradients aligned with params()."""
    grads = []
    first_param_layer = next(i for i, l in enumerate(self.layers) if l['kind'] in ('conv', 'dense'))
    for i in range(len(self.layers) - 1, -1, -1):
      l = self.layers[i]
      xin, y = tape[i]
      k = l['kind']
      if k == 'dense':
        dz = act_bwd(y, dy, l.get('act'))
        dx, dw, db = dense_bwd(xin, l['w'], dz)
        g = [dw] + ([db] if l.get('b') is not None else [])
        grads = g + grads
        dy = dx
      elif k == 'conv':
        dz = act_bwd(y, dy, l.get('act'))
        dx, dw, db = conv2d_bwd(xin, l['w'], dz, l['stride'], need_dx=i > first_param_layer)
        g = [dw] + ([db] if l.get('b') is not None else [])
        grads = g + grads
        dy = dx
      elif k == 'flatten':
        dy = dy.reshape(xin.shape)
      elif k == 'cast_scale':
        dy = None
    return grads
