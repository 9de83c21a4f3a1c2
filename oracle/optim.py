"""Optimiser / target-update formulas in numpy float32 (TEST INFRASTRUCTURE).

Adam / RMSProp follow TensorFlow's documented update rules (training_ops ApplyAdam /
ApplyRMSProp / ApplyCenteredRMSProp); the reference only calls optimizer.apply_gradients
(agents/dqn/dqn_agent.py:444) and pins no post-step values ("parity unpinned").
soft_variables_update: utils/common.py:250-346; Periodically: utils/common.py:450-507;
clip_gradient_norms: utils/eager_utils.py:227-246 (tf.clip_by_norm per variable).
"""
import numpy as np

f32 = np.float32


class AdamTF(object):
  def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
    self.lr, self.b1, self.b2, self.eps = f32(lr), f32(beta1), f32(beta2), f32(eps)
    self.t = 0
    self.m = None
    self.v = None

  def apply(self, params, grads):
    if self.m is None:
      self.m = [np.zeros_like(p) for p in params]
      self.v = [np.zeros_like(p) for p in params]
    self.t += 1
    t = f32(self.t)
    lr_t = f32(self.lr * np.sqrt(f32(1) - np.power(self.b2, t, dtype=f32)) /
               (f32(1) - np.power(self.b1, t, dtype=f32)))
    for p, g, m, v in zip(params, grads, self.m, self.v):
      m += (g - m) * (f32(1) - self.b1)
      v += (g * g - v) * (f32(1) - self.b2)
      p -= (m * lr_t) / (np.sqrt(v) + self.eps)


class RMSPropTF(object):
  def __init__(self, lr, decay=0.9, momentum=0.0, eps=1e-10, centered=False, ms_init=1.0):
    self.lr, self.decay, self.momentum, self.eps = f32(lr), f32(decay), f32(momentum), f32(eps)
    self.centered = centered
    self.ms_init = f32(ms_init)
    self.ms = self.mg = self.mom = None

  def apply(self, params, grads):
    if self.ms is None:
      self.ms = [np.full_like(p, self.ms_init) for p in params]
      self.mg = [np.zeros_like(p) for p in params]
      self.mom = [np.zeros_like(p) for p in params]
    for p, g, ms, mg, mom in zip(params, grads, self.ms, self.mg, self.mom):
      ms += (g * g - ms) * (f32(1) - self.decay)
      denom = ms + self.eps
      if self.centered:
        mg += (g - mg) * (f32(1) - self.decay)
        denom = ms - mg * mg + self.eps
      mom[...] = mom * self.momentum + self.lr * g / np.sqrt(denom)
      p -= mom


def soft_variables_update(source, target, tau=1.0):
  # utils/common.py:300-346
  if tau == 0.0:
    return
  for s, t in zip(source, target):
    if tau == 1.0:
      t[...] = s
    else:
      t[...] = (f32(1 - tau) * t + f32(tau) * s).astype(f32)


class Periodically(object):
  # utils/common.py:450-507: fires when ++counter % period == 0; period 1 -> always
  def __init__(self, body, period):
    self.body, self.period, self.counter = body, period, 0

  def __call__(self):
    if self.period is None:
      return False
    if self.period == 1:
      self.body()
      return True
    self.counter += 1
    if self.counter % self.period == 0:
      self.body()
      return True
    return False


def clip_by_norm(g, clip):
  # tf.clip_by_norm: g * clip / max(l2, clip)
  l2 = np.sqrt(np.sum(g.astype(f32) * g, dtype=f32))
  return (g * (f32(clip) / np.maximum(l2, f32(clip)))).astype(f32)


def clip_by_global_norm(grads, clip):
  norm = np.sqrt(sum(np.sum(g.astype(f32) * g, dtype=f32) for g in grads), dtype=f32)
  scale = f32(clip) / np.maximum(norm, f32(clip))
  return [(g * scale).astype(f32) for g in grads], norm
