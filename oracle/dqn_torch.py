"""torch-CPU restatement of the DQN train step (TEST INFRASTRUCTURE / CPU baseline).

Same algorithm as oracle/dqn.py (which follows agents/dqn/dqn_agent.py:412-645 line by line),
but with torch CPU tensors, autograd and oneDNN/MKL multi-threaded conv/matmul: the closest
stand-in available here for the reference's own `tf.function` CPU path (TensorFlow is not
installable in this image).  Used as bench.py's `cpu_baseline` /
`--impl reference` arm and cross-checked against oracle/dqn.py in tests/test_oracle_nn.py.
"""
import numpy as np
import torch


class TorchSequential(object):
  """Mirror of oracle.nn.Sequential with torch parameters (requires_grad)."""

  def __init__(self, layers):
    self.layers = []
    for l in layers:
      l = dict(l)
      if l['kind'] in ('conv', 'dense'):
        l['w'] = torch.tensor(np.asarray(l['w'], dtype=np.float32), requires_grad=True)
        l['b'] = torch.tensor(np.asarray(l['b'], dtype=np.float32), requires_grad=True)
      self.layers.append(l)

  def params(self):
    out = []
    for l in self.layers:
      if l['kind'] in ('conv', 'dense'):
        out += [l['w'], l['b']]
    return out

  def forward(self, x):
    for l in self.layers:
      k = l['kind']
      if k == 'cast_scale':
        x = x.float() / l['divisor']
      elif k == 'conv':
        w = l['w'].permute(3, 2, 0, 1)                      # HWIO -> OIHW
        x = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, l['b'], stride=l['stride'])
        x = x.permute(0, 2, 3, 1)
      elif k == 'flatten':
        x = x.reshape(x.shape[0], -1)
      elif k == 'dense':
        x = x @ l['w'] + l['b']
      if l.get('act') == 'relu':
        x = torch.relu(x)
      elif l.get('act') == 'tanh':
        x = torch.tanh(x)
    return x


def _huber(target, pred):
  e = (pred - target).abs()
  quad = torch.clamp(e, max=1.0)
  return 0.5 * quad * quad + (e - quad)


class DqnTorchOracle(object):

  def __init__(self, layers, lr=2.5e-4, decay=0.95, momentum=0.0, eps=1e-5, centered=True,
               gamma=0.99, reward_scale=1.0, loss_fn='huber', target_update_tau=1.0,
               target_update_period=2500, optimizer='rmsprop', ddqn=False):
    self.q_net = TorchSequential(layers)
    self.target_net = TorchSequential(layers)
    self.gamma, self.reward_scale, self.loss_fn = gamma, reward_scale, loss_fn
    self.tau, self.period, self.counter = target_update_tau, target_update_period, 0
    self.ddqn = ddqn
    self.opt = dict(kind=optimizer, lr=lr, decay=decay, momentum=momentum, eps=eps,
                    centered=centered, t=0)
    ps = self.q_net.params()
    self.slots = [dict(ms=torch.ones_like(p), mg=torch.zeros_like(p), mom=torch.zeros_like(p),
                       m=torch.zeros_like(p), v=torch.zeros_like(p)) for p in ps]
    self.train_step_counter = 0

  def _apply(self, grads):
    o = self.opt
    with torch.no_grad():
      if o['kind'] == 'adam':
        o['t'] += 1
        lr_t = o['lr'] * np.sqrt(1 - 0.999 ** o['t']) / (1 - 0.9 ** o['t'])
      for p, g, s in zip(self.q_net.params(), grads, self.slots):
        if o['kind'] == 'adam':
          s['m'] += (g - s['m']) * (1 - 0.9)
          s['v'] += (g * g - s['v']) * (1 - 0.999)
          p -= (s['m'] * lr_t) / (s['v'].sqrt() + o['eps'])
        else:
          s['ms'] += (g * g - s['ms']) * (1 - o['decay'])
          denom = s['ms'] + o['eps']
          if o['centered']:
            s['mg'] += (g - s['mg']) * (1 - o['decay'])
            denom = s['ms'] - s['mg'] * s['mg'] + o['eps']
          s['mom'].mul_(o['momentum']).add_(o['lr'] * g / denom.sqrt())
          p -= s['mom']

  def train(self, exp):
    """exp: dict of numpy/torch [B,T,...] arrays (step_type, observation, action, reward,
    discount). Returns the scalar loss (python float)."""
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    obs, st = t(exp['observation']), t(exp['step_type'])
    act, rew, disc = t(exp['action']), t(exp['reward']).float(), t(exp['discount']).float()
    B, T = rew.shape
    n = T - 1
    R = torch.zeros(B)
    for k in range(n - 1, -1, -1):
      R = R * (self.gamma * disc[:, k]) + rew[:, k]
    D = (self.gamma ** (n - 1)) * torch.prod(disc[:, :n], dim=1)
    q = self.q_net.forward(obs[:, 0])
    with torch.no_grad():
      nt = self.target_net.forward(obs[:, -1])
      sel = self.q_net.forward(obs[:, -1]) if self.ddqn else nt
      nq = nt.gather(1, sel.argmax(dim=1, keepdim=True)).squeeze(1)
      target = self.reward_scale * R + self.gamma * D * nq
    qsa = q.gather(1, act[:, 0].long().unsqueeze(1)).squeeze(1)
    l = _huber(target, qsa) if self.loss_fn == 'huber' else (target - qsa) ** 2
    valid = (st[:, 0] != 2).float()
    loss = (valid * l).sum() / B
    grads = torch.autograd.grad(loss, self.q_net.params())
    self._apply(grads)
    self.train_step_counter += 1
    self.counter += 1
    if self.period == 1 or self.counter % self.period == 0:
      with torch.no_grad():
        for s, tg in zip(self.q_net.params(), self.target_net.params()):
          if self.tau == 1.0:
            tg.copy_(s)
          else:
            tg.mul_(1 - self.tau).add_(self.tau * s)
    return float(loss.detach())
