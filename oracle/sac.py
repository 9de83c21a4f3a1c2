"""numpy restatement of SacAgent losses and train step (TEST INFRASTRUCTURE).

Follows agents/sac/sac_agent.py: _train :314-410 (critic step, then actor step with the UPDATED
critics, then alpha step with the UPDATED actor; three independent samples), target updater
:493-535, _actions_and_log_probs :537-557, critic_loss :559-643, actor_loss :645-694,
alpha_loss :696-739, target entropy default :281-296;
agents/sac/tanh_normal_projection_network.py:112-143 (loc, std = exp(log_std));
distributions/utils.py:40-160 + distributions/tanh_bijector_stable.py:68-82 (tanh-squashed
log-prob); agents/ddpg/critic_network.py:163-178 (concat(obs, action) MLP).
The three noise draws per step are supplied by the caller (eps arrays), because the reference's
sampling stream is unpinned.
"""
import copy

import numpy as np

from oracle import optim

f32 = np.float32
LOG2PI = f32(np.log(2 * np.pi))
LOG2 = f32(np.log(2.0))


def softplus(x):
  return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20)))).astype(f32)


def sample_and_log_prob(head, eps, amin, amax):
  """head [N, 2A] -> (action, logp, u)."""
  A = head.shape[1] // 2
  loc, ls = head[:, :A], head[:, A:]
  u = (loc + np.exp(ls) * eps).astype(f32)
  half, shift = (amax - amin) / 2, (amax + amin) / 2
  action = (shift + half * np.tanh(u)).astype(f32)
  logp = np.sum(-0.5 * eps * eps - ls - 0.5 * LOG2PI - np.log(half)
                - 2.0 * (LOG2 - u - softplus(-2.0 * u)), axis=1).astype(f32)
  return action, logp, u


def _agg(per_example, weights, global_batch=None):
  x = per_example if weights is None else np.where(weights == 0, f32(0), per_example * weights)
  gb = f32(global_batch if global_batch is not None else x.shape[0])
  return f32(np.sum(x, dtype=f32) / gb)


def critic_loss(q1, q2, tq1, tq2, next_logp, reward, discount, log_alpha, gamma=1.0,
                reward_scale=1.0, weights=None, global_batch=None):
  """:603-634 with td_errors_loss_fn = squared_difference. Returns (loss, td_targets)."""
  tq = np.minimum(tq1, tq2) - np.exp(f32(log_alpha)) * next_logp
  y = (f32(reward_scale) * reward + f32(gamma) * discount * tq).astype(f32)
  per = ((y - q1) ** 2 + (y - q2) ** 2).astype(f32)
  return _agg(per, weights, global_batch), y


def actor_loss(q1, q2, logp, log_alpha, weights=None, global_batch=None):
  """:676-690."""
  per = (np.exp(f32(log_alpha)) * logp - np.minimum(q1, q2)).astype(f32)
  return _agg(per, weights, global_batch)


def alpha_loss(logp, log_alpha, target_entropy, use_log_alpha=True, weights=None,
               global_batch=None):
  """:716-735."""
  diff = (-logp - f32(target_entropy)).astype(f32)
  coef = f32(log_alpha) if use_log_alpha else np.exp(f32(log_alpha))
  return _agg((coef * diff).astype(f32), weights, global_batch)


class SacOracle(object):
  """actor: oracle.nn.Sequential ending in a linear Dense(2A); critics: Sequentials on
  concat(obs, action) ending in Dense(1)."""

  def __init__(self, actor, critic1, critic2, amin, amax, actor_opt, critic_opt, alpha_opt,
               gamma=0.99, reward_scale=1.0, tau=0.005, period=1, target_entropy=None,
               initial_log_alpha=0.0, critic_loss_weight=0.5, actor_loss_weight=1.0,
               alpha_loss_weight=1.0):
    self.actor, self.c1, self.c2 = actor, critic1, critic2
    self.t1, self.t2 = copy.deepcopy(critic1), copy.deepcopy(critic2)
    self.amin, self.amax = np.asarray(amin, f32), np.asarray(amax, f32)
    A = self.amin.shape[0]
    self.target_entropy = f32(-A / 2.0 if target_entropy is None else target_entropy)   # :281-296
    self.log_alpha = np.array([initial_log_alpha], f32)
    self.actor_opt, self.critic_opt, self.alpha_opt = actor_opt, critic_opt, alpha_opt
    self.gamma, self.reward_scale = gamma, reward_scale
    self.cw, self.aw, self.alw = critic_loss_weight, actor_loss_weight, alpha_loss_weight
    self.train_step_counter = 0
    self._update_target = optim.Periodically(
        lambda: optim.soft_variables_update(self.c1.params() + self.c2.params(),
                                            self.t1.params() + self.t2.params(), tau), period)

  def _q(self, net, obs, act, keep=False):
    x = np.concatenate([obs, act], axis=1).astype(f32)
    if keep:
      q, tape = net.forward(x, keep=True)
      return q[:, 0], tape
    return net.forward(x)[:, 0]

  def train(self, exp, eps_next, eps_actor, eps_alpha, weights=None):
    """exp: [B, 2, ...] arrays (observation, action, reward, discount)."""
    obs, obs2 = exp['observation'][:, 0], exp['observation'][:, 1]
    act = exp['action'][:, 0]
    r, d = exp['reward'][:, 0], exp['discount'][:, 0]
    B = r.shape[0]
    la = float(self.log_alpha[0])
    # ---- critic step (:341-362)
    a2, logp2, _ = sample_and_log_prob(self.actor.forward(obs2), eps_next, self.amin, self.amax)
    tq1, tq2 = self._q(self.t1, obs2, a2), self._q(self.t2, obs2, a2)
    q1, tape1 = self._q(self.c1, obs, act, keep=True)
    q2, tape2 = self._q(self.c2, obs, act, keep=True)
    closs, y = critic_loss(q1, q2, tq1, tq2, logp2, r, d, la, self.gamma, self.reward_scale, weights)
    closs = f32(self.cw * closs)
    w = np.ones(B, f32) if weights is None else np.asarray(weights, f32)
    g1 = (-2 * (y - q1) * self.cw * w / B).astype(f32)[:, None]
    g2 = (-2 * (y - q2) * self.cw * w / B).astype(f32)[:, None]
    grads = self.c1.backward(tape1, g1) + self.c2.backward(tape2, g2)
    self.critic_opt.apply(self.c1.params() + self.c2.params(), grads)
    # ---- actor step with the updated critics (:364-377)
    head, atape = self.actor.forward(obs, keep=True)
    A = head.shape[1] // 2
    a, logp, u = sample_and_log_prob(head, eps_actor, self.amin, self.amax)
    qa1, t1 = self._q(self.c1, obs, a, keep=True)
    qa2, t2 = self._q(self.c2, obs, a, keep=True)
    aloss = f32(self.aw * actor_loss(qa1, qa2, logp, la, weights))
    k = (self.aw * w / B).astype(f32)
    first = qa1 <= qa2
    dq1 = np.where(first, -k, 0).astype(f32)[:, None]
    dq2 = np.where(first, 0, -k).astype(f32)[:, None]
    dx1 = _input_grad(self.c1, t1, dq1)[:, -A:]
    dx2 = _input_grad(self.c2, t2, dq2)[:, -A:]
    da = dx1 + dx2
    g = (np.exp(f32(la)) * k).astype(f32)
    half = (self.amax - self.amin) / 2
    t = np.tanh(u)
    du = da * half * (1 - t * t) + g[:, None] * 2 * t
    std = np.exp(head[:, A:])
    dhead = np.concatenate([du, du * std * eps_actor - g[:, None]], axis=1).astype(f32)
    self.actor_opt.apply(self.actor.params(), self.actor.backward(atape, dhead))
    # ---- alpha step with the updated actor (:379-390)
    _, logp3, _ = sample_and_log_prob(self.actor.forward(obs), eps_alpha, self.amin, self.amax)
    alloss = f32(self.alw * alpha_loss(logp3, la, self.target_entropy, True, weights))
    dla = f32(self.alw * np.sum(np.where(w == 0, 0, (-logp3 - self.target_entropy) * w), dtype=f32) / B)
    self.alpha_opt.apply([self.log_alpha], [np.array([dla], f32)])
    self.train_step_counter += 1
    self._update_target()
    return dict(loss=f32(closs + aloss + alloss), critic_loss=closs, actor_loss=aloss,
                alpha_loss=alloss)


def _input_grad(net, tape, dy):
  """d(output)/d(input) chain for an all-dense Sequential (no parameter gradients)."""
  from oracle import nn
  for i in range(len(net.layers) - 1, -1, -1):
    l = net.layers[i]
    xin, y = tape[i]
    dz = nn.act_bwd(y, dy, l.get('act'))
    dy = (dz @ l['w'].T).astype(f32)
  return dy
