"""Pins oracle/sac.py to the scalar goldens of agents/sac/sac_agent_test.py and checks the
tanh-Normal log-prob and the hand-written actor gradient against torch."""
import numpy as np
import torch

from oracle import nn as onn
from oracle import optim as ooptim
from oracle import sac as osac

f32 = np.float32


def _dummy_critic(obs, act):
  # DummyCriticNet (sac_agent_test.py:84-130): q = obs @ [0, 1] + act @ [1]
  return (obs[:, 1] + act[:, 0]).astype(f32)


def test_critic_loss_golden():  # :269-316: 2 * MSE([7.3, 19.1], [7, 10])
  obs, act = np.array([[1, 2], [3, 4]], f32), np.array([[5], [6]], f32)
  obs2, next_act = np.array([[5, 6], [7, 8]], f32), np.ones((2, 1), f32)     # DummyActorPolicy: spec.maximum
  q = _dummy_critic(obs, act)
  tq = _dummy_critic(obs2, next_act)
  loss, y = osac.critic_loss(q, q, tq, tq, np.full(2, 10, f32), np.array([10, 20], f32),
                             np.array([.9, .9], f32), log_alpha=0.0)
  np.testing.assert_allclose(y, [7.3, 19.1], rtol=1e-6)
  np.testing.assert_allclose(loss, 2 * np.mean((np.array([7.3, 19.1]) - np.array([7., 10.])) ** 2), rtol=1e-6)


def test_actor_loss_golden():  # :353-373: (2*10 - (2+1) - (4+1)) / 2 = 6
  q = _dummy_critic(np.array([[1, 2], [3, 4]], f32), np.ones((2, 1), f32))
  np.testing.assert_allclose(osac.actor_loss(q, q, np.full(2, 10, f32), 0.0), 6.0, rtol=1e-6)


def test_alpha_loss_golden():  # :375-396: 4 * (-10 - 3) = -52
  np.testing.assert_allclose(osac.alpha_loss(np.full(2, 10, f32), 4.0, 3.0), -52.0, rtol=1e-6)


def test_tanh_normal_log_prob_matches_torch():
  rng = np.random.RandomState(0)
  N, A = 50, 6
  head = (rng.randn(N, 2 * A) * .5).astype(f32)
  eps = rng.randn(N, A).astype(f32)
  amin, amax = -np.ones(A, f32) * 2, np.ones(A, f32) * 2
  a, logp, u = osac.sample_and_log_prob(head, eps, amin, amax)
  # float64 reference: log|da/du| = log(half) + log(1 - tanh(u)^2) = log(half) + log(sech(u)^2)
  h64 = torch.tensor(head.astype(np.float64))
  base = torch.distributions.Normal(h64[:, :A], torch.exp(h64[:, A:]))
  tu = h64[:, :A] + torch.exp(h64[:, A:]) * torch.tensor(eps.astype(np.float64))
  log_det = np.log(2.0) + 2 * (np.log(2.0) - torch.logaddexp(tu, -tu))
  want = (base.log_prob(tu) - log_det).sum(-1)
  np.testing.assert_allclose(logp, want.numpy(), rtol=2e-5, atol=2e-5)
  assert np.abs(a).max() <= 2.0


def _mlp(rng, n_in, hidden, n_out):
  layers = []
  for h in hidden:
    layers.append(dict(kind='dense', w=(rng.randn(n_in, h) * .3).astype(f32), b=(rng.randn(h) * .05).astype(f32), act='relu'))
    n_in = h
  layers.append(dict(kind='dense', w=(rng.randn(n_in, n_out) * .3).astype(f32), b=np.zeros(n_out, f32), act=None))
  return onn.Sequential(layers)


def make_oracle(rng, D=5, A=3, hidden=(16, 16), lr=3e-4):
  return osac.SacOracle(_mlp(rng, D, hidden, 2 * A), _mlp(rng, D + A, hidden, 1), _mlp(rng, D + A, hidden, 1),
                        -np.ones(A, f32), np.ones(A, f32), ooptim.AdamTF(lr, eps=1e-7),
                        ooptim.AdamTF(lr, eps=1e-7), ooptim.AdamTF(lr, eps=1e-7), gamma=0.99,
                        reward_scale=0.1, tau=0.005)


def test_actor_gradient_matches_autograd():
  rng = np.random.RandomState(1)
  D, A, B = 5, 3, 12
  orc = make_oracle(rng, D, A)
  obs = rng.randn(B, D).astype(f32)
  eps = rng.randn(B, A).astype(f32)
  # torch restatement of the actor loss
  ap = [torch.tensor(p, requires_grad=True) for p in orc.actor.params()]
  def run(params, x):
    for i in range(0, len(params), 2):
      x = x @ params[i] + params[i + 1]
      if i + 2 < len(params):
        x = torch.relu(x)
    return x
  head = run(ap, torch.tensor(obs))
  loc, ls = head[:, :A], head[:, A:]
  u = loc + torch.exp(ls) * torch.tensor(eps)
  a = torch.tanh(u)
  logp = (-0.5 * torch.tensor(eps) ** 2 - ls - 0.5 * np.log(2 * np.pi)
          - 2 * (np.log(2.0) - u - torch.nn.functional.softplus(-2 * u))).sum(-1)
  cp1 = [torch.tensor(p) for p in orc.c1.params()]
  cp2 = [torch.tensor(p) for p in orc.c2.params()]
  x = torch.cat([torch.tensor(obs), a], 1)
  q1, q2 = run(cp1, x)[:, 0], run(cp2, x)[:, 0]
  loss = (logp - torch.minimum(q1, q2)).mean()
  loss.backward()
  want = [p.grad.numpy() for p in ap]
  # oracle: run only the actor part of train() by replaying its formulas
  before = [p.copy() for p in orc.actor.params()]
  class Capture(object):
    def apply(self, params, grads):
      self.grads = [g.copy() for g in grads]
  cap = Capture()
  orc.actor_opt = cap
  orc.critic_opt = Capture()
  orc.alpha_opt = Capture()
  exp = dict(observation=np.stack([obs, obs], 1), action=np.zeros((B, 2, A), f32),
             reward=np.zeros((B, 2), f32), discount=np.ones((B, 2), f32))
  out = orc.train(exp, eps, eps, eps)
  np.testing.assert_allclose(out['actor_loss'], loss.item(), rtol=2e-5)
  for g, w in zip(cap.grads, want):
    np.testing.assert_allclose(g, w, rtol=2e-4, atol=2e-6)
  for p, b in zip(orc.actor.params(), before):
    assert np.array_equal(p, b)


def test_train_order_and_targets():
  rng = np.random.RandomState(2)
  D, A, B = 5, 3, 8
  orc = make_oracle(rng, D, A, lr=1e-2)
  t_before = [p.copy() for p in orc.t1.params()]
  c_before = [p.copy() for p in orc.c1.params()]
  exp = dict(observation=rng.randn(B, 2, D).astype(f32), action=rng.rand(B, 2, A).astype(f32) * 2 - 1,
             reward=rng.rand(B, 2).astype(f32), discount=np.ones((B, 2), f32))
  out = orc.train(exp, *[rng.randn(B, A).astype(f32) for _ in range(3)])
  assert np.isfinite(out['loss']) and orc.train_step_counter == 1
  for t, tb, c, cb in zip(orc.t1.params(), t_before, orc.c1.params(), c_before):
    np.testing.assert_allclose(t, 0.995 * tb + 0.005 * c, rtol=1e-5, atol=1e-7)   # Polyak after the step
    assert not np.array_equal(c, cb)
  assert orc.log_alpha[0] != 0.0
