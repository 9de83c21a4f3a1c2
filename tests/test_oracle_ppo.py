"""Pins oracle/ppo.py to the reference goldens of agents/ppo/ppo_agent_test.py,
agents/ppo/ppo_utils_test.py and utils/tensor_normalizer_test.py, and checks its hand-written
backward against torch CPU autograd."""
import numpy as np
import torch

from oracle import nn as onn
from oracle import optim as ooptim
from oracle import ppo as oppo

f32 = np.float32

# DummyActorNet (ppo_agent_test.py:50-128): Dense(2) kernel [[2,1],[1,1]], bias [5,5] -> (loc, scale)
OBS = np.array([[1, 2], [3, 4]], f32)
LOC = (OBS @ np.array([[2.], [1.]], f32) + 5).astype(f32)       # [[9],[15]]
SCALE = (OBS @ np.array([[1.], [1.]], f32) + 5).astype(f32)     # [[8],[12]]
VALUE = (OBS @ np.array([2., 1.], f32) + 5).astype(f32)         # DummyValueNet :151-171 -> [9, 15]


def test_policy_gradient_loss_golden():  # ppo_agent_test.py:947-985 -> -0.0164646133
  logp = oppo.normal_log_prob(LOC, SCALE, np.array([[0.], [1.]], f32))
  loss, clip_frac = oppo.policy_gradient_loss(logp, np.array([.9, .3], f32), np.array([1.9, 1.], f32),
                                              np.ones(2, f32), clip_eps=10.0)
  np.testing.assert_allclose(loss, -0.0164646133, rtol=1e-5)
  assert clip_frac == 0


def test_value_estimation_loss_golden():  # :919-942 -> 123.205
  loss = oppo.value_estimation_loss(VALUE, np.array([1.9, 1.0], f32), np.ones(2, f32), vf_coef=1.0)
  np.testing.assert_allclose(loss, 123.205, rtol=1e-6)


def test_entropy_regularization_loss_golden():  # :864-914 -> -3.70111 * coef
  ent = oppo.normal_entropy(SCALE)
  np.testing.assert_allclose(oppo.entropy_regularization_loss(ent, np.ones(2, f32), 0.1), -0.370111,
                             rtol=1e-5)
  assert oppo.entropy_regularization_loss(ent, np.ones(2, f32), 0.0) == 0


def test_get_loss_components_golden():  # :644-727: time dimension of 2 -> every term x 2/4
  # [B=2, T=2] batch made of the same two rows repeated over time, weights zero at t=1
  sh = lambda x: np.stack([x, x], axis=1)
  w = np.array([[1, 0], [1, 0]], f32)
  logp = oppo.normal_log_prob(sh(LOC), sh(SCALE), sh(np.array([[0.], [1.]], f32)))
  pg, _ = oppo.policy_gradient_loss(logp, sh(np.array([.9, .3], f32)), sh(np.array([1.9, 1.], f32)), w, 10.0)
  np.testing.assert_allclose(pg, -0.0164646133 * 2 / 4, rtol=1e-5)
  ve = oppo.value_estimation_loss(sh(VALUE), sh(np.array([1.9, 1.0], f32)), w, 1.0)
  np.testing.assert_allclose(ve, 123.205 * 2 / 4, rtol=1e-6)
  en = oppo.entropy_regularization_loss(oppo.normal_entropy(sh(SCALE)), w, 0.1)
  np.testing.assert_allclose(en, -0.370111 * 2 / 4, rtol=1e-5)


def test_trajectory_mask():  # ppo_utils_test.py:35-64
  st = np.array([[0, 1, 2, 0], [1, 1, 1, 2]], np.int32)
  ret = np.array([[1, 2, 3, 0], [1, 0, 3, 0]], f32)
  adv = np.array([[1, 2, 3, 0], [1, 0, 3, 1]], f32)
  m = oppo.make_trajectory_mask(st, ret, adv)
  assert m.tolist() == [[1, 1, 0, 0], [1, 0, 1, 0]]


def test_return_and_advantage_use_value_ops_with_sic_final_value():
  # ppo_agent_test.py:290-347 drives GAE through the agent; here: structure of :440-479, :617-719
  rng = np.random.RandomState(0)
  B, T = 3, 6
  r, d = rng.rand(B, T - 1).astype(f32), np.ones((B, T - 1), f32)
  nst = np.ones((B, T - 1), np.int32)
  nst[0, 2] = 2
  vp = rng.rand(B, T).astype(f32)
  ret, adv = oppo.compute_return_and_advantage(r, d, nst, vp, gamma=0.99, lam=0.95)
  assert ret.shape == adv.shape == (B, T - 1)
  # last step of the return bootstraps from V[:, -1]; the GAE bootstraps from V[:, -2] (sic)
  np.testing.assert_allclose(ret[:, -1], r[:, -1] + 0.99 * vp[:, -1], rtol=1e-6)
  np.testing.assert_allclose(adv[:, -1], r[:, -1] + 0.99 * vp[:, -2] - vp[:, -2], rtol=1e-5, atol=1e-6)
  # the episode end zeroes the discount at (0, 2)
  np.testing.assert_allclose(ret[0, 2], r[0, 2], rtol=1e-6)


def test_streaming_normalizer_matches_batch_stats():  # utils/tensor_normalizer_test.py:33-91,253-479
  rng = np.random.RandomState(1)
  x = (rng.randn(4000, 3) * [1, 5, .1] + [0, 2, -1]).astype(f32)
  n = oppo.StreamingNormalizer((3,))
  for chunk in np.split(x, 8):
    n.update(chunk)
  np.testing.assert_allclose(n.avg, x.mean(0), rtol=1e-4, atol=1e-5)
  np.testing.assert_allclose(n.m2 / n.count, x.var(0), rtol=1e-4)
  out = n.normalize(x, clip_value=5.0)
  assert abs(out.mean()) < 1e-2 and abs(out.std() - 1) < 2e-2 and out.max() <= 5.0
  np.testing.assert_allclose(n.normalize(x[:2], clip_value=0, center_mean=False),
                             x[:2] / np.sqrt(x.var(0) + 1e-3), rtol=1e-3)


def _make_oracle(rng, obs_dim=5, A=3, hidden=(8, 6), **kw):
  def mlp(out, scale_last):
    layers, n_in = [], obs_dim
    for h in hidden:
      layers.append(dict(kind='dense', w=(rng.randn(n_in, h) * .4).astype(f32), b=(rng.randn(h) * .1).astype(f32), act='tanh'))
      n_in = h
    layers.append(dict(kind='dense', w=(rng.randn(n_in, out) * scale_last).astype(f32), b=np.zeros(out, f32), act=None))
    return onn.Sequential(layers)
  if kw.get('normalize_observations'):
    kw['obs_dim'] = obs_dim
  return oppo.PPOOracle(mlp(A, .3), (rng.randn(A) * .2).astype(f32), mlp(1, .3), -np.ones(A, f32) * 2,
                        np.ones(A, f32) * 2, ooptim.AdamTF(1e-3, eps=1e-7), **kw)


def test_oracle_backward_matches_autograd():
  rng = np.random.RandomState(2)
  B, T, A, D = 4, 5, 3, 5
  orc = _make_oracle(rng, D, A, clip_eps=0.2, vf_coef=0.5, ent_coef=0.01, value_clip=0.3, logp_clip=8.0)
  N = B * T
  obs = rng.randn(N, D).astype(f32)
  action = rng.randn(N, A).astype(f32)
  old_logp = (rng.randn(N) * .5 - 3).astype(f32)
  ret, adv, v_old = rng.randn(N).astype(f32), rng.randn(N).astype(f32), rng.randn(N).astype(f32) * .3
  w = (rng.rand(N) > .2).astype(f32)
  info, grads = orc.loss_and_grads(obs, action, old_logp, ret, adv, v_old, w, B, T)
  # torch re-statement with autograd
  tp = [torch.tensor(p, requires_grad=True) for p in orc.params()]
  na = len(orc.actor.params())
  def run(params, x):
    for i in range(0, len(params), 2):
      x = x @ params[i] + params[i + 1]
      if i + 2 < len(params):
        x = torch.tanh(x)
    return x
  xo = torch.tensor(obs)
  m_raw = run(tp[:na], xo)
  loc = 2.0 * torch.tanh(m_raw)
  scale = torch.nn.functional.softplus(tp[na]).expand_as(loc)
  v = run(tp[na + 1:], xo)[:, 0]
  dist = torch.distributions.Normal(loc, scale)
  logp = dist.log_prob(torch.tensor(action)).sum(-1)
  ent = dist.entropy().sum(-1)
  lp = torch.clamp(logp, -8.0, 8.0)
  ratio = torch.exp(lp - torch.tensor(old_logp))
  a_t, w_t = torch.tensor(adv), torch.tensor(w)
  pg = (-torch.minimum(ratio * a_t, torch.clamp(ratio, .8, 1.2) * a_t) * w_t).reshape(B, T).mean(1).sum() / B
  r_t, vo = torch.tensor(ret), torch.tensor(v_old)
  vc = vo + torch.clamp(v - vo, -.3, .3)
  ve = .5 * (torch.maximum((r_t - v) ** 2, (r_t - vc) ** 2) * w_t).reshape(B, T).mean(1).sum() / B
  en = .01 * (-ent * w_t).reshape(B, T).mean(1).sum() / B
  total = pg + ve + en
  np.testing.assert_allclose(info['loss'], total.item(), rtol=2e-5)
  np.testing.assert_allclose(info['pg'], pg.item(), rtol=2e-5)
  total.backward()
  for g, p in zip(grads, tp):
    np.testing.assert_allclose(g, p.grad.numpy(), rtol=2e-4, atol=2e-6)


def test_oracle_train_runs_epochs():
  rng = np.random.RandomState(3)
  B, T, A, D = 6, 7, 2, 4
  orc = _make_oracle(rng, D, A, num_epochs=3, clip_eps=0.2)
  exp = dict(observation=rng.randn(B, T, D).astype(f32), action=rng.randn(B, T, A).astype(f32),
             loc=rng.randn(B, T, A).astype(f32) * .1, scale=np.full((B, T, A), .8, f32),
             reward=rng.rand(B, T).astype(f32), discount=np.ones((B, T), f32),
             step_type=np.ones((B, T), np.int32), next_step_type=np.ones((B, T), np.int32))
  infos = orc.train(exp)
  assert len(infos) == 3 == orc.train_step_counter
  assert all(np.isfinite(i['loss']) for i in infos) and infos[0]['loss'] != infos[-1]['loss']


# ---- KL penalty (ppo_agent.py:1514-1690) ---------------------------------------------------------
def test_kl_cutoff_loss_golden():  # ppo_agent_test.py:1045-1078: coef * (0.74 - 0.5)^2, 0 when coef 0
  kl = np.array([[1.5, -0.5, 6.5, -1.5, -2.3]], f32)
  for coef in (0.0, 30.0):
    got = oppo.kl_cutoff_loss(kl, kl_cutoff_factor=5.0, adaptive_kl_target=0.1, kl_cutoff_coef=coef)
    np.testing.assert_allclose(got, coef * 0.24 ** 2, rtol=1e-5)
  assert oppo.kl_cutoff_loss(kl, 0.0, 0.1, 30.0) == 0.0                       # :1518-1519


def test_adaptive_kl_loss_and_beta_update_goldens():  # :1080-1164
  beta = f32(1.0)
  # :1113-1124 loss moves with beta; :1149-1164 beta 1.0 -> 1.0 -> 1.5 -> 1.0
  assert oppo.adaptive_kl_loss(np.array([10.0], f32), beta) == 10.0
  beta = oppo.update_adaptive_kl_beta(beta, np.array([10.0], f32), 10.0, 0.5)
  assert beta == 1.0
  beta = oppo.update_adaptive_kl_beta(beta, np.array([100.0], f32), 10.0, 0.5)
  assert beta == 1.5
  assert oppo.adaptive_kl_loss(np.array([100.0], f32), beta) > 100.0
  beta = oppo.update_adaptive_kl_beta(beta, np.array([1.0], f32), 10.0, 0.5)
  np.testing.assert_allclose(beta, 1.0, rtol=1e-6)
  assert oppo.update_adaptive_kl_beta(None, np.array([1.0], f32), 10.0, 0.5) is None
  assert oppo.adaptive_kl_loss(np.array([1.0], f32), None) == 0.0


def test_normal_kl_matches_torch_distributions():
  rng = np.random.RandomState(5)
  la, lb = rng.randn(7, 3).astype(f32), rng.randn(7, 3).astype(f32)
  sa, sb = (rng.rand(7, 3) + .3).astype(f32), (rng.rand(7, 3) + .3).astype(f32)
  want = torch.distributions.kl_divergence(torch.distributions.Normal(torch.tensor(la), torch.tensor(sa)),
                                           torch.distributions.Normal(torch.tensor(lb), torch.tensor(sb))).sum(-1)
  np.testing.assert_allclose(oppo.normal_kl(la, sa, lb, sb), want.numpy(), rtol=2e-5, atol=1e-6)


def test_oracle_kl_penalty_backward_matches_autograd():
  rng = np.random.RandomState(7)
  B, T, A, D = 4, 5, 3, 5
  kw = dict(kl_cutoff_factor=2.0, kl_cutoff_coef=50.0, initial_adaptive_kl_beta=1.3,
            adaptive_kl_target=0.01, adaptive_kl_tolerance=0.3)
  orc = _make_oracle(rng, D, A, clip_eps=0.2, vf_coef=0.5, **kw)
  N = B * T
  obs = rng.randn(N, D).astype(f32)
  action = rng.randn(N, A).astype(f32)
  old_loc, old_scale = (rng.randn(N, A) * .3).astype(f32), (rng.rand(N, A) * .5 + .5).astype(f32)
  old_logp = oppo.normal_log_prob(old_loc, old_scale, action)
  ret, adv = rng.randn(N).astype(f32), rng.randn(N).astype(f32)
  w = (rng.rand(N) > .2).astype(f32)
  info, grads = orc.loss_and_grads(obs, action, old_logp, ret, adv, None, w, B, T, old_loc=old_loc,
                                   old_scale=old_scale)
  tp = [torch.tensor(p, requires_grad=True) for p in orc.params()]
  na = len(orc.actor.params())

  def run(params, x):
    for i in range(0, len(params), 2):
      x = x @ params[i] + params[i + 1]
      if i + 2 < len(params):
        x = torch.tanh(x)
    return x
  xo = torch.tensor(obs)
  loc = 2.0 * torch.tanh(run(tp[:na], xo))
  scale = torch.nn.functional.softplus(tp[na]).expand_as(loc)
  v = run(tp[na + 1:], xo)[:, 0]
  dist = torch.distributions.Normal(loc, scale)
  logp = dist.log_prob(torch.tensor(action)).sum(-1)
  ratio = torch.exp(logp - torch.tensor(old_logp))
  a_t, w_t = torch.tensor(adv), torch.tensor(w)
  pg = (-torch.minimum(ratio * a_t, torch.clamp(ratio, .8, 1.2) * a_t) * w_t).reshape(B, T).mean(1).sum() / B
  ve = .5 * (((torch.tensor(ret) - v) ** 2) * w_t).reshape(B, T).mean(1).sum() / B
  kl = torch.distributions.kl_divergence(
      torch.distributions.Normal(torch.tensor(old_loc), torch.tensor(old_scale)), dist).sum(-1) * w_t
  mean_kl = kl.mean()
  klp = 50.0 * torch.clamp(mean_kl - 2.0 * 0.01, min=0.0) ** 2 + 1.3 * mean_kl
  total = pg + ve + klp
  assert klp.item() > 1e-3                      # the penalty is active in this case
  np.testing.assert_allclose(info['kl'], klp.item(), rtol=2e-5)
  np.testing.assert_allclose(info['loss'], total.item(), rtol=2e-5)
  total.backward()
  for g, p in zip(grads, tp):
    np.testing.assert_allclose(g, p.grad.numpy(), rtol=2e-4, atol=2e-6)


def test_oracle_train_with_kl_and_observation_normalizer():
  rng = np.random.RandomState(11)
  B, T, A, D = 6, 7, 2, 4
  orc = _make_oracle(rng, D, A, num_epochs=3, clip_eps=0.0, kl_cutoff_factor=2.0, kl_cutoff_coef=1000.0,
                     initial_adaptive_kl_beta=1.0, adaptive_kl_target=0.01, adaptive_kl_tolerance=0.3,
                     normalize_observations=True, normalize_rewards=True)
  exp = dict(observation=rng.randn(B, T, D).astype(f32), action=rng.randn(B, T, A).astype(f32),
             loc=rng.randn(B, T, A).astype(f32) * .1, scale=np.full((B, T, A), .8, f32),
             reward=rng.rand(B, T).astype(f32), discount=np.ones((B, T), f32),
             step_type=np.ones((B, T), np.int32), next_step_type=np.ones((B, T), np.int32))
  infos = orc.train(exp)
  assert len(infos) == 3 and all(np.isfinite(i['loss']) for i in infos)
  assert infos[0]['kl'] > 0 and orc.beta != 1.0          # far from the behaviour policy -> beta moved
  assert orc.obs_normalizer.count[0] > 1 and orc.reward_normalizer.count > 1
