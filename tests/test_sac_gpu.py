"""SAC on the GPU: reference scalar goldens through the fused kernels, tanh-Normal sampling vs
the oracle, and SacAgent.train parity (losses within 1e-5 relative, supplied noise)."""
import numpy as np
import pytest
import torch

from agents_b200 import _lib
from agents_b200 import optimizers
from agents_b200.agents.sac import sac_agent
from agents_b200.networks import critic_network
from agents_b200.networks import layers as L
from agents_b200.networks import tanh_normal_projection_network as tnp
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import time_step as ts
from agents_b200.trajectories import trajectory
from oracle import nn as onn
from oracle import optim as ooptim
from oracle import sac as osac

pytestmark = pytest.mark.gpu
f32 = np.float32


def _d(cuda, a):
  return torch.as_tensor(np.ascontiguousarray(a, dtype=f32), device=cuda)


def test_reference_goldens_through_kernels(cuda):
  # agents/sac/sac_agent_test.py:269-316 (critic 2*MSE([7.3,19.1],[7,10])), :353-373 (actor 6),
  # :375-396 (alpha -52) with log_pi mocked to 10
  la = torch.zeros(4, device=cuda)
  q, tq = _d(cuda, [7., 10.]), _d(cuda, [7., 9.])
  logpi, rew, disc = _d(cuda, [10., 10.]), _d(cuda, [10., 20.]), _d(cuda, [.9, .9])   # keep alive
  loss = torch.empty(1, device=cuda); dq1 = torch.empty(2, device=cuda); dq2 = torch.empty(2, device=cuda)
  y = torch.empty(2, device=cuda)
  _lib.call('b200rl_sac_critic_loss', _lib.ptr(q), _lib.ptr(q), _lib.ptr(tq), _lib.ptr(tq),
            _lib.ptr(logpi), _lib.ptr(rew), _lib.ptr(disc), None,
            _lib.ptr(la), 2, 1.0, 1.0, 1.0, 2.0, _lib.ptr(loss), _lib.ptr(dq1), _lib.ptr(dq2), _lib.ptr(y), None,
            _lib.stream())
  np.testing.assert_allclose(y.cpu().numpy(), [7.3, 19.1], rtol=1e-6)
  np.testing.assert_allclose(loss.item(), 2 * np.mean((np.array([7.3, 19.1]) - [7., 10.]) ** 2), rtol=1e-6)
  qa = _d(cuda, [3., 5.])
  dl = torch.empty(2, device=cuda)
  _lib.call('b200rl_sac_actor_loss', _lib.ptr(qa), _lib.ptr(qa), _lib.ptr(logpi), None, _lib.ptr(la), 2,
            1.0, 2.0, _lib.ptr(loss), _lib.ptr(dl), _lib.ptr(dq1), _lib.ptr(dq2), None, _lib.stream())
  np.testing.assert_allclose(loss.item(), 6.0, rtol=1e-6)
  la[0] = 4.0
  dla = torch.empty(4, device=cuda)
  _lib.call('b200rl_sac_alpha_loss', _lib.ptr(logpi), None, _lib.ptr(la), 2, 3.0, 1, 1.0, 2.0,
            _lib.ptr(loss), _lib.ptr(dla), None, _lib.stream())
  np.testing.assert_allclose(loss.item(), -52.0, rtol=1e-6)
  np.testing.assert_allclose(dla[0].item(), -13.0, rtol=1e-6)


def test_sample_and_logp_parity(cuda):
  rng = np.random.RandomState(0)
  N, A = 1000, 6
  head = (rng.randn(N, 2 * A) * .7).astype(f32)
  eps = rng.randn(N, A).astype(f32)
  amin, amax = -np.ones(A, f32), np.ones(A, f32) * 3
  wa, wl, wu = osac.sample_and_log_prob(head, eps, amin, amax)
  act = torch.empty(N, A, device=cuda); logp = torch.empty(N, device=cuda)
  u = torch.empty(N, A, device=cuda); e = torch.empty(N, A, device=cuda)
  th, tmin, tmax, teps = _d(cuda, head), _d(cuda, amin), _d(cuda, amax), _d(cuda, eps)   # keep alive
  _lib.call('b200rl_sac_sample', _lib.ptr(th), N, A, _lib.ptr(tmin), _lib.ptr(tmax),
            _lib.ptr(teps), 0, None, _lib.ptr(act), A, _lib.ptr(logp), _lib.ptr(u), _lib.ptr(e), _lib.stream())
  np.testing.assert_allclose(act.cpu().numpy(), wa, rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(logp.cpu().numpy(), wl, rtol=1e-5, atol=1e-4)
  np.testing.assert_allclose(u.cpu().numpy(), wu, rtol=1e-6, atol=1e-6)
  # device-drawn noise is N(0,1) and advances the call counter
  rngs = torch.zeros(2, dtype=torch.int64, device=cuda)
  _lib.call('b200rl_sac_sample', _lib.ptr(th), N, A, _lib.ptr(tmin), _lib.ptr(tmax),
            None, 7, _lib.ptr(rngs), _lib.ptr(act), A, _lib.ptr(logp), _lib.ptr(u), _lib.ptr(e), _lib.stream())
  z = e.cpu().numpy()
  assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05 and rngs.cpu().tolist() == [1, 0]


def _mirror(net):
  return onn.Sequential([dict(kind='dense', w=l.kernel.cpu().numpy().copy(), b=l.bias.cpu().numpy().copy(),
                              act=l.activation) for l in net.layers if isinstance(l, L.Dense)])


def test_sac_agent_train_parity(cuda):
  rng = np.random.RandomState(5)
  D, A, B = 17, 6, 256
  obs_spec = tensor_spec.TensorSpec((D,), torch.float32, 'observation')
  act_spec = tensor_spec.BoundedTensorSpec((A,), torch.float32, -1.0, 1.0, 'action')
  actor = tnp.TanhNormalActorNetwork(obs_spec, act_spec, fc_layer_params=(64, 64), device=cuda).set_seed(1)
  critic = critic_network.CriticNetwork((obs_spec, act_spec), joint_fc_layer_params=(64, 64), device=cuda).set_seed(2)
  agent = sac_agent.SacAgent(ts.time_step_spec(obs_spec), act_spec, critic_network=critic, actor_network=actor,
                             actor_optimizer=optimizers.Adam(3e-4), critic_optimizer=optimizers.Adam(3e-4),
                             alpha_optimizer=optimizers.Adam(3e-4), target_update_tau=0.005,
                             target_update_period=1, gamma=0.99, reward_scale_factor=0.1)
  agent.initialize()
  orc = osac.SacOracle(_mirror(actor), _mirror(agent._critic_network_1), _mirror(agent._critic_network_2),
                       -np.ones(A, f32), np.ones(A, f32), ooptim.AdamTF(3e-4, eps=1e-7),
                       ooptim.AdamTF(3e-4, eps=1e-7), ooptim.AdamTF(3e-4, eps=1e-7), gamma=0.99,
                       reward_scale=0.1, tau=0.005)
  assert not np.array_equal(orc.c1.params()[0], orc.c2.params()[0])     # twin critics differ
  for step in range(4):
    e = dict(observation=rng.randn(B, 2, D).astype(f32), action=(rng.rand(B, 2, A) * 2 - 1).astype(f32),
             reward=rng.rand(B, 2).astype(f32), discount=(rng.rand(B, 2) > .05).astype(f32))
    noise = [rng.randn(B, A).astype(f32) for _ in range(3)]
    want = orc.train(e, *noise)
    d = lambda a: torch.as_tensor(a, device=cuda)
    traj = trajectory.Trajectory(torch.ones(B, 2, dtype=torch.int32, device=cuda), d(e['observation']), d(e['action']), (),
                                 torch.ones(B, 2, dtype=torch.int32, device=cuda), d(e['reward']), d(e['discount']))
    got = agent.train(traj, noise=noise)
    np.testing.assert_allclose(got.extra.critic_loss.item(), want['critic_loss'], rtol=1e-5)
    np.testing.assert_allclose(got.extra.actor_loss.item(), want['actor_loss'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got.extra.alpha_loss.item(), want['alpha_loss'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got.loss.item(), want['loss'], rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(agent.log_alpha.item(), orc.log_alpha[0], rtol=1e-4, atol=1e-7)
  for v, w in zip(actor.variables, orc.actor.params()):
    np.testing.assert_allclose(v.cpu().numpy(), w, rtol=1e-3, atol=1e-5)
  for v, w in zip(agent._target_critic_network_1.variables, orc.t1.params()):
    np.testing.assert_allclose(v.cpu().numpy(), w, rtol=1e-4, atol=1e-6)
  assert int(agent.train_step_counter.item()) == 4
  agent.check_numerics()
  # device-noise path + policies
  traj_dev = agent.train(traj)
  assert np.isfinite(traj_dev.loss.item())
  step = agent.collect_policy.action(ts.restart(d(e['observation'][:, 0]), batch_size=B))
  assert tuple(step.action.shape) == (B, A) and float(step.action.abs().max()) <= 1.0
  g1 = agent.policy.action(ts.restart(d(e['observation'][:, 0]), batch_size=B)).action
  g2 = agent.policy.action(ts.restart(d(e['observation'][:, 0]), batch_size=B)).action
  assert torch.equal(g1, g2)
