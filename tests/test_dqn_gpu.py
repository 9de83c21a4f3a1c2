"""DqnAgent / DdqnAgent on the GPU: reference scalar-loss goldens through the product API and
multi-step train parity against the oracle (fp32 losses within 1e-5 relative)."""
import numpy as np
import pytest
import torch

from agents_b200 import _lib
from agents_b200 import optimizers
from agents_b200.agents.dqn import dqn_agent
from agents_b200.networks import layers as L
from agents_b200.networks import q_network
from agents_b200.networks import sequential
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import time_step as ts
from agents_b200.trajectories import trajectory
from agents_b200.utils import common
from oracle import dqn as odqn
from oracle import nn as onn
from oracle import optim as ooptim

pytestmark = pytest.mark.gpu
f32 = np.float32
AGENTS = [dqn_agent.DqnAgent, dqn_agent.DdqnAgent]


def _dummy_net(cuda, l2=0.0):
  # agents/dqn/dqn_agent_test.py:38-69 DummyNet
  return sequential.Sequential(
      [L.Dense(2, kernel_initializer=np.array([[2, 1], [1, 1]], f32), bias_initializer=[1, 1],
               kernel_regularizer_l2=l2)], device=cuda)


def _specs(with_mask=False):
  obs = tensor_spec.TensorSpec([2], torch.float32)
  if with_mask:
    obs = (obs, tensor_spec.BoundedTensorSpec([2], torch.int32, 0, 1))
  return ts.time_step_spec(obs), tensor_spec.BoundedTensorSpec((), torch.int32, 0, 1)


def _exp(cuda, obs_seq, step_types, rewards, discounts, masks=None):
  B, T = 2, len(obs_seq)
  t = lambda a, dt: torch.as_tensor(np.stack(a, axis=1), device=cuda).to(dt)
  obs = t([np.asarray(o, f32) for o in obs_seq], torch.float32)
  if masks is not None:
    obs = (obs, t([np.asarray(m, np.int32) for m in masks], torch.int32))
  return trajectory.Trajectory(
      step_type=t([np.full(B, s) for s in step_types], torch.int32), observation=obs,
      action=t([np.array([0, 1])] * T, torch.int32), policy_info=(),
      next_step_type=t([np.full(B, 1)] * T, torch.int32),
      reward=t([np.asarray(r, f32) for r in rewards], torch.float32),
      discount=t([np.asarray(d, f32) for d in discounts], torch.float32))


@pytest.mark.parametrize('agent_class', AGENTS)
def test_loss_goldens(cuda, agent_class):
  """agents/dqn/dqn_agent_test.py: :178-218 26.0, :220-267 9.8, :269-299 33.0, :301-355 47.42,
  :416-481 21.5."""
  tss, acts = _specs()
  two = [[[1, 2], [3, 4]], [[5, 6], [7, 8]]]
  agent = agent_class(tss, acts, q_network=_dummy_net(cuda), optimizer=None)
  agent.initialize()
  loss = agent.loss(_exp(cuda, two, [0, 1], [[10, 20]] * 2, [[.9, .9]] * 2)).loss
  np.testing.assert_allclose(loss.item(), 26.0, rtol=1e-6)
  loss = agent.loss(_exp(cuda, [two[0], [[-5, 6], [-7, 8]]], [0, 1], [[10, 20]] * 2, [[.9, .9]] * 2)).loss
  np.testing.assert_allclose(loss.item(), 9.8, rtol=1e-6)
  agent = agent_class(tss, acts, q_network=_dummy_net(cuda, l2=1.0), optimizer=None)
  agent.initialize()
  loss = agent.loss(_exp(cuda, two, [0, 1], [[10, 20]] * 2, [[.9, .9]] * 2)).loss
  np.testing.assert_allclose(loss.item(), 33.0, rtol=1e-6)
  agent = agent_class(tss, acts, q_network=_dummy_net(cuda), optimizer=None, n_step_update=2)
  agent.initialize()
  exp = _exp(cuda, two + [[[9, 10], [11, 12]]], [0, 1, 1], [[10, 20]] * 3, [[.9, .9]] * 3)
  np.testing.assert_allclose(agent.loss(exp).loss.item(), 47.42, rtol=1e-6)
  agent = agent_class(tss, acts, q_network=_dummy_net(cuda), optimizer=None, n_step_update=3)
  agent.initialize()
  exp = _exp(cuda, two + [[[9, 10], [11, 12]], [[13, 14], [15, 16]]], [1, 1, 2, 0],
             [[10, 20], [10, 20], [0, 0], [0, 0]], [[.9, .9], [0, 0], [1, 1], [1, 1]])
  np.testing.assert_allclose(agent.loss(exp).loss.item(), 21.5, rtol=1e-6)


@pytest.mark.parametrize('agent_class', AGENTS)
def test_loss_masked_actions_golden(cuda, agent_class):  # :483-561 -> 23.75
  tss, acts = _specs(with_mask=True)
  agent = agent_class(tss, acts, q_network=_dummy_net(cuda), optimizer=None,
                      observation_and_action_constraint_splitter=lambda x: (x[0], x[1]))
  agent.initialize()
  exp = _exp(cuda, [[[1, 2], [3, 4]], [[5, 6], [7, 8]]], [0, 1], [[10, 20]] * 2, [[.9, .9]] * 2,
             masks=[[[1, 1], [1, 1]], [[0, 1], [1, 0]]])
  np.testing.assert_allclose(agent.loss(exp).loss.item(), 23.75, rtol=1e-6)


def test_d3qn_goldens(cuda):
  """D3qnAgent (dqn_agent.py:704-753): the unmasked goldens like DDQN, and 26.0 on the masked
  case because its argmax ignores the action constraint (dqn_agent_test.py:483-561, :556)."""
  tss, acts = _specs()
  two = [[[1, 2], [3, 4]], [[5, 6], [7, 8]]]
  agent = dqn_agent.D3qnAgent(tss, acts, q_network=_dummy_net(cuda), optimizer=None)
  agent.initialize()
  loss = agent.loss(_exp(cuda, two, [0, 1], [[10, 20]] * 2, [[.9, .9]] * 2)).loss
  np.testing.assert_allclose(loss.item(), 26.0, rtol=1e-6)
  loss = agent.loss(_exp(cuda, [two[0], [[-5, 6], [-7, 8]]], [0, 1], [[10, 20]] * 2, [[.9, .9]] * 2)).loss
  np.testing.assert_allclose(loss.item(), 9.8, rtol=1e-6)
  tss, acts = _specs(with_mask=True)
  agent = dqn_agent.D3qnAgent(tss, acts, q_network=_dummy_net(cuda), optimizer=None,
                              observation_and_action_constraint_splitter=lambda x: (x[0], x[1]))
  agent.initialize()
  exp = _exp(cuda, two, [0, 1], [[10, 20]] * 2, [[.9, .9]] * 2,
             masks=[[[1, 1], [1, 1]], [[0, 1], [1, 0]]])
  np.testing.assert_allclose(agent.loss(exp).loss.item(), 26.0, rtol=1e-6)


def test_sequence_length_validation(cuda):
  tss, acts = _specs()
  agent = dqn_agent.DqnAgent(tss, acts, q_network=_dummy_net(cuda), optimizer=None)
  exp = _exp(cuda, [[[1, 2], [3, 4]]] * 3, [0, 1, 1], [[1, 1]] * 3, [[1, 1]] * 3)
  with pytest.raises(ValueError, match='sequence_length'):
    agent.loss(exp)
  with pytest.raises(ValueError, match='Only scalar actions'):
    dqn_agent.DqnAgent(tss, tensor_spec.BoundedTensorSpec((2,), torch.int32, 0, 1),
                       q_network=_dummy_net(cuda), optimizer=None)


@pytest.mark.parametrize('loss_kind', ['huber', 'squared'])
@pytest.mark.parametrize('B,A,T', [(2, 2, 2), (64, 2, 2), (256, 6, 2), (300, 18, 4), (1025, 3, 3)])
def test_td_loss_kernel_parity(cuda, loss_kind, B, A, T):
  rng = np.random.RandomState(B + A + T)
  q, nt, ns = [rng.randn(B, A).astype(f32) * 3 for _ in range(3)]
  actions = rng.randint(0, A, size=B).astype(np.int32)
  st0 = rng.randint(0, 3, size=B).astype(np.int32)
  rew = rng.randn(B, T).astype(f32)
  disc = (rng.rand(B, T) > 0.2).astype(f32)
  w = rng.rand(B).astype(f32)
  w[::7] = 0
  mask = (rng.rand(B, A) > 0.3).astype(np.int32)
  mask[:, 0] = 1
  for use_w, use_m in [(False, False), (True, True)]:
    want = odqn.dqn_loss(q, nt, ns, actions, st0, rew, disc, gamma=0.99, reward_scale=0.5,
                         loss_fn=loss_kind, weights=w if use_w else None,
                         next_mask=mask if use_m else None)
    d = lambda a: torch.as_tensor(a, device=cuda)
    loss = torch.empty(1, device=cuda); tdl = torch.empty(B, device=cuda)
    tde = torch.empty(B, device=cuda); dq = torch.empty(B, A, device=cuda)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    tq, tnt, tns, ta, ts0, tr, td_, tw, tm = map(d, (q, nt, ns, actions, st0, rew, disc, w, mask))
    _lib.call('b200rl_dqn_td_loss', _lib.ptr(tq), _lib.ptr(tnt), _lib.ptr(tns),
              _lib.ptr(tm) if use_m else None, _lib.ptr(ta), _lib.ptr(ts0), _lib.ptr(tr), _lib.ptr(td_),
              _lib.ptr(tw) if use_w else None, 1, 1, B, A, T, 0.99, 0.5,
              _lib.LOSS_HUBER if loss_kind == 'huber' else _lib.LOSS_SQUARED, float(B),
              _lib.ptr(loss), _lib.ptr(tdl), _lib.ptr(tde), _lib.ptr(dq), _lib.ptr(flag), _lib.stream())
    np.testing.assert_allclose(loss.item(), want['weighted'], rtol=1e-5)
    np.testing.assert_allclose(tdl.cpu().numpy(), want['td_loss'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(tde.cpu().numpy(), want['td_error'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dq.cpu().numpy(), want['dq'], rtol=1e-6, atol=1e-9)
    assert flag.item() == 0


def _mirror_oracle(net):
  olayers = []
  for l in net.layers:
    if isinstance(l, L.CastScale): olayers.append(dict(kind='cast_scale', divisor=l.divisor))
    elif isinstance(l, L.Conv2D): olayers.append(dict(kind='conv', w=l.kernel.cpu().numpy().copy(), b=l.bias.cpu().numpy().copy(), stride=l.stride, act=l.activation))
    elif isinstance(l, L.Flatten): olayers.append(dict(kind='flatten'))
    else: olayers.append(dict(kind='dense', w=l.kernel.cpu().numpy().copy(), b=l.bias.cpu().numpy().copy(), act=l.activation))
  return onn.Sequential(olayers)


def _random_exp(rng, B, T, obs_shape, A, u8):
  obs = rng.randint(0, 256, size=(B, T) + obs_shape).astype(np.uint8) if u8 else rng.randn(B, T, *obs_shape).astype(f32)
  return dict(observation=obs, step_type=rng.randint(0, 3, size=(B, T)).astype(np.int32),
              action=rng.randint(0, A, size=(B, T)).astype(np.int32),
              next_step_type=rng.randint(0, 3, size=(B, T)).astype(np.int32),
              reward=rng.rand(B, T).astype(f32), discount=(rng.rand(B, T) > 0.1).astype(f32))


def _to_traj(cuda, e):
  d = lambda a: torch.as_tensor(a, device=cuda)
  return trajectory.Trajectory(d(e['step_type']), d(e['observation']), d(e['action']), (),
                               d(e['next_step_type']), d(e['reward']), d(e['discount']))


@pytest.mark.parametrize('agent_class,opt', [(dqn_agent.DqnAgent, 'adam'), (dqn_agent.DdqnAgent, 'rmsprop')])
def test_train_parity_conv_net(cuda, agent_class, opt):
  """5 train steps of a small Atari-style conv Q-net vs the numpy oracle: same seeds/inputs,
  loss within 1e-5 relative each step, parameters tracking to 1e-4."""
  rng = np.random.RandomState(0)
  A, B, n = 4, 32, 2
  obs_spec = tensor_spec.TensorSpec((20, 20, 2), torch.uint8)
  act_spec = tensor_spec.BoundedTensorSpec((), torch.int32, 0, A - 1)
  net = q_network.QNetwork(obs_spec, act_spec, preprocessing_layers=L.CastScale(255.),
                           conv_layer_params=((8, 4, 2), (16, 3, 1)), fc_layer_params=(32,),
                           device=cuda).set_seed(1)
  if opt == 'adam':
    optimizer, oopt = optimizers.AdamOptimizer(1e-3), ooptim.AdamTF(1e-3, eps=1e-8)
  else:
    optimizer = optimizers.RMSPropOptimizer(2.5e-4, decay=0.95, momentum=0.0, epsilon=1e-5, centered=True)
    oopt = ooptim.RMSPropTF(2.5e-4, decay=0.95, momentum=0.0, eps=1e-5, centered=True, ms_init=1.0)
  agent = agent_class(ts.time_step_spec(obs_spec), act_spec, q_network=net, optimizer=optimizer,
                      n_step_update=n, target_update_tau=0.5, target_update_period=2, gamma=0.99,
                      reward_scale_factor=1.0, gradient_clipping=10.0)
  agent.initialize()
  orc = odqn.DqnOracle(_mirror_oracle(net), oopt, gamma=0.99, n_step_update=n, loss_fn='huber',
                       target_update_tau=0.5, target_update_period=2, gradient_clipping=10.0,
                       ddqn=agent_class is dqn_agent.DdqnAgent)
  for step in range(5):
    e = _random_exp(rng, B, n + 1, (20, 20, 2), A, True)
    want = orc.train(e)
    got = agent.train(_to_traj(cuda, e))
    np.testing.assert_allclose(got.loss.item(), want['loss'], rtol=1e-5)
    np.testing.assert_allclose(got.extra.td_loss.cpu().numpy(), want['td_loss'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got.extra.td_error.cpu().numpy(), want['td_error'], rtol=1e-4, atol=1e-6)
  assert int(agent.train_step_counter.item()) == 5 == orc.train_step_counter
  for v, w in zip(net.variables, orc.q_net.params()):
    np.testing.assert_allclose(v.cpu().numpy(), w, rtol=1e-3, atol=2e-5)
  for v, w in zip(agent._target_q_network.variables, orc.target_net.params()):
    np.testing.assert_allclose(v.cpu().numpy(), w, rtol=1e-3, atol=2e-5)
  agent.check_numerics()


def test_cartpole_config_train_parity(cuda):
  """BASELINE config #1 shapes: obs f32[4], A=2, batch 64, T=2, Dense(100) Q-net, squared loss,
  Adam 1e-3, tau=0.05 every 5 steps (agents/dqn/examples/v2/train_eval.py:172-188)."""
  rng = np.random.RandomState(3)
  obs_spec = tensor_spec.TensorSpec((4,), torch.float32)
  act_spec = tensor_spec.BoundedTensorSpec((), torch.int64, 0, 1)
  net = sequential.Sequential([L.Dense(100, activation='relu'), L.Dense(2)], input_spec=obs_spec,
                              device=cuda).set_seed(2)
  agent = dqn_agent.DqnAgent(ts.time_step_spec(obs_spec), act_spec, q_network=net,
                             optimizer=optimizers.AdamOptimizer(1e-3),
                             td_errors_loss_fn=common.element_wise_squared_loss, gamma=0.99,
                             target_update_tau=0.05, target_update_period=5, epsilon_greedy=0.1)
  agent.initialize()
  orc = odqn.DqnOracle(_mirror_oracle(net), ooptim.AdamTF(1e-3, eps=1e-8), gamma=0.99,
                       loss_fn='squared', target_update_tau=0.05, target_update_period=5)
  for step in range(12):
    e = _random_exp(rng, 64, 2, (4,), 2, False)
    want = orc.train(e)
    tr = _to_traj(cuda, e)
    tr = tr._replace(action=tr.action.long())
    got = agent.train(tr)
    np.testing.assert_allclose(got.loss.item(), want['loss'], rtol=1e-5)
  for v, w in zip(agent._target_q_network.variables, orc.target_net.params()):
    np.testing.assert_allclose(v.cpu().numpy(), w, rtol=1e-4, atol=1e-6)


def test_graph_captured_train_step_matches_eager(cuda):
  """common.function (CUDA-graph capture) replays produce the same losses as eager calls."""
  rng = np.random.RandomState(4)
  obs_spec = tensor_spec.TensorSpec((4,), torch.float32)
  act_spec = tensor_spec.BoundedTensorSpec((), torch.int32, 0, 1)

  def make():
    net = sequential.Sequential([L.Dense(32, activation='relu'), L.Dense(2)], input_spec=obs_spec,
                                device=cuda).set_seed(9)
    a = dqn_agent.DqnAgent(ts.time_step_spec(obs_spec), act_spec, q_network=net,
                           optimizer=optimizers.AdamOptimizer(1e-2), gamma=0.9,
                           target_update_tau=1.0, target_update_period=3)
    a.initialize()
    return a

  eager, graphed = make(), make()
  train_fn = common.function(graphed.train)
  for step in range(8):
    e = _random_exp(rng, 16, 2, (4,), 2, False)
    l1 = eager.train(_to_traj(cuda, e)).loss.item()
    l2 = train_fn(_to_traj(cuda, e)).loss.item()
    np.testing.assert_allclose(l2, l1, rtol=1e-6)
  assert int(graphed.train_step_counter.item()) == 8 and graphed._train_step_host == 8
  assert torch.equal(eager._q_network.flat_params, graphed._q_network.flat_params)
  assert torch.equal(eager._target_q_network.flat_params, graphed._target_q_network.flat_params)
