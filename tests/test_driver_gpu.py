"""DynamicStepDriver collect -> store on the GPU: replays drivers/dynamic_step_driver_test.py
with the reference's mock env/policy restated on torch, then the device-resident envs."""
import numpy as np
import pytest
import torch

from agents_b200 import optimizers
from agents_b200.agents.dqn import dqn_agent
from agents_b200.drivers import dynamic_step_driver
from agents_b200.environments import random_tf_environment
from agents_b200.environments import tf_environment
from agents_b200.networks import layers as L
from agents_b200.networks import q_network
from agents_b200.networks import sequential
from agents_b200.policies import q_policy
from agents_b200.policies import tf_policy
from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import policy_step
from agents_b200.trajectories import time_step as ts
from oracle import env as oenv

pytestmark = pytest.mark.gpu


class EnvMock(tf_environment.TFEnvironment):
  """drivers/test_utils.py:43-93 PyEnvironmentMock (state += action, episode ends at state>=3),
  batched to [1] like TFPyEnvironment does."""

  def __init__(self, device, final_state=3):
    obs = tensor_spec.TensorSpec([], torch.int32, 'observation')
    act = tensor_spec.BoundedTensorSpec([], torch.int32, 1, 2, 'action')
    super().__init__(ts.time_step_spec(obs), act, batch_size=1)
    self._dev, self._final = device, final_state
    self._state = 0
    self._cur = None

  def _current_time_step(self):
    if self._cur is None:
      return self._reset()
    return self._cur

  def _reset(self):
    self._state = 0
    self._cur = ts.restart(torch.zeros(1, dtype=torch.int32, device=self._dev), batch_size=1)
    return self._cur

  def _step(self, action):
    if self._cur is None or self._state >= self._final:
      return self._reset()
    self._state += int(action.item())
    obs = torch.full((1,), self._state, dtype=torch.int32, device=self._dev)
    rew = torch.ones(1, device=self._dev)
    self._cur = ts.transition(obs, rew) if self._state < self._final else ts.termination(obs, rew)
    return self._cur


class PolicyMock(tf_policy.TFPolicy):
  """drivers/test_utils.py:96-161 TFPolicyMock: actions alternate 1,2; info = 2*action."""

  def __init__(self, time_step_spec, action_spec, device):
    super().__init__(time_step_spec, action_spec,
                     policy_state_spec=tensor_spec.BoundedTensorSpec((), torch.int32, 1, 2),
                     info_spec=action_spec)
    self._dev = device

  def get_initial_state(self, batch_size=None):
    return torch.zeros(batch_size or 1, dtype=torch.int32, device=self._dev)

  def _action(self, time_step, policy_state, seed):
    policy_state = torch.where(time_step.is_first(), torch.zeros_like(policy_state), policy_state)
    action = (policy_state % 2 + 1).to(torch.int32)
    return policy_step.PolicyStep(action, policy_state + 1, (action * 2).to(torch.int32))


def test_one_step_replay_buffer_observers_golden(cuda):
  # drivers/dynamic_step_driver_test.py:121-166 (and :168-199 with num_steps=6)
  for num_steps, runs in [(1, 6), (6, 1)]:
    env = EnvMock(cuda)
    policy = PolicyMock(env.time_step_spec(), env.action_spec(), cuda)
    rb = rb_mod.TFUniformReplayBuffer(policy.trajectory_spec, batch_size=1, max_length=1000, device=cuda)
    driver = dynamic_step_driver.DynamicStepDriver(env, policy, num_steps=num_steps,
                                                   observers=[rb.add_batch])
    time_step, policy_state = None, None
    for _ in range(runs):
      time_step, policy_state = driver.run(time_step, policy_state)
    tr = rb.gather_all()
    assert tr.step_type.cpu().tolist() == [[0, 1, 2, 0, 1, 2, 0, 1]]
    assert tr.observation.cpu().tolist() == [[0, 1, 3, 0, 1, 3, 0, 1]]
    assert tr.action.cpu().tolist() == [[1, 2, 1, 1, 2, 1, 1, 2]]
    assert tr.policy_info.cpu().tolist() == [[2, 4, 2, 2, 4, 2, 2, 4]]
    assert tr.next_step_type.cpu().tolist() == [[1, 2, 0, 1, 2, 0, 1, 2]]
    assert tr.reward.cpu().tolist() == [[1., 1., 0., 1., 1., 0., 1., 1.]]
    assert tr.discount.cpu().tolist() == [[1., 0., 1., 1., 0., 1., 1., 0.]]


def test_step_counts(cuda):  # :69-119: boundary steps are stored but not counted
  env = EnvMock(cuda)
  policy = PolicyMock(env.time_step_spec(), env.action_spec(), cuda)
  seen = []
  driver = dynamic_step_driver.DynamicStepDriver(env, policy, num_steps=5,
                                                 observers=[lambda t: seen.append(int(t.step_type.item()))])
  driver.run()
  assert sum(1 for s in seen if s != 2) == 5 and len(seen) == 5 + seen.count(2)
  seen.clear()
  driver.run(maximum_iterations=3)
  assert len(seen) == 3


def test_random_env_driver_with_dqn_collect_policy(cuda):
  """Atari-shape collect: EpsilonGreedy(QPolicy) -> RandomTFEnvironment -> add_batch; contents
  are checked against the oracle env stream and the trajectory alignment rule
  (row t = obs_t, action_t, reward_{t+1}; trajectories/trajectory.py:50-72)."""
  B_env, steps = 8, 12
  obs_spec = tensor_spec.TensorSpec((84, 84, 4), torch.uint8, 'observation')
  act_spec = tensor_spec.BoundedTensorSpec((), torch.int32, 0, 5, 'action')
  tss = ts.time_step_spec(obs_spec)
  env = random_tf_environment.RandomTFEnvironment(tss, act_spec, batch_size=B_env,
                                                  episode_end_probability=0.25, seed=3, device=cuda)
  net = q_network.QNetwork(obs_spec, act_spec, preprocessing_layers=L.CastScale(255.),
                           conv_layer_params=((8, 8, 4),), fc_layer_params=(16,), device=cuda).set_seed(0)
  agent = dqn_agent.DqnAgent(tss, act_spec, q_network=net, optimizer=optimizers.AdamOptimizer(1e-3),
                             epsilon_greedy=0.5)
  agent.initialize()
  rb = rb_mod.TFUniformReplayBuffer(agent.collect_data_spec, batch_size=B_env, max_length=64, device=cuda)
  driver = dynamic_step_driver.DynamicStepDriver(env, agent.collect_policy, observers=[rb.add_batch],
                                                 num_steps=B_env)
  for _ in range(steps):
    driver.run(maximum_iterations=1)
  tr = rb.gather_all()
  assert tuple(tr.observation.shape) == (B_env, steps, 84, 84, 4)
  # oracle env stream: call 0 is the reset, then one call per step
  seed = (3 ^ random_tf_environment._ENV_SEED_TAG) & 0xFFFFFFFFFFFFFFFF
  st = np.full(B_env, 2, np.int32)
  ts_list = []
  for call in range(steps + 1):
    st, obs, rew, disc = oenv.random_env_step(st, 84 * 84 * 4, True, 0.25, seed, call)
    ts_list.append((st.copy(), obs.reshape(B_env, 84, 84, 4), rew, disc))
  for t in range(steps):
    np.testing.assert_array_equal(tr.step_type[:, t].cpu().numpy(), ts_list[t][0])
    np.testing.assert_array_equal(tr.observation[:, t].cpu().numpy(), ts_list[t][1])
    np.testing.assert_array_equal(tr.next_step_type[:, t].cpu().numpy(), ts_list[t + 1][0])
    np.testing.assert_array_equal(tr.reward[:, t].cpu().numpy(), ts_list[t + 1][2])
    np.testing.assert_array_equal(tr.discount[:, t].cpu().numpy(), ts_list[t + 1][3])
  a = tr.action.cpu().numpy()
  assert a.min() >= 0 and a.max() <= 5 and len(np.unique(a)) > 1


def test_cartpole_collect_and_train_loop(cuda):
  """BASELINE config #1 end to end (agents/dqn/examples/v2/train_eval.py:151-297): CartPole,
  buffer 10k, batch 64, 1 env, collect 1 step / train 1 step; loss stays finite and > 0
  (the reference's own smoke assertion, train_eval_test.py:30-42)."""
  env = random_tf_environment.CartPoleTFEnvironment(batch_size=1, seed=1, device=cuda)
  tss, act_spec = env.time_step_spec(), env.action_spec()
  net = sequential.Sequential([L.Dense(100, activation='relu'), L.Dense(2)],
                              input_spec=tss.observation, device=cuda).set_seed(0)
  from agents_b200.utils import common
  agent = dqn_agent.DqnAgent(tss, act_spec, q_network=net, optimizer=optimizers.AdamOptimizer(1e-3),
                             td_errors_loss_fn=common.element_wise_squared_loss, gamma=0.99,
                             target_update_tau=0.05, target_update_period=5, epsilon_greedy=0.1)
  agent.initialize()
  rb = rb_mod.TFUniformReplayBuffer(agent.collect_data_spec, batch_size=1, max_length=10000, device=cuda)
  random_policy = q_policy.RandomTFPolicy(tss, act_spec)
  dynamic_step_driver.DynamicStepDriver(env, random_policy, observers=[rb.add_batch], num_steps=200).run()
  assert int(rb.num_frames()) >= 200
  driver = dynamic_step_driver.DynamicStepDriver(env, agent.collect_policy, observers=[rb.add_batch], num_steps=1)
  ds = iter(rb.as_dataset(sample_batch_size=64, num_steps=2).prefetch(3))
  losses = []
  for _ in range(30):
    driver.run()
    exp, _ = next(ds)
    losses.append(agent.train(exp).loss.item())
  assert all(np.isfinite(losses)) and losses[-1] > 0
  assert int(agent.train_step_counter.item()) == 30
  frames = rb.gather_all()
  assert frames.observation.shape[1] == int(rb.num_frames())
  assert float(frames.observation.abs().max()) < 5.0


def test_host_env_bridge_replays_driver_golden(cuda):
  """drivers/dynamic_step_driver_test.py:121-199 through the real bridge: numpy PyEnvironmentMock
  -> TFPyEnvironment (pinned staging, async upload) -> DynamicStepDriver -> ring."""
  from agents_b200.environments import tf_py_environment
  from py_env_mocks import PyEnvironmentMock
  for num_steps, runs, isolation in [(1, 6, False), (6, 1, True)]:
    py_env = PyEnvironmentMock()
    env = tf_py_environment.TFPyEnvironment(py_env, check_dims=True, isolation=isolation, device=cuda)
    assert env.batch_size == 1 and env.pyenv.envs[0] is py_env
    policy = PolicyMock(env.time_step_spec(), env.action_spec(), cuda)
    rb = rb_mod.TFUniformReplayBuffer(policy.trajectory_spec, batch_size=1, max_length=1000, device=cuda)
    driver = dynamic_step_driver.DynamicStepDriver(env, policy, num_steps=num_steps,
                                                   observers=[rb.add_batch])
    time_step, policy_state = None, None
    for _ in range(runs):
      time_step, policy_state = driver.run(time_step, policy_state)
    tr = rb.gather_all()
    assert tr.step_type.cpu().tolist() == [[0, 1, 2, 0, 1, 2, 0, 1]]
    assert tr.observation.cpu().tolist() == [[0, 1, 3, 0, 1, 3, 0, 1]]
    assert tr.action.cpu().tolist() == [[1, 2, 1, 1, 2, 1, 1, 2]]
    assert tr.next_step_type.cpu().tolist() == [[1, 2, 0, 1, 2, 0, 1, 2]]
    assert tr.reward.cpu().tolist() == [[1., 1., 0., 1., 1., 0., 1., 1.]]
    assert tr.discount.cpu().tolist() == [[1., 0., 1., 1., 0., 1., 1., 0.]]
    assert py_env.actions_taken == [1, 2, 1, 2, 1, 2]        # the step after LAST ignores its action
    env.close()


def test_host_env_bridge_batched_and_check_dims(cuda):
  from agents_b200.environments import batched_py_environment
  from agents_b200.environments import tf_py_environment
  from py_env_mocks import PyEnvironmentMock
  py_env = batched_py_environment.BatchedPyEnvironment([PyEnvironmentMock(3), PyEnvironmentMock(4)])
  env = tf_py_environment.TFPyEnvironment(py_env, check_dims=True, device=cuda)
  t0 = env.reset()
  assert t0.step_type.device.type == 'cuda' and t0.step_type.tolist() == [0, 0]
  seen = [t0.observation.clone()]
  for a in (1, 2, 1, 1, 2):
    t = env.step(torch.full((2,), a, dtype=torch.int32, device=cuda))
    seen.append(t.observation.clone())                  # staging sets alternate: clone what we keep
  assert torch.stack(seen).cpu().tolist() == [[0, 0], [1, 1], [3, 3], [0, 4], [1, 0], [3, 2]]
  assert env.current_time_step().step_type.tolist() == [2, 1]
  with pytest.raises(ValueError, match='major dimension is batch_size'):
    env.step(torch.ones(3, dtype=torch.int32, device=cuda))
  env.close()
