"""Host-side mirror of the reference interface: nests, specs, records, validation (no GPU)."""
import collections

import numpy as np
import pytest
import torch

from agents_b200.agents import tf_agent
from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import policy_step
from agents_b200.trajectories import time_step as ts
from agents_b200.trajectories import trajectory
from agents_b200.utils import nest


def test_nest_flatten_pack_roundtrip():
  Pair = collections.namedtuple('Pair', ['a', 'b'])
  s = Pair(a=(1, {'z': 2, 'y': 3}), b=[4, ()])
  flat = nest.flatten(s)
  assert flat == [1, 3, 2, 4]          # dict keys sorted, () has no leaves
  assert nest.pack_sequence_as(s, [10, 30, 20, 40]) == Pair(a=(10, {'z': 20, 'y': 30}), b=[40, ()])
  assert nest.map_structure(lambda x: x * 2, s).b == [8, ()]
  with pytest.raises(ValueError):
    nest.assert_same_structure((1, 2), (1, (2,)))


def test_specs():
  s = tensor_spec.TensorSpec([84, 84, 4], torch.uint8, 'obs')
  assert s.row_bytes == 28224 and s.shape == (84, 84, 4)
  b = tensor_spec.BoundedTensorSpec([], np.int32, 0, 5)
  assert b.dtype == torch.int32 and int(b.maximum) == 5
  assert s == tensor_spec.TensorSpec((84, 84, 4), np.uint8)


def test_time_step_constructors():
  obs = torch.zeros(3, 4)
  r = ts.restart(obs, batch_size=3)
  assert r.step_type.tolist() == [0, 0, 0] and r.discount.tolist() == [1, 1, 1]
  t = ts.transition(obs, torch.ones(3), discount=0.5)
  assert t.step_type.tolist() == [1, 1, 1] and t.discount.tolist() == [.5, .5, .5]
  e = ts.termination(obs, torch.ones(3))
  assert e.step_type.tolist() == [2, 2, 2] and e.discount.tolist() == [0, 0, 0]
  assert bool(e.is_last().all())
  spec = ts.time_step_spec(tensor_spec.TensorSpec([4], torch.float32))
  assert spec.step_type.dtype == torch.int32 and spec.discount.maximum == 1.0


def test_trajectory_from_transition_and_flags():
  obs = torch.arange(2.)
  t0 = ts.restart(obs, batch_size=2)
  t1 = ts.termination(obs + 1, torch.ones(2))
  tr = trajectory.from_transition(t0, policy_step.PolicyStep(torch.tensor([1, 2])), t1)
  assert tr.policy_info == () and tr.is_first().all() and tr.is_last().all()
  assert not tr.is_boundary().any()
  assert tr.reward.tolist() == [1, 1] and tr.discount.tolist() == [0, 0]


def test_to_transition_slices_time():
  B, T = 2, 3
  tr = trajectory.Trajectory(
      step_type=torch.zeros(B, T, dtype=torch.int32), observation=torch.arange(6.).reshape(B, T),
      action=torch.zeros(B, T), policy_info=(), next_step_type=torch.ones(B, T, dtype=torch.int32),
      reward=torch.ones(B, T), discount=torch.ones(B, T))
  time_steps, policy_steps, next_time_steps = trajectory.to_transition(tr)
  assert time_steps.observation.tolist() == [[0, 1], [3, 4]]
  assert next_time_steps.observation.tolist() == [[1, 2], [4, 5]]
  assert time_steps.reward.sum() == 0


def test_validate_trajectory_errors():
  spec = trajectory.Trajectory(
      step_type=tensor_spec.TensorSpec([], torch.int32), observation=tensor_spec.TensorSpec([2], torch.float32),
      action=tensor_spec.TensorSpec([], torch.int32), policy_info=(),
      next_step_type=tensor_spec.TensorSpec([], torch.int32),
      reward=tensor_spec.TensorSpec([], torch.float32), discount=tensor_spec.TensorSpec([], torch.float32))

  def make(B, T, obs_dim=2):
    return trajectory.Trajectory(
        step_type=torch.zeros(B, T, dtype=torch.int32), observation=torch.zeros(B, T, obs_dim),
        action=torch.zeros(B, T, dtype=torch.int32), policy_info=(),
        next_step_type=torch.zeros(B, T, dtype=torch.int32), reward=torch.zeros(B, T),
        discount=torch.zeros(B, T))

  tf_agent.validate_trajectory(make(4, 2), spec, 2)
  with pytest.raises(ValueError, match='sequence_length'):
    tf_agent.validate_trajectory(make(4, 3), spec, 2)
  with pytest.raises(ValueError, match='two outer dimensions'):
    tf_agent.validate_trajectory(make(4, 2, obs_dim=3), spec, 2)


def test_valid_range_ids_host():
  f = rb_mod._valid_range_ids
  assert f(-1, 10) == (0, 0)
  assert f(0, 10) == (0, 1)
  assert f(9, 10, 2) == (0, 9)
  assert f(10, 10, 2) == (1, 10)
  assert f(14, 10) == (5, 15)
  assert f(0, 10, 2) == (0, 0)


def test_replay_buffer_requires_cuda_device():
  with pytest.raises(ValueError, match='no CPU fallback'):
    rb_mod.TFUniformReplayBuffer(tensor_spec.TensorSpec([], torch.int64), batch_size=1,
                                 device='cpu')


def test_checkpointer_roundtrip_and_retention(tmp_path):
  """utils/common.py:1045-1100: latest checkpoint restored on construction, max_to_keep honoured."""
  import torch
  from agents_b200.utils import common

  class Box(object):
    def __init__(self, v):
      self.v = v
    def state_dict(self):
      return {'v': self.v}
    def load_state_dict(self, s):
      self.v = s['v']

  step, box = torch.tensor(0, dtype=torch.int64), Box(1)
  ck = common.Checkpointer(str(tmp_path), max_to_keep=2, global_step=step, box=box)
  assert not ck.checkpoint_exists and ck.initialize_or_restore() is False
  for s in (5, 10, 15):
    step.fill_(s)
    box.v = s * 2
    ck.save(step)
  import os
  assert sorted(os.listdir(str(tmp_path))) == ['ckpt-10.pt', 'ckpt-15.pt']
  step2, box2 = torch.tensor(0, dtype=torch.int64), Box(-1)
  ck2 = common.Checkpointer(str(tmp_path), max_to_keep=2, global_step=step2, box=box2)
  assert ck2.checkpoint_exists and ck2.initialize_or_restore() is True
  assert int(step2) == 15 and box2.v == 30


def test_batched_py_environment_contract():
  """environments/batched_py_environment.py:60-200 + py_environment.py:185-239."""
  import numpy as np
  from agents_b200.environments import batched_py_environment
  from py_env_mocks import PyEnvironmentMock
  envs = [PyEnvironmentMock(3), PyEnvironmentMock(4)]
  for threaded in (False, True):
    env = batched_py_environment.BatchedPyEnvironment([PyEnvironmentMock(3), PyEnvironmentMock(4)],
                                                      multithreading=threaded)
    assert env.batched and env.batch_size == 2
    t = env.step(np.array([1, 1], np.int32))            # no current step -> reset (:233-236)
    assert t.step_type.tolist() == [0, 0] and t.observation.tolist() == [0, 0]
    obs, types = [], []
    for a in (1, 2, 1, 1, 2):
      t = env.step(np.array([a, a], np.int32))
      obs.append(t.observation.tolist()); types.append(t.step_type.tolist())
    assert obs == [[1, 1], [3, 3], [0, 4], [1, 0], [3, 2]]
    assert types == [[1, 1], [2, 1], [0, 2], [1, 0], [2, 1]]
    assert t.reward.dtype == np.float32 and t.discount.tolist() == [0.0, 1.0]
    env.close()
  with pytest.raises(ValueError, match='already batched'):
    batched_py_environment.BatchedPyEnvironment(
        [batched_py_environment.BatchedPyEnvironment(envs, multithreading=False)])
  with pytest.raises(ValueError, match='must be a list or tuple'):
    batched_py_environment.BatchedPyEnvironment(envs[0])
