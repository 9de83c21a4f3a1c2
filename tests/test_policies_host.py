"""Host-side wiring of the discrete policies (q_policy / greedy_policy / epsilon_greedy_policy /
random_tf_policy): the single b200rl_epsilon_greedy launch is replaced by a recording stub, so that
argument plumbing (epsilon conventions -1 / eps / 2, masks, dtypes, seeds, PolicyStep packing) is
checked without a GPU.  The launch itself is checked on the GPU in tests/test_optim_env_gpu.py and
tests/test_driver_gpu.py."""
import numpy as np
import pytest
import torch

from agents_b200 import _lib
from agents_b200.policies import epsilon_greedy_policy
from agents_b200.policies import greedy_policy
from agents_b200.policies import q_policy
from agents_b200.policies import random_tf_policy
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import time_step as ts


class _QNet(object):
  variables = ['w']

  def __call__(self, obs, step_type=None):
    return obs.float() * torch.arange(1, 4, dtype=torch.float32), ()


@pytest.fixture
def launches(monkeypatch):
  calls = []

  def call(name, *args):
    calls.append((name, args))
    return 0
  monkeypatch.setattr(_lib, 'call', call)
  monkeypatch.setattr(_lib, 'ptr', lambda t: t)
  monkeypatch.setattr(_lib, 'stream', lambda: 0)
  return calls


def _specs():
  obs = tensor_spec.TensorSpec((3,), torch.float32, 'observation')
  act = tensor_spec.BoundedTensorSpec((), torch.int64, 0, 2, 'action')
  return ts.time_step_spec(obs), act


def _step(b=4):
  return ts.TimeStep(torch.zeros(b, dtype=torch.int32), torch.zeros(b), torch.ones(b),
                     torch.ones(b, 3))


def test_module_layout_matches_the_reference():
  assert q_policy.GreedyPolicy is greedy_policy.GreedyPolicy
  assert q_policy.EpsilonGreedyPolicy is epsilon_greedy_policy.EpsilonGreedyPolicy
  assert q_policy.RandomTFPolicy is random_tf_policy.RandomTFPolicy
  assert issubclass(epsilon_greedy_policy.EpsilonGreedyPolicy, greedy_policy._Selecting)


def test_greedy_and_epsilon_greedy_pass_their_epsilon(launches):
  tss, act = _specs()
  policy = q_policy.QPolicy(tss, act, q_network=_QNet())
  assert policy.num_actions == 3 and policy.variables() == ['w']
  step = greedy_policy.GreedyPolicy(policy).action(_step())
  name, args = launches[-1]
  assert name == 'b200rl_epsilon_greedy' and args[2:5] == (4, 3, -1.0)   # (B, A, eps): never random
  assert step.action.dtype == torch.int64 and step.action.shape == (4,) and step.info == ()
  eps = [0.25]
  collect = epsilon_greedy_policy.EpsilonGreedyPolicy(policy, epsilon=lambda: eps[0], seed=7)
  collect.action(_step())
  assert launches[-1][1][4] == 0.25 and launches[-1][1][1] is None        # no mask
  eps[0] = 0.5                                                            # callable: read per call
  collect.action(_step())
  assert launches[-1][1][4] == 0.5
  assert launches[-1][1][5] != launches[0][1][5]                          # its own Philox key
  assert collect.wrapped_policy is policy and collect.variables() == ['w']
  policy.action(_step())                                                  # QPolicy itself acts greedily
  assert launches[-1][1][4] == -1.0


def test_masks_reach_the_launch_as_int32(launches):
  tss, act = _specs()
  split = lambda obs: (obs, torch.tensor([[1, 0, 1]] * obs.shape[0], dtype=torch.bool))
  policy = q_policy.QPolicy(tss, act, q_network=_QNet(),
                            observation_and_action_constraint_splitter=split)
  epsilon_greedy_policy.EpsilonGreedyPolicy(policy, epsilon=0.1).action(_step(2))
  mask = launches[-1][1][1]
  assert mask.dtype == torch.int32 and mask.tolist() == [[1, 0, 1], [1, 0, 1]]
  rnd = random_tf_policy.RandomTFPolicy(tss, act, seed=3,
                                        observation_and_action_constraint_splitter=split)
  out = rnd.action(_step(2))
  name, args = launches[-1]
  assert args[4] == 2.0 and args[0].shape == (2, 3) and float(args[0].abs().sum()) == 0.0
  assert args[1].dtype == torch.int32 and out.action.dtype == torch.int64


def test_only_scalar_actions():
  tss, act = _specs()
  with pytest.raises(ValueError, match='Only scalar actions'):
    q_policy.QPolicy(tss, (act, act), q_network=_QNet())
