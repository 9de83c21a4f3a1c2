"""RandomPyPolicy (tf_agents/policies/random_py_policy.py:37-175): uniform host-side actions.

Used for the initial collect through `PyDriver` (e.g. `dqn/examples/v2/train_eval.py:205-214`
with a py environment) and by the Reverb observer tests.  Bounded integer specs are sampled over
the INCLUSIVE range, floats over [minimum, maximum) (`specs/array_spec.py:28-84`); an optional
action mask (`observation_and_action_constraint_splitter`) restricts scalar integer actions to
the allowed ones (`random_py_policy.py:140-165`).  Outer dimensions come from `outer_dims` or from
the observation's leading dimensions.
"""
import numpy as np

from agents_b200.policies import py_policy
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import policy_step
from agents_b200.utils import nest


def _sample(spec, rng, outer):
  dtype = tensor_spec.as_numpy_dtype(spec.dtype)
  shape = tuple(outer) + tuple(spec.shape)
  lo = getattr(spec, 'minimum', None)
  hi = getattr(spec, 'maximum', None)
  if np.issubdtype(dtype, np.floating):
    lo = np.finfo(dtype).min / 2 if lo is None else lo
    hi = np.finfo(dtype).max / 2 if hi is None else hi
    return rng.uniform(lo, hi, size=shape).astype(dtype)
  info = np.iinfo(dtype)
  lo = info.min if lo is None else np.asarray(lo, np.int64)
  hi = info.max if hi is None else np.asarray(hi, np.int64)
  return rng.integers(lo, np.asarray(hi, np.int64) + 1, size=shape, dtype=np.int64).astype(dtype)


class RandomPyPolicy(py_policy.PyPolicy):
  """Returns random samples of the given action_spec."""

  def __init__(self, time_step_spec, action_spec, policy_state_spec=(), info_spec=(), seed=None,
               outer_dims=None, observation_and_action_constraint_splitter=None):
    super().__init__(time_step_spec, action_spec, policy_state_spec, info_spec)
    if observation_and_action_constraint_splitter is not None:
      if nest.is_nested(action_spec) or not hasattr(action_spec, 'minimum'):
        raise NotImplementedError(
            'RandomPyPolicy only supports action constraints for BoundedArraySpec action specs.')
      if tuple(action_spec.shape) not in ((), (1,)):
        raise NotImplementedError(
            'RandomPyPolicy only supports action constraints for action specs shaped as () or '
            '(1,) or their equivalent list forms.')
    self._splitter = observation_and_action_constraint_splitter
    self._outer_dims = outer_dims
    self._rng = np.random.default_rng(seed)

  def _outer(self, time_step):
    if self._outer_dims is not None:
      return tuple(self._outer_dims)
    obs_specs = nest.flatten(self._time_step_spec.observation) if self._time_step_spec else []
    if not obs_specs:
      return ()
    first = np.asarray(nest.flatten(time_step.observation)[0])
    return first.shape[:first.ndim - len(obs_specs[0].shape)]

  def _action(self, time_step, policy_state):
    outer = self._outer(time_step)
    if self._splitter is not None:
      _, mask = self._splitter(time_step.observation)
      mask = np.asarray(mask).reshape(-1, np.asarray(mask).shape[-1]) > 0
      spec = self._action_spec
      lo = int(np.asarray(spec.minimum).reshape(-1)[0])
      picks = np.array([lo + self._rng.choice(np.flatnonzero(m)) for m in mask])
      action = picks.reshape(tuple(outer) + tuple(spec.shape)).astype(
          tensor_spec.as_numpy_dtype(spec.dtype))
    else:
      action = nest.map_structure(lambda s: _sample(s, self._rng, outer), self._action_spec)
    info = nest.map_structure(lambda s: _sample(s, self._rng, outer), self._info_spec)
    return policy_step.PolicyStep(action, policy_state, info)
