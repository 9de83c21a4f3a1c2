"""PyPolicy (tf_agents/policies/py_policy.py:33-210): host-side (numpy) policy contract used by
`PyDriver`.  `PyTFEagerPolicy` (py_tf_eager_policy.py) runs a device policy of this package behind
that contract; `RandomPyPolicy` (random_py_policy.py) samples the action spec on the host."""
import abc


class PyPolicy(abc.ABC):

  def __init__(self, time_step_spec, action_spec, policy_state_spec=(), info_spec=()):
    self._time_step_spec = time_step_spec
    self._action_spec = action_spec
    self._policy_state_spec = policy_state_spec
    self._info_spec = info_spec

  @property
  def time_step_spec(self):
    return self._time_step_spec

  @property
  def action_spec(self):
    return self._action_spec

  @property
  def policy_state_spec(self):
    return self._policy_state_spec

  @property
  def info_spec(self):
    return self._info_spec

  def get_initial_state(self, batch_size=None):
    return self._get_initial_state(batch_size)

  def action(self, time_step, policy_state=(), seed=None):
    return self._action(time_step, policy_state)

  def _get_initial_state(self, batch_size=None):
    return ()

  @abc.abstractmethod
  def _action(self, time_step, policy_state):
    pass


def __getattr__(name):
  # `py_policy.PyTFEagerPolicy` was this module's spelling before the class moved to its
  # reference-named module; resolved lazily (py_tf_eager_policy imports PyPolicy from here)
  if name == 'PyTFEagerPolicy':
    from agents_b200.policies import py_tf_eager_policy
    return py_tf_eager_policy.PyTFEagerPolicy
  raise AttributeError(name)
