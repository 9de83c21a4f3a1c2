"""PolicyStep record (tf_agents/trajectories/policy_step.py:31-77)."""
import collections


class PolicyStep(collections.namedtuple('PolicyStep', ['action', 'state', 'info'])):
  __slots__ = ()

  def replace(self, **kwargs):
    return self._replace(**kwargs)


PolicyStep.__new__.__defaults__ = ((), (), ())
