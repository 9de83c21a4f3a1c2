"""DynamicStepDriver (tf_agents/drivers/dynamic_step_driver.py:48-224).

Steps the environment with the policy until `sum(counter) >= num_steps`, calling every observer
with the Trajectory of each step; `counter` counts non-boundary steps per batch entry (:170).

The reference runs a `tf.while_loop(parallel_iterations=1)` whose condition is evaluated on
every iteration.  Here each iteration only ENQUEUES device work (policy launch, env-step launch,
observer launches), and the loop condition is resolved with as few host reads as possible:
an iteration adds at most `batch_size` to the counter, so `ceil(remaining / batch_size)`
iterations can always be enqueued before the counter has to be looked at — the sequence of
iterations is exactly the reference's.  With `maximum_iterations` no read is needed at all once
that bound is reached, which is what makes a collect step CUDA-graph capturable.
"""
import torch

from agents_b200.drivers import driver
from agents_b200.environments import tf_environment
from agents_b200.policies import tf_policy
from agents_b200.trajectories import time_step as ts
from agents_b200.trajectories import trajectory


class DynamicStepDriver(driver.Driver):
  """A driver that takes N steps in an environment using a policy."""

  def __init__(self, env, policy, observers=None, transition_observers=None, num_steps=1):
    if not isinstance(env, tf_environment.TFEnvironment):
      raise ValueError('`env` must be an instance of tf_environment.TFEnvironment.')
    if not isinstance(policy, tf_policy.TFPolicy):
      raise ValueError('`policy` must be an instance of tf_policy.TFPolicy.')
    super(DynamicStepDriver, self).__init__(env, policy, observers, transition_observers)
    self._num_steps = num_steps

  def _loop_body(self, counter, time_step, policy_state):
    """One iteration of the reference loop body (:124-172)."""
    action_step = self.policy.action(time_step, policy_state)
    policy_state = action_step.state
    next_time_step = self.env.step(action_step.action)
    traj = trajectory.from_transition(time_step, action_step, next_time_step)
    for observer in self._observers:
      observer(traj)
    for observer in self._transition_observers:
      observer((time_step, action_step, next_time_step))
    counter += (traj.step_type != ts.StepType.LAST).sum()   # += ~is_boundary (:170)
    return next_time_step, policy_state

  def run(self, time_step=None, policy_state=None, maximum_iterations=None):
    """Returns (time_step, policy_state) after the loop (:176-224)."""
    if time_step is None:
      time_step = self.env.current_time_step()
    if policy_state is None:
      policy_state = self.policy.get_initial_state(self.env.batch_size)
    batch = max(int(self.env.batch_size or 1), 1)
    counter = torch.zeros((), dtype=torch.int64, device=time_step.step_type.device)
    counted, done_iters = 0, 0
    while counted < self._num_steps:
      k = -(-(self._num_steps - counted) // batch)
      if maximum_iterations is not None:
        k = min(k, int(maximum_iterations) - done_iters)
      if k <= 0:
        break
      for _ in range(k):
        time_step, policy_state = self._loop_body(counter, time_step, policy_state)
      done_iters += k
      if maximum_iterations is not None and done_iters >= int(maximum_iterations):
        break
      counted = int(counter.item())       # the only host read; skipped under maximum_iterations
    return time_step, policy_state
