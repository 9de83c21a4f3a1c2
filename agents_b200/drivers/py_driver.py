"""PyDriver (tf_agents/drivers/py_driver.py:33-146): runs a host (numpy) policy in a host
PyEnvironment and feeds observers with numpy Trajectories.

Semantics kept from the reference: at least one of `max_steps` / `max_episodes` must be > 0
(:86-92); the loop stops when either budget is reached (:113); boundary transitions
(step_type == LAST) are passed to the observers but do not count as steps (:139);
`end_episode_on_boundary` selects whether an episode is counted at its boundary or at its LAST
transition (:134-137); for non-batched envs the policy state is re-initialised at the first step
of every episode after the first (:115-116); observers see the policy state that PRODUCED the
action (:122-127).

To store the numpy trajectories in the HBM ring use `PinnedAddBatch(replay_buffer)` below as the
observer: it stages each batched Trajectory through pinned memory and calls `add_batch` with
device tensors (the reference goes through `tf.numpy_function` in `TFPyEnvironment` or through
Reverb for this).
"""
import numpy as np
import torch

from agents_b200.drivers import driver
from agents_b200.trajectories import trajectory
from agents_b200.utils import nest


class PyDriver(driver.Driver):
  """A driver that runs a python policy in a python environment."""

  def __init__(self, env, policy, observers, transition_observers=None, info_observers=None,
               max_steps=None, max_episodes=None, end_episode_on_boundary=True):
    max_steps = max_steps or 0
    max_episodes = max_episodes or 0
    if max_steps < 1 and max_episodes < 1:
      raise ValueError('Either `max_steps` or `max_episodes` should be greater than 0.')
    super(PyDriver, self).__init__(env, policy, observers, transition_observers, info_observers)
    self._max_steps = max_steps or np.inf
    self._max_episodes = max_episodes or np.inf
    self._end_episode_on_boundary = end_episode_on_boundary

  @property
  def info_observers(self):
    return self._info_observers

  def run(self, time_step, policy_state=()):
    """Runs the policy from `time_step` / `policy_state`; returns the final pair."""
    num_steps = 0
    num_episodes = 0
    while num_steps < self._max_steps and num_episodes < self._max_episodes:
      if not self.env.batched and np.all(time_step.is_first()) and num_episodes > 0:
        policy_state = self._policy.get_initial_state(self.env.batch_size or 1)
      action_step = self.policy.action(time_step, policy_state)
      next_time_step = self.env.step(action_step.action)
      action_step_with_previous_state = action_step._replace(state=policy_state)
      traj = trajectory.from_transition(time_step, action_step_with_previous_state, next_time_step)
      for observer in self._transition_observers:
        observer((time_step, action_step_with_previous_state, next_time_step))
      for observer in self.observers:
        observer(traj)
      for observer in self.info_observers:
        observer(self.env.get_info())
      if self._end_episode_on_boundary:
        num_episodes += np.sum(traj.is_boundary())
      else:
        num_episodes += np.sum(traj.is_last())
      num_steps += np.sum(~np.asarray(traj.is_boundary()))
      time_step = next_time_step
      policy_state = action_step.state
    return time_step, policy_state


class PinnedAddBatch(object):
  """Observer: numpy batched Trajectory -> pinned staging -> async upload -> `rb.add_batch`.

  Two pinned/device buffer sets alternate; an event per set guards the reuse of its pinned
  memory, so the host may prepare step t+1 while step t is still being written into the ring."""

  def __init__(self, replay_buffer, device=None):
    self._rb = replay_buffer
    self._spec = replay_buffer.data_spec
    self._flat_specs = nest.flatten(self._spec)
    self._device = torch.device(device) if device is not None else replay_buffer.device
    b = replay_buffer.batch_size
    self._pinned, self._dev, self._done, self._used = [], [], [], [False, False]
    for _ in range(2):
      self._pinned.append([torch.empty((b,) + tuple(s.shape), dtype=s.dtype).pin_memory()
                           for s in self._flat_specs])
      self._dev.append([torch.empty((b,) + tuple(s.shape), dtype=s.dtype, device=self._device)
                        for s in self._flat_specs])
      self._done.append(torch.cuda.Event())
    self._slot = 0

  def __call__(self, traj):
    slot = self._slot
    self._slot ^= 1
    if self._used[slot]:
      self._done[slot].synchronize()
    flat = nest.flatten(traj)
    if len(flat) != len(self._flat_specs):
      raise ValueError('Trajectory does not match the replay buffer data_spec.')
    with torch.cuda.device(self._device):
      for pin, dev, leaf in zip(self._pinned[slot], self._dev[slot], flat):
        np.copyto(pin.numpy(), np.asarray(leaf).reshape(pin.shape), casting='same_kind')
        dev.copy_(pin, non_blocking=True)
      self._rb.add_batch(nest.pack_sequence_as(self._spec, list(self._dev[slot])))
      self._done[slot].record()
    self._used[slot] = True
