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

  def _transition(self, time_step, policy_state):
    """One interaction: (time_step, policy_state) -> (transition, next policy_state).  The
    PolicyStep handed to observers carries the state the policy was CALLED with, as the
    reference's driver does, so recurrent policies can be re-run from stored data."""
    step = self.policy.action(time_step, policy_state)
    next_time_step = self.env.step(step.action)
    return (time_step, step._replace(state=policy_state), next_time_step), step.state

  def _notify(self, transition):
    traj = trajectory.from_transition(*transition)
    for obs in self._transition_observers:
      obs(transition)
    for obs in self.observers:
      obs(traj)
    if self.info_observers:
      info = self.env.get_info()
      for obs in self.info_observers:
        obs(info)
    return traj

  def run(self, time_step, policy_state=()):
    """Steps the environment with the policy until `max_steps` non-boundary steps or
    `max_episodes` episode ends have been seen (whichever budget is finite); returns the final
    `(time_step, policy_state)` (reference drivers/py_driver.py:100-146)."""
    steps_left, episodes_left = self._max_steps, self._max_episodes
    episodes_seen = 0
    while steps_left > 0 and episodes_left > 0:
      # an unbatched environment that auto-reset gets a fresh policy state for the new episode
      if episodes_seen and not self.env.batched and np.all(time_step.is_first()):
        policy_state = self._policy.get_initial_state(self.env.batch_size or 1)
      transition, policy_state = self._transition(time_step, policy_state)
      traj = self._notify(transition)
      boundary = np.asarray(traj.is_boundary())
      ended = boundary if self._end_episode_on_boundary else np.asarray(traj.is_last())
      n_ended = int(np.sum(ended))
      episodes_seen += n_ended
      episodes_left -= n_ended
      steps_left -= int(boundary.size - np.sum(boundary))
      time_step = transition[2]
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
