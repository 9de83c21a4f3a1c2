"""Actor (tf_agents/train/actor.py:33-330): manages the interaction between a host policy and a
host environment through `PyDriver`, with metric observers.

Kept from the reference: constructor arguments; metrics are appended to the observers (:114-118);
only PyEnvironments are supported (:138-154); `run()` continues from the stored time step and
policy state (:175-187); `reset()` (:231-235); `collect_metrics` / `eval_metrics` helper lists
(:267-330).  Summaries: instead of tf.summary event files, `write_metric_summaries` appends one
JSON line per call to `<summary_dir>/metrics.jsonl` (TensorBoard plumbing is out of scope,
DESIGN.md §8).
"""
import json
import os

import numpy as np

from agents_b200.drivers import py_driver
from agents_b200.environments import py_environment
from agents_b200.environments import tf_environment
from agents_b200.metrics import py_metrics


class Actor(object):

  def __init__(self, env, policy, train_step, steps_per_run=None, episodes_per_run=None,
               observers=None, transition_observers=None, info_observers=None, metrics=None,
               reference_metrics=None, image_metrics=None, summary_dir=None,
               summary_interval=1000, end_episode_on_boundary=True, name=''):
    self._env = env
    self._policy = policy
    self._train_step = train_step
    self._metrics = list(metrics or [])
    self._image_metrics = list(image_metrics or [])
    self._reference_metrics = list(reference_metrics or [])
    obs = list(observers or [])
    for m in self._metrics + self._image_metrics:      # de-duplicated, order preserved
      if not any(m is o for o in obs):
        obs.append(m)
    self._observers = obs
    self._transition_observers = list(transition_observers or [])
    self._info_observers = list(info_observers or [])
    self._summary_dir = summary_dir
    self._write_summaries = bool(summary_dir)
    self._summary_interval = summary_interval
    self._last_summary = -summary_interval
    self._name = name
    if isinstance(env, py_environment.PyEnvironment):
      self._driver = py_driver.PyDriver(
          env, policy, self._observers, transition_observers=self._transition_observers,
          info_observers=self._info_observers, max_steps=steps_per_run,
          max_episodes=episodes_per_run, end_episode_on_boundary=end_episode_on_boundary)
    elif isinstance(env, tf_environment.TFEnvironment):
      raise ValueError("Actor doesn't support TFEnvironments yet.")
    else:
      raise ValueError('Unknown environment type.')
    self.reset()

  @property
  def metrics(self):
    return self._metrics

  @property
  def image_metrics(self):
    return self._image_metrics

  @property
  def train_step(self):
    return self._train_step

  @property
  def policy(self):
    return self._policy

  def _train_step_value(self):
    t = self._train_step
    return int(t.item()) if hasattr(t, 'item') else int(t)

  def run(self):
    self._time_step, self._policy_state = self._driver.run(self._time_step, self._policy_state)
    if (self._write_summaries and self._summary_interval > 0 and
        self._train_step_value() - self._last_summary >= self._summary_interval):
      self.write_metric_summaries()
      self._last_summary = self._train_step_value()

  def run_and_log(self):
    self.run()
    self.log_metrics()

  def write_metric_summaries(self):
    if not self._metrics or not self._write_summaries:
      return
    os.makedirs(self._summary_dir, exist_ok=True)
    rec = {'train_step': self._train_step_value(), 'actor': self._name}
    for m in self._metrics:
      rec[m.name] = float(np.asarray(m.result()))
    for ref in self._reference_metrics:
      rec['vs_' + ref.name] = float(np.asarray(ref.result()))
    with open(os.path.join(self._summary_dir, 'metrics.jsonl'), 'a') as f:
      f.write(json.dumps(rec) + '\n')

  def log_metrics(self):
    """Returns (and prints) 'name = value' for every metric (actor.py:216-229)."""
    line = ', '.join('{} = {}'.format(m.name, m.result()) for m in self._metrics)
    print('{} step = {}: {}'.format(self._name or 'Actor', self._train_step_value(), line))
    return line

  def reset(self):
    self._time_step = self._env.reset()
    self._policy_state = self._policy.get_initial_state(self._env.batch_size or 1)


def collect_metrics(buffer_size):
  """Standard collection metrics (actor.py:267-283)."""
  return [py_metrics.NumberOfEpisodes(), py_metrics.EnvironmentSteps(),
          py_metrics.AverageReturnMetric(buffer_size=buffer_size),
          py_metrics.AverageEpisodeLengthMetric(buffer_size=buffer_size)]


def eval_metrics(buffer_size):
  """Standard evaluation metrics (actor.py:286-300)."""
  return [py_metrics.AverageReturnMetric(buffer_size=buffer_size),
          py_metrics.AverageEpisodeLengthMetric(buffer_size=buffer_size)]
