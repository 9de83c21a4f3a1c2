"""N>1 host logic on CPU: world_size-2 gloo process groups (no GPU).

Restates the idea of train/learner_test.py:442-540 (testLossLearnerDifferentDistStrat): the
same weights trained under 1 replica x batch 8 and 2 replicas x batch 4 must end up with equal
losses and equal variables, because per-example losses are divided by the GLOBAL batch and the
gradient is SUM-all-reduced.  The agent here is a CPU fake (like the reference's FakePPOAgent,
train/ppo_learner_test_utils.py:31-75) that uses the same hooks (replicas, _grad_sync) the CUDA
agents use.
"""
import os
import socket
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agents_b200.agents import tf_agent
from agents_b200.train import learner as learner_lib
from agents_b200.train.utils import strategy_utils


class FakeAgent(object):
  """Linear regression 'agent': loss = sum((x@w - y)^2) / global_batch, SGD."""

  def __init__(self):
    self._w = torch.tensor([0.5, -1.0, 2.0])
    self._q_network = type('N', (), {'flat_params': self._w})()
    self.train_step_counter = torch.zeros((), dtype=torch.int64)
    self._train_step_host = 0
    self.replicas = 1
    self._grad_sync = None

  def train(self, experience):
    x, y = experience
    gb = x.shape[0] * self.replicas
    err = x @ self._w - y
    loss = (err ** 2).sum() / gb
    grad = 2 * (x.t() @ err) / gb
    if self._grad_sync is not None:
      self._grad_sync(grad)
    self._w -= 0.1 * grad
    self.train_step_counter += 1
    self._train_step_host += 1
    return tf_agent.LossInfo(loss, ())

  def loss(self, experience):
    x, y = experience
    return tf_agent.LossInfo(((x @ self._w - y) ** 2).sum() / (x.shape[0] * self.replicas), ())


def _data(steps=5):
  g = torch.Generator().manual_seed(0)
  return [(torch.randn(8, 3, generator=g), torch.randn(8, generator=g)) for _ in range(steps)]


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  strategy = strategy_utils.ProcessGroupStrategy()
  lo, hi = strategy.shard_range(8)
  agent = FakeAgent()
  if rank == 1:
    agent._w += 5.0                      # must be overwritten by the rank-0 broadcast
  data = [((x[lo:hi], y[lo:hi]), None) for x, y in _data()]
  lrn = learner_lib.Learner(out_dir, agent.train_step_counter, agent, strategy=strategy, checkpoint_interval=0)
  losses = [float(lrn.run(iterations=1, iterator=iter([d])).loss) for d in data]
  torch.save({'w': agent._w, 'losses': losses, 'range': (lo, hi)}, os.path.join(out_dir, f'r{rank}.pt'))
  dist.destroy_process_group()


def test_two_replicas_match_single_replica():
  single = FakeAgent()
  lrn = learner_lib.Learner(tempfile.mkdtemp(), single.train_step_counter, single,
                            strategy=strategy_utils.SingleProcessStrategy(), checkpoint_interval=0)
  want = [float(lrn.run(iterations=1, iterator=iter([(d, None)])).loss) for d in _data()]
  out = tempfile.mkdtemp()
  mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  r0, r1 = torch.load(os.path.join(out, 'r0.pt')), torch.load(os.path.join(out, 'r1.pt'))
  assert r0['range'] == (0, 4) and r1['range'] == (4, 8)
  np.testing.assert_allclose(r0['losses'], want, rtol=1e-5)         # SUM-reduced LossInfo
  np.testing.assert_allclose(r1['losses'], want, rtol=1e-5)
  assert torch.allclose(r0['w'], single._w, rtol=1e-5, atol=1e-6)   # mirrored variables
  assert torch.equal(r0['w'], r1['w'])


def test_checkpoint_restore_roundtrip():
  root = tempfile.mkdtemp()
  a = FakeAgent()
  lrn = learner_lib.Learner(root, a.train_step_counter, a, strategy=strategy_utils.SingleProcessStrategy(),
                            checkpoint_interval=2)
  lrn.run(iterations=3, iterator=iter([(d, None) for d in _data()]))
  assert os.listdir(os.path.join(root, 'train', 'checkpoints'))
  b = FakeAgent()
  learner_lib.Learner(root, b.train_step_counter, b, strategy=strategy_utils.SingleProcessStrategy(),
                      checkpoint_interval=2)
  assert torch.equal(a._w, b._w) and int(b.train_step_counter) == 3 == b._train_step_host


def test_shard_range_errors_and_default_strategy():
  s = strategy_utils.SingleProcessStrategy()
  assert s.shard_range(7) == (0, 7) and s.num_replicas_in_sync == 1
  assert isinstance(strategy_utils.get_strategy(), strategy_utils.SingleProcessStrategy)


def _ppo_worker(rank, world, port, out_dir):
  """PPOLearner under 2 gloo replicas, full-sequence mode (host only): the number of train
  calls per replica is num_samples * num_epochs / num_replicas (train/ppo_learner.py:283-300)
  and the LossInfo returned by run() is SUM-reduced."""
  import collections
  from agents_b200.train import ppo_learner
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  strategy = strategy_utils.ProcessGroupStrategy()
  agent = FakeAgent()
  agent._compute_value_and_advantage_in_train = False
  agent.update_normalizers_in_train = False
  agent.update_observation_normalizer = lambda obs: None
  agent.update_reward_normalizer = lambda r: None
  calls = []
  agent.train = lambda exp: (calls.append(1), agent.__class__.loss(agent, exp))[1]
  Traj = collections.namedtuple('Traj', ['observation', 'reward', 'xy'])
  g = torch.Generator().manual_seed(rank)
  x, y = torch.randn(4, 3, generator=g), torch.randn(4, generator=g)

  class Exp(Traj):                       # what FakeAgent.loss unpacks: (x, y)
    def __iter__(self):
      return iter(self.xy)
  sample = Exp(observation=x, reward=torch.zeros(4, 6), xy=(x, y))
  ds = lambda: [sample] * 3
  lrn = ppo_learner.PPOLearner(out_dir, agent.train_step_counter, agent, ds, ds, num_samples=3, num_epochs=4,
                               strategy=strategy, checkpoint_interval=0)
  info = lrn.run()
  local = float(agent.__class__.loss(agent, sample).loss)
  torch.save({'calls': len(calls), 'frames': lrn.num_frames_for_training, 'loss': float(info.loss),
              'local': local}, os.path.join(out_dir, f'p{rank}.pt'))
  dist.destroy_process_group()


def test_ppo_learner_iterations_under_two_replicas():
  out = tempfile.mkdtemp()
  mp.spawn(_ppo_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  r0, r1 = torch.load(os.path.join(out, 'p0.pt')), torch.load(os.path.join(out, 'p1.pt'))
  assert r0['calls'] == r1['calls'] == 3 * 4 // 2            # 12 batches over 2 replicas
  assert r0['frames'] == r1['frames'] == 3 * 4 * 6
  np.testing.assert_allclose(r0['loss'], r0['local'] + r1['local'], rtol=1e-5)   # SUM over replicas
  np.testing.assert_allclose(r1['loss'], r0['loss'], rtol=1e-6)
