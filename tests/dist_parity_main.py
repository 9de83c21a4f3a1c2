"""torchrun entry (2 ranks, NCCL): DQN, PPO and SAC trained data-parallel (local batch b, grads
SUM-all-reduced, losses divided by the global batch) must match one replica with batch 2b —
the multi-replica oracle of train/learner_test.py:442-540, on real GPUs."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from agents_b200 import optimizers  # noqa: E402
from agents_b200.agents.dqn import dqn_agent  # noqa: E402
from agents_b200.agents.ppo import ppo_clip_agent  # noqa: E402
from agents_b200.networks import actor_distribution_network, value_network  # noqa: E402
from agents_b200.networks import layers as L  # noqa: E402
from agents_b200.networks import q_network  # noqa: E402
from agents_b200.specs import tensor_spec  # noqa: E402
from agents_b200.trajectories import time_step as ts  # noqa: E402
from agents_b200.trajectories import trajectory  # noqa: E402
from agents_b200.train import learner as learner_lib  # noqa: E402
from agents_b200.train.utils import strategy_utils  # noqa: E402


def dqn_case(dev, strategy, rank, world):
  obs_spec = tensor_spec.TensorSpec((20, 20, 2), torch.uint8)
  act_spec = tensor_spec.BoundedTensorSpec((), torch.int32, 0, 3)

  def make():
    net = q_network.QNetwork(obs_spec, act_spec, preprocessing_layers=L.CastScale(255.),
                             conv_layer_params=((8, 4, 2),), fc_layer_params=(32,), device=dev).set_seed(3)
    a = dqn_agent.DqnAgent(ts.time_step_spec(obs_spec), act_spec, q_network=net,
                           optimizer=optimizers.AdamOptimizer(1e-3), gamma=0.99, target_update_period=2)
    a.initialize()
    return a

  g = torch.Generator().manual_seed(0)
  B = 32
  batches = []
  for _ in range(4):
    batches.append(trajectory.Trajectory(
        torch.randint(0, 3, (B, 2), generator=g, dtype=torch.int32), torch.randint(0, 256, (B, 2, 20, 20, 2), generator=g, dtype=torch.uint8),
        torch.randint(0, 4, (B, 2), generator=g, dtype=torch.int32), (), torch.randint(0, 3, (B, 2), generator=g, dtype=torch.int32),
        torch.rand(B, 2, generator=g), (torch.rand(B, 2, generator=g) > .1).float()))
  to = lambda tr, lo, hi: trajectory.Trajectory(*[(x[lo:hi].to(dev) if isinstance(x, torch.Tensor) else x) for x in tr])
  single = make()
  want = [single.train(to(b, 0, B)).loss.item() for b in batches]
  dp = make()
  lrn = learner_lib.Learner(tempfile.mkdtemp(), dp.train_step_counter, dp, strategy=strategy, checkpoint_interval=0)
  lo, hi = strategy.shard_range(B)
  got = [lrn.run(iterations=1, iterator=iter([(to(b, lo, hi), None)])).loss.item() for b in batches]
  np.testing.assert_allclose(got, want, rtol=2e-5)
  assert torch.allclose(single._q_network.flat_params, dp._q_network.flat_params, rtol=1e-4, atol=1e-6)
  assert torch.allclose(single._target_q_network.flat_params, dp._target_q_network.flat_params, rtol=1e-4, atol=1e-6)


def ppo_case(dev, strategy, rank, world):
  D, A, B, T = 17, 6, 16, 12
  obs_spec = tensor_spec.TensorSpec((D,), torch.float32)
  act_spec = tensor_spec.BoundedTensorSpec((A,), torch.float32, -1.0, 1.0)

  def make():
    actor = actor_distribution_network.ActorDistributionNetwork(obs_spec, act_spec, fc_layer_params=(32, 16),
                                                                activation_fn='tanh', device=dev).set_seed(1)
    value = value_network.ValueNetwork(obs_spec, fc_layer_params=(32, 16), activation_fn='tanh', device=dev).set_seed(2)
    a = ppo_clip_agent.PPOClipAgent(ts.time_step_spec(obs_spec), act_spec, optimizer=optimizers.Adam(1e-3),
                                    actor_net=actor, value_net=value, importance_ratio_clipping=0.2, use_gae=True,
                                    num_epochs=2, normalize_observations=False, normalize_rewards=False,
                                    gradient_clipping=0.5)
    a.initialize()
    return a

  g = torch.Generator().manual_seed(1)
  tr = trajectory.Trajectory(
      torch.ones(B, T, dtype=torch.int32), torch.randn(B, T, D, generator=g), torch.rand(B, T, A, generator=g) * 2 - 1,
      {'dist_params': {'loc': torch.randn(B, T, A, generator=g) * .1, 'scale': torch.full((B, T, A), .7)}},
      torch.ones(B, T, dtype=torch.int32), torch.rand(B, T, generator=g), torch.ones(B, T))
  from agents_b200.utils import nest
  to = lambda lo, hi: nest.map_structure(lambda x: x[lo:hi].to(dev), tr)
  single = make()
  want = single.train(to(0, B)).loss.item()
  dp = make()
  lrn = learner_lib.Learner(tempfile.mkdtemp(), dp.train_step_counter, dp, strategy=strategy, checkpoint_interval=0)
  lo, hi = strategy.shard_range(B)
  got = lrn.run(iterations=1, iterator=iter([(to(lo, hi), None)])).loss.item()
  np.testing.assert_allclose(got, want, rtol=5e-5)
  assert torch.allclose(single._flat_params, dp._flat_params, rtol=1e-3, atol=1e-5)


def main():
  rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
  torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
  dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
  dist.init_process_group('nccl', device_id=dev)
  strategy = strategy_utils.ProcessGroupStrategy()
  dqn_case(dev, strategy, rank, world)
  ppo_case(dev, strategy, rank, world)
  dist.barrier()
  if rank == 0:
    print('DIST_PARITY_OK', flush=True)
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
