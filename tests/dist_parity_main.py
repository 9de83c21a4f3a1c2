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
                                    num_epochs=2, normalize_observations=True, normalize_rewards=True,
                                    gradient_clipping=0.5)
    a.initialize()
    return a

  g = torch.Generator().manual_seed(1)
  from agents_b200.utils import nest
  single = make()
  dp = make()
  lrn = learner_lib.Learner(tempfile.mkdtemp(), dp.train_step_counter, dp, strategy=strategy, checkpoint_interval=0)
  lo, hi = strategy.shard_range(B)
  for call in range(3):          # the 2nd and 3rd calls normalise with the merged statistics of the 1st
    tr = trajectory.Trajectory(
        torch.ones(B, T, dtype=torch.int32), torch.randn(B, T, D, generator=g) * (1 + call) + call,
        torch.rand(B, T, A, generator=g) * 2 - 1,
        {'dist_params': {'loc': torch.randn(B, T, A, generator=g) * .1, 'scale': torch.full((B, T, A), .7)}},
        torch.ones(B, T, dtype=torch.int32), torch.rand(B, T, generator=g) * (1 + call), torch.ones(B, T))
    to = lambda a, b: nest.map_structure(lambda x: x[a:b].to(dev), tr)
    want = single.train(to(0, B)).loss.item()
    got = lrn.run(iterations=1, iterator=iter([(to(lo, hi), None)])).loss.item()
    np.testing.assert_allclose(got, want, rtol=5e-5)
  assert torch.allclose(single._flat_params, dp._flat_params, rtol=1e-3, atol=1e-5)
  for a, b in zip(single._observation_normalizer.variables[:3], dp._observation_normalizer.variables[:3]):
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)          # count / mean / m2 of the GLOBAL batches
  for a, b in zip(single._reward_normalizer.variables[:3], dp._reward_normalizer.variables[:3]):
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
  RESULTS['ppo_param_rel'] = float((single._flat_params - dp._flat_params).abs().max() /
                                   single._flat_params.abs().max())


def sampler_case(dev, strategy, rank, world):
  """Parity-mode sampling: the ranks' slices of get_next_global, concatenated, are
  bit-identical to what ONE buffer holding all segments returns for the same seed."""
  from agents_b200.replay_buffers import sharded_replay_buffer as srb
  from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
  from agents_b200.utils import nest
  spec = trajectory.Trajectory(
      tensor_spec.TensorSpec([], torch.int32, 'step_type'), tensor_spec.TensorSpec((12, 12, 4), torch.uint8, 'observation'),
      tensor_spec.TensorSpec([], torch.int32, 'action'), (), tensor_spec.TensorSpec([], torch.int32, 'next_step_type'),
      tensor_spec.TensorSpec([], torch.float32, 'reward'), tensor_spec.TensorSpec([], torch.float32, 'discount'))
  B_env, Lm, Bs, T = 8, 16, 32, 2
  full = rb_mod.TFUniformReplayBuffer(spec, batch_size=B_env, max_length=Lm, device=dev, seed=77)
  shard = srb.ShardedUniformReplayBuffer(spec, batch_size=B_env, max_length=Lm, strategy=strategy, device=dev, seed=77)
  lo, hi = shard.segment_range
  g = torch.Generator().manual_seed(5)
  for t in range(23):                                          # wraps the ring
    items = trajectory.Trajectory(
        torch.randint(0, 3, (B_env,), generator=g, dtype=torch.int32),
        torch.randint(0, 256, (B_env, 12, 12, 4), generator=g, dtype=torch.uint8),
        torch.arange(B_env, dtype=torch.int32), (), torch.randint(0, 3, (B_env,), generator=g, dtype=torch.int32),
        torch.full((B_env,), float(t)), torch.ones(B_env))
    full.add_batch(nest.map_structure(lambda x: x.to(dev), items))
    shard.add_batch(nest.map_structure(lambda x: x[lo:hi].to(dev), items))
  for _ in range(3):
    want, winfo = full.get_next(sample_batch_size=Bs, num_steps=T)
    got, ginfo = shard.get_next_global(Bs, T)
    a, b = rank * (Bs // world), (rank + 1) * (Bs // world)
    for x, y in zip(nest.flatten(want), nest.flatten(got)):
      assert torch.equal(x[a:b], y), 'sharded parity-mode batch differs from the single-buffer batch'
    assert torch.equal(winfo.ids[a:b], ginfo.ids)
    assert torch.allclose(winfo.probabilities[a:b], ginfo.probabilities)
  RESULTS['sampler_bit_exact'] = 1.0


RESULTS = {}


def main():
  rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
  torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
  dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
  strategy_utils.configure_nccl_env()
  dist.init_process_group('nccl', device_id=dev)
  strategy = strategy_utils.ProcessGroupStrategy()
  dqn_case(dev, strategy, rank, world)
  ppo_case(dev, strategy, rank, world)
  sampler_case(dev, strategy, rank, world)
  dist.barrier()
  if rank == 0:
    import json
    RESULTS['replicas'] = world
    print('DIST_PARITY_OK ' + json.dumps(RESULTS), flush=True)
    out = os.path.join(ROOT, 'gpurun_out')
    try:
      os.makedirs(out, exist_ok=True)
      json.dump(RESULTS, open(os.path.join(out, 'dist_parity.json'), 'w'))
    except OSError:
      pass
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
