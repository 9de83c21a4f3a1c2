"""PPO sharded-update benchmark (BASELINE.json configs[2]): synthetic MuJoCo-shape experience
(17-d obs, 6-d act), 4096 envs x T=128 in total, PPOClipAgent (GAE, clip 0.2, (200,100) tanh nets,
25 epochs, Adam).  Strong scaling: the 4096 trajectories are sharded over the ranks
(strategy.shard_range); every epoch all-reduces the 48 k-parameter gradient and the two
advantage-moment scalars.  Launch: python profiles/ppo_bench.py            (1 GPU)
        torchrun --nproc-per-node N profiles/ppo_bench.py                  (N GPUs)
Prints one JSON line on rank 0: train() ms (max over ranks, CUDA events) and env-steps/s."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import _lib, optimizers  # noqa: E402
from agents_b200.agents.ppo import ppo_clip_agent  # noqa: E402
from agents_b200.networks import actor_distribution_network, value_network  # noqa: E402
from agents_b200.specs import tensor_spec  # noqa: E402
from agents_b200.trajectories import time_step as ts  # noqa: E402
from agents_b200.trajectories import trajectory  # noqa: E402
from agents_b200.train import learner as learner_lib  # noqa: E402
from agents_b200.train.utils import strategy_utils  # noqa: E402
from agents_b200.utils import common  # noqa: E402


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  strategy = strategy_utils.get_strategy()
  B_total, T, D, A, epochs = 4096, 128, 17, 6, 25
  lo, hi = strategy.shard_range(B_total)
  B = hi - lo
  obs_spec = tensor_spec.TensorSpec((D,), torch.float32, 'observation')
  act_spec = tensor_spec.BoundedTensorSpec((A,), torch.float32, -1.0, 1.0, 'action')
  actor = actor_distribution_network.ActorDistributionNetwork(obs_spec, act_spec, fc_layer_params=(200, 100),
                                                              activation_fn='tanh', device=dev).set_seed(1)
  value = value_network.ValueNetwork(obs_spec, fc_layer_params=(200, 100), activation_fn='tanh', device=dev).set_seed(2)
  agent = ppo_clip_agent.PPOClipAgent(ts.time_step_spec(obs_spec), act_spec, optimizer=optimizers.Adam(3e-4),
                                      actor_net=actor, value_net=value, importance_ratio_clipping=0.2, use_gae=True,
                                      lambda_value=0.95, discount_factor=0.99, num_epochs=epochs,
                                      normalize_observations=False, normalize_rewards=False)
  agent.initialize()
  lrn = learner_lib.Learner('/tmp/ppo_bench', agent.train_step_counter, agent, strategy=strategy, checkpoint_interval=0)
  g = torch.Generator(device=dev).manual_seed(100 + strategy.rank)
  r = lambda *s: torch.rand(*s, device=dev, generator=g)
  exp = trajectory.Trajectory(
      torch.ones(B, T, dtype=torch.int32, device=dev), torch.randn(B, T, D, device=dev, generator=g), r(B, T, A) * 2 - 1,
      {'dist_params': {'loc': (r(B, T, A) - .5) * .2, 'scale': torch.full((B, T, A), .7, device=dev)}},
      torch.ones(B, T, dtype=torch.int32, device=dev), r(B, T), torch.ones(B, T, device=dev))
  use_graph = os.environ.get('PPO_BENCH_GRAPH', '1') == '1'
  fn = common.function(agent.train, warmup=1) if use_graph else agent.train
  fn(exp)                                  # eager warm-up
  ok = 1
  try:
    fn(exp)                                # capture + first replay
  except Exception as e:
    if not use_graph:
      raise
    sys.stderr.write(f'[rank {strategy.rank}] graph capture failed ({e}); eager\n')
    ok = 0
    torch.cuda.synchronize()
    agent.train(exp)                       # same number of collectives as a successful rank
  if world > 1 and use_graph:
    flag = torch.tensor([ok], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = int(flag.item())
  if use_graph and not ok:
    use_graph, fn = False, agent.train
  fn(exp)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  reps = 5
  c0 = _lib.launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    info = fn(exp)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / reps
  if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
  if strategy.rank == 0:
    print(json.dumps(dict(bench='ppo_sharded_update', n_gpus=world, envs_total=B_total, envs_per_gpu=B, T=T,
                          epochs=epochs, train_ms=ms, ms_per_epoch=ms / epochs,
                          samples_per_s=B_total * T * epochs / (ms * 1e-3), cuda_graph=use_graph,
                          loss=float(info.loss.item()), scaling='strong')), flush=True)
  if world > 1:
    # CUDA graphs that captured NCCL kernels make the communicator teardown hang on this stack:
    # leave without running destructors once every rank is done.
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == '__main__':
  main()
