"""SAC sharded-update benchmark (BASELINE.json configs[3] / SURVEY §8d config 4): obs 17, act 6,
1 M-slot ring of 108-byte rows, global batch 1024 sharded over the ranks (128/GPU at 8 GPUs),
nets (256, 256) relu, tau 0.005, reward_scale 0.1, Adam 3e-4.  One step = get_next(B/N, 2) +
SacAgent.train (critic -> actor -> alpha, three optimiser applies, Polyak), captured in one CUDA
graph; gradients are all-reduced per optimiser by the Learner hooks.
Launch: python profiles/sac_bench.py   |   torchrun --nproc-per-node N profiles/sac_bench.py
NOT RUN in round 1 (written after the GPU budget was spent) - first thing to measure in round 2."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import optimizers  # noqa: E402
from agents_b200.agents.sac import sac_agent  # noqa: E402
from agents_b200.networks import critic_network  # noqa: E402
from agents_b200.networks import tanh_normal_projection_network as tnp  # noqa: E402
from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod  # noqa: E402
from agents_b200.specs import tensor_spec  # noqa: E402
from agents_b200.train import learner as learner_lib  # noqa: E402
from agents_b200.train.utils import strategy_utils  # noqa: E402
from agents_b200.trajectories import time_step as ts  # noqa: E402
from agents_b200.trajectories import trajectory  # noqa: E402
from agents_b200.utils import common  # noqa: E402


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  strategy = strategy_utils.get_strategy()
  D, A, B_global, B_env, L = 17, 6, 1024, 256, 4096
  B = B_global // world
  obs_spec = tensor_spec.TensorSpec((D,), torch.float32, 'observation')
  act_spec = tensor_spec.BoundedTensorSpec((A,), torch.float32, -1.0, 1.0, 'action')
  actor = tnp.TanhNormalActorNetwork(obs_spec, act_spec, fc_layer_params=(256, 256), device=dev).set_seed(1)
  critic = critic_network.CriticNetwork((obs_spec, act_spec), joint_fc_layer_params=(256, 256),
                                        device=dev).set_seed(2)
  agent = sac_agent.SacAgent(ts.time_step_spec(obs_spec), act_spec, critic_network=critic, actor_network=actor,
                             actor_optimizer=optimizers.Adam(3e-4), critic_optimizer=optimizers.Adam(3e-4),
                             alpha_optimizer=optimizers.Adam(3e-4), target_update_tau=0.005,
                             target_update_period=1, gamma=0.99, reward_scale_factor=0.1)
  agent.initialize()
  learner_lib.Learner('/tmp/sac_bench', agent.train_step_counter, agent, strategy=strategy, checkpoint_interval=0)
  rb = rb_mod.TFUniformReplayBuffer(agent.collect_data_spec, batch_size=B_env, max_length=L, device=dev,
                                    seed=0x5eed0000 + strategy.rank)
  g = torch.Generator(device=dev).manual_seed(100 + strategy.rank)
  for _ in range(64):                      # enough history for T=2 windows; rows are synthetic
    rb.add_batch(trajectory.Trajectory(
        torch.ones(B_env, dtype=torch.int32, device=dev), torch.randn(B_env, D, device=dev, generator=g),
        torch.rand(B_env, A, device=dev, generator=g) * 2 - 1, (),
        torch.ones(B_env, dtype=torch.int32, device=dev), torch.rand(B_env, device=dev, generator=g),
        torch.ones(B_env, device=dev)))

  def step():
    exp, _ = rb.get_next(sample_batch_size=B, num_steps=2)
    return agent.train(exp).loss

  fn = common.function(step, warmup=1)
  fn()
  ok = 1
  try:
    fn()
  except Exception as e:  # keep ranks in lock-step if a capture fails somewhere
    sys.stderr.write(f'[rank {strategy.rank}] graph capture failed ({e}); eager\n')
    ok = 0
    torch.cuda.synchronize()
    step()
  if world > 1:
    flag = torch.tensor([ok], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = int(flag.item())
  if not ok:
    fn = step
  for _ in range(10):
    fn()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  K = 300
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(K):
    loss = fn()
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / K
  if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
  if strategy.rank == 0:
    print(json.dumps(dict(bench='sac_sharded_update', n_gpus=world, global_batch=B_global, per_gpu_batch=B,
                          ms_per_step=ms, steps_per_s=1000.0 / ms, cuda_graph=bool(ok),
                          loss=float(loss.item()), scaling='strong')), flush=True)
  if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == '__main__':
  main()
