"""DQN CartPole iteration benchmark (BASELINE.json configs[0] / SURVEY §8d config 1): the
reference's train_eval loop `collect 1 step -> sample 64 x 2 -> train` with obs f32[4], A=2,
B_env=1, L=10 000, net Dense(100, relu) -> Dense(2), squared loss, Adam 1e-3, tau=0.05 every 5
steps, gamma 0.99, epsilon 0.1, on the HBM-resident CartPole dynamics (csrc/env.cu).
Reports iterations/s eager and with the whole iteration (driver step + add_batch + get_next +
train) replayed as one CUDA graph.
NOT RUN in round 1 (written after the GPU budget was spent) - first thing to measure in round 2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agents_b200 import optimizers  # noqa: E402
from agents_b200.agents.dqn import dqn_agent  # noqa: E402
from agents_b200.drivers import dynamic_step_driver  # noqa: E402
from agents_b200.environments import random_tf_environment  # noqa: E402
from agents_b200.networks import layers as L  # noqa: E402
from agents_b200.networks import sequential  # noqa: E402
from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod  # noqa: E402
from agents_b200.utils import common  # noqa: E402


def main():
  dev = torch.device('cuda:0')
  env = random_tf_environment.CartPoleTFEnvironment(batch_size=1, seed=0, device=dev, action_dtype=torch.int32)
  tss, act_spec = env.time_step_spec(), env.action_spec()
  net = sequential.Sequential([L.Dense(100, activation='relu'), L.Dense(2)], input_spec=tss.observation,
                              device=dev).set_seed(0)
  agent = dqn_agent.DqnAgent(tss, act_spec, q_network=net, optimizer=optimizers.AdamOptimizer(1e-3), gamma=0.99,
                             epsilon_greedy=0.1, target_update_tau=0.05, target_update_period=5,
                             td_errors_loss_fn=common.element_wise_squared_loss)
  agent.initialize()
  rb = rb_mod.TFUniformReplayBuffer(agent.collect_data_spec, batch_size=1, max_length=10000, device=dev)
  driver = dynamic_step_driver.DynamicStepDriver(env, agent.collect_policy, observers=[rb.add_batch], num_steps=1)
  state = {'ts': None, 'ps': None}
  for _ in range(1000):                                   # initial_collect_steps (train_eval.py:99)
    state['ts'], state['ps'] = driver.run(state['ts'], state['ps'], maximum_iterations=1)

  def iteration():
    state['ts'], state['ps'] = driver.run(state['ts'], state['ps'], maximum_iterations=1)
    exp, _ = rb.get_next(sample_batch_size=64, num_steps=2)
    return agent.train(exp).loss

  def timed(fn, n):
    for _ in range(20):
      fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
      loss = fn()
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) * 1e-3), float(loss.item())

  eager, loss = timed(iteration, 500)
  out = dict(bench='dqn_cartpole_iteration', eager_iters_per_s=eager, loss=loss)
  try:
    graph_fn = common.function(iteration, warmup=1)
    out['graph_iters_per_s'], out['loss_graph'] = timed(graph_fn, 2000)
    rb.sync_last_id_from_device()
  except Exception as e:  # the driver's host bookkeeping may not be capturable on this stack
    out['graph_error'] = f'{type(e).__name__}: {e}'
  print(json.dumps(out), flush=True)


if __name__ == '__main__':
  main()
