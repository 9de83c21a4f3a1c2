"""PPOLearner on the GPU: preprocess_sequence and the minibatch pipeline against oracle/ppo.py +
oracle/ppo_learner.py (same seeded inputs, same shuffle order), losses within 2e-5 relative."""
import numpy as np
import pytest
import torch

from agents_b200 import optimizers
from agents_b200.agents.ppo import ppo_clip_agent
from agents_b200.networks import actor_distribution_network
from agents_b200.networks import layers as L
from agents_b200.networks import value_network
from agents_b200.specs import tensor_spec
from agents_b200.train import ppo_learner
from agents_b200.trajectories import time_step as ts
from agents_b200.trajectories import trajectory
from oracle import nn as onn
from oracle import optim as ooptim
from oracle import ppo as oppo
from oracle import ppo_learner as opl

pytestmark = pytest.mark.gpu
f32 = np.float32
D, A = 17, 6


def _build(cuda, **kw):
  obs_spec = tensor_spec.TensorSpec((D,), torch.float32, 'observation')
  act_spec = tensor_spec.BoundedTensorSpec((A,), torch.float32, -1.0, 1.0, 'action')
  actor = actor_distribution_network.ActorDistributionNetwork(
      obs_spec, act_spec, fc_layer_params=(64, 32), activation_fn='tanh', device=cuda).set_seed(1)
  value = value_network.ValueNetwork(obs_spec, fc_layer_params=(64, 32), activation_fn='tanh',
                                     device=cuda).set_seed(2)
  agent = ppo_clip_agent.PPOClipAgent(
      ts.time_step_spec(obs_spec), act_spec, optimizer=optimizers.Adam(1e-3), actor_net=actor,
      value_net=value, importance_ratio_clipping=0.2, use_gae=True, lambda_value=0.95,
      discount_factor=0.99, num_epochs=1, normalize_observations=False,
      compute_value_and_advantage_in_train=False, update_normalizers_in_train=False, **kw)
  agent.initialize()
  return agent, actor, value


def _mirror(net):
  return onn.Sequential([dict(kind='dense', w=l.kernel.cpu().numpy().copy(), b=l.bias.cpu().numpy().copy(),
                              act=l.activation) for l in net.layers if isinstance(l, L.Dense)])


def _oracle(actor, value, normalize_rewards):
  return oppo.PPOOracle(_mirror(actor), actor._std.bias.cpu().numpy().copy(), _mirror(value),
                        -np.ones(A, f32), np.ones(A, f32), ooptim.AdamTF(1e-3, eps=1e-7), num_epochs=1,
                        clip_eps=0.2, vf_coef=0.5, gamma=0.99, lam=0.95, normalize_rewards=normalize_rewards)


def _experience(rng, B, T):
  e = dict(observation=rng.randn(B, T, D).astype(f32), action=np.clip(rng.randn(B, T, A) * .5, -1, 1).astype(f32),
           loc=(rng.randn(B, T, A) * .2).astype(f32), scale=(rng.rand(B, T, A) * .3 + .5).astype(f32),
           reward=rng.rand(B, T).astype(f32), discount=np.ones((B, T), f32),
           step_type=np.ones((B, T), np.int32), next_step_type=np.ones((B, T), np.int32),
           value_prediction=rng.randn(B, T).astype(f32))
  ends = rng.rand(B, T) < 0.05
  e['next_step_type'][ends] = 2
  e['discount'][ends] = 0
  e['step_type'][:, 1:][ends[:, :-1]] = 2
  return e


def _to_traj(cuda, e):
  d = lambda a: torch.as_tensor(a, device=cuda)
  info = {'dist_params': {'loc': d(e['loc']), 'scale': d(e['scale'])},
          'value_prediction': d(e['value_prediction'])}
  return trajectory.Trajectory(d(e['step_type']), d(e['observation']), d(e['action']), info,
                               d(e['next_step_type']), d(e['reward']), d(e['discount']))


def test_preprocess_sequence_parity(cuda):
  rng = np.random.RandomState(11)
  agent, actor, value = _build(cuda, normalize_rewards=False)
  orc = _oracle(actor, value, False)
  e = _experience(rng, 6, 12)
  got = agent.preprocess_sequence(_to_traj(cuda, e))
  want = orc.preprocess_sequence(e)
  assert set(got.policy_info) == {'dist_params', 'value_prediction', 'return', 'advantage'}
  scale = np.abs(want['return']).max()
  np.testing.assert_allclose(got.policy_info['return'].cpu().numpy(), want['return'], rtol=1e-5, atol=1e-5 * scale)
  np.testing.assert_allclose(got.policy_info['advantage'].cpu().numpy(), want['advantage'], rtol=1e-5,
                             atol=1e-5 * scale)
  np.testing.assert_array_equal(got.policy_info['value_prediction'].cpu().numpy(), e['value_prediction'])
  # [T, ...] input keeps its rank (ppo_agent.py:744-747, :802-803)
  one = agent.preprocess_sequence(trajectory.Trajectory(*[
      torch.utils._pytree.tree_map(lambda t: t[0], f) for f in _to_traj(cuda, e)]))
  assert tuple(one.policy_info['return'].shape) == (12,)


@pytest.mark.parametrize('normalize_rewards', [False, True])
def test_minibatch_run_matches_oracle(cuda, tmp_path, normalize_rewards):
  rng = np.random.RandomState(5)
  B, T, mb, epochs, buf, seed = 16, 13, 30, 3, 50, 9   # 41 minibatches per inner dataset, 39 used per run
  agent, actor, value = _build(cuda, normalize_rewards=normalize_rewards)
  orc = _oracle(actor, value, normalize_rewards)
  batches = [_experience(rng, B, T) for _ in range(4)]        # 2 runs x num_samples=2
  cursor = {'train': 0, 'norm': 0}

  def dataset(kind):
    def gen():
      while True:
        e = batches[cursor[kind] % len(batches)]
        cursor[kind] += 1
        yield agent.preprocess_sequence(_to_traj(cuda, e)), ()
    return gen

  lrn = ppo_learner.PPOLearner(str(tmp_path), agent.train_step_counter, agent, dataset('train'),
                               dataset('norm'), num_samples=2, num_epochs=epochs, minibatch_size=mb,
                               shuffle_buffer_size=buf, checkpoint_interval=0, seed=seed)
  n = 2 * B * T
  per_run = opl.iterations_per_run(n, 2, epochs, mb, 1)
  leftover = []
  for run in range(2):
    info = lrn.run()
    assert lrn.num_frames_for_training == n
    # ---- oracle: same order of side effects -----------------------------------------------------
    es = batches[2 * run:2 * run + 2]
    if orc.reward_normalizer is not None:
      for e in es:                                            # _update_normalizers first (:270)
        orc.reward_normalizer.update(e['reward'])
    need = per_run - len(leftover)
    stream = list(leftover)
    if need > 0:                                              # a new inner dataset is cached
      pre = [orc.preprocess_sequence(e) for e in es]
      flat = {k: np.concatenate([p[k].reshape((B * T,) + p[k].shape[2:]) for p in pre], 0) for k in pre[0]}
      rows = opl.minibatch_rows(n, epochs, mb, buf, (seed ^ ppo_learner._SHUFFLE_SEED_TAG), run)
      stream += [{k: v[r][:, None] for k, v in flat.items()} for r in rows]
    infos = [orc.train(m, preprocessed=True, update_normalizers=False)[-1] for m in stream[:per_run]]
    leftover = stream[per_run:]
    np.testing.assert_allclose(info.loss.item(), infos[-1]['loss'], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(info.extra.value_estimation_loss.item(), infos[-1]['ve'], rtol=2e-5)
  assert int(agent.train_step_counter.item()) == 2 * per_run
  for v, w in zip(actor.variables + value.variables, orc.actor.params() + [orc.std_bias] + orc.value.params()):
    np.testing.assert_allclose(v.cpu().numpy(), w, rtol=2e-3, atol=3e-5)
  agent.check_numerics()


def test_full_sequence_mode(cuda, tmp_path):
  """minibatch_size=None: num_samples * num_epochs full [B, T] train calls (:293-294)."""
  rng = np.random.RandomState(2)
  agent, actor, value = _build(cuda, normalize_rewards=False)
  e = _experience(rng, 8, 9)
  ds = lambda: iter([(agent.preprocess_sequence(_to_traj(cuda, e)), ())] * 3)
  lrn = ppo_learner.PPOLearner(str(tmp_path), agent.train_step_counter, agent, ds, ds, num_samples=3,
                               num_epochs=2, checkpoint_interval=0)
  lrn.run()
  assert int(agent.train_step_counter.item()) == 6
  assert lrn.train_step_numpy == 6
