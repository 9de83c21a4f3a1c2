"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_goldens.py).

CPU part: the oracle reproduces the committed bytes (drift guard).  GPU part: the CUDA path is
compared with the committed bytes through the product API — bit-exact for the ring, 1e-5
relative (to the trajectory's return scale) for the scans."""
import os

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import replay as oreplay
from oracle import value_ops as ovo

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
LEAVES = [('step_type', (), np.int32), ('observation', (17,), np.float32), ('action', (6,), np.float32),
          ('next_step_type', (), np.int32), ('reward', (), np.float32), ('discount', (), np.float32)]
B_ENV, L, ADDS, SEED = 5, 16, 21, 7
SAMPLES = [(4, 1), (7, 2), (16, 3)]


def _load(name):
  return np.load(os.path.join(HERE, name))


def test_oracle_reproduces_replay_fixture():
  g = _load('replay_mujoco_shape.npz')
  orc = oreplay.UniformReplayOracle([s for _, s, _ in LEAVES], [d for _, _, d in LEAVES], B_ENV, L,
                                    seed=SEED)
  for k in range(ADDS):
    orc.add_batch([g[f'add{k}_{n}'] for n, _, _ in LEAVES])
  assert orc.last_id == int(g['last_id']) == ADDS - 1
  for j, (B, T) in enumerate(SAMPLES):
    want, ids, rows, prob = orc.get_next(B, T)
    for (n, _, _), w in zip(LEAVES, want):
      np.testing.assert_array_equal(w, g[f'sample{j}_{n}'])
    np.testing.assert_array_equal(ids, g[f'sample{j}_ids'])
    np.testing.assert_array_equal(rows, g[f'sample{j}_rows'])
    np.testing.assert_array_equal(prob, g[f'sample{j}_prob'])
  for (n, _, _), w in zip(LEAVES, orc.gather_all()):
    np.testing.assert_array_equal(w, g[f'gather_all_{n}'])


def test_oracle_reproduces_value_ops_and_philox_fixtures():
  g = _load('value_ops.npz')
  r, d, v, fv = g['rewards'], g['discounts'], g['values'], g['final_value']
  np.testing.assert_array_equal(ovo.discounted_return(r, d, fv, time_major=False), g['returns'])
  np.testing.assert_array_equal(ovo.discounted_return(r, d, None, time_major=False), g['returns_no_final'])
  np.testing.assert_array_equal(ovo.generalized_advantage_estimation(v, fv, d, r, 0.95, False), g['gae_095'])
  np.testing.assert_array_equal(ovo.generalized_advantage_estimation(v, fv, d, r, 1.0, False), g['gae_100'])
  p = _load('philox.npz')
  words = np.array([philox.philox(e, int(p['call']), int(p['seed'])) for e in range(64)], dtype=np.uint32)
  np.testing.assert_array_equal(words, p['words'])


@pytest.mark.gpu
def test_cuda_ring_matches_replay_fixture(cuda):
  from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
  from agents_b200.specs import tensor_spec
  from agents_b200.trajectories import trajectory
  from agents_b200.utils import nest
  g = _load('replay_mujoco_shape.npz')
  td = {np.int32: torch.int32, np.float32: torch.float32}
  sp = {n: tensor_spec.TensorSpec(s, td[d], n) for n, s, d in LEAVES}
  spec = trajectory.Trajectory(step_type=sp['step_type'], observation=sp['observation'], action=sp['action'],
                               policy_info=(), next_step_type=sp['next_step_type'], reward=sp['reward'],
                               discount=sp['discount'])
  rb = rb_mod.TFUniformReplayBuffer(spec, batch_size=B_ENV, max_length=L, device=cuda, seed=SEED)
  for k in range(ADDS):
    rb.add_batch(nest.pack_sequence_as(
        spec, [torch.as_tensor(g[f'add{k}_{n}'], device=cuda) for n, _, _ in LEAVES]))
  for j, (B, T) in enumerate(SAMPLES):
    data, info = rb.get_next(sample_batch_size=B, num_steps=T)
    for (n, _, _), got in zip(LEAVES, nest.flatten(data)):
      np.testing.assert_array_equal(got.cpu().numpy(), g[f'sample{j}_{n}'])
    np.testing.assert_array_equal(info.ids.cpu().numpy(), g[f'sample{j}_ids'])
    np.testing.assert_array_equal(info.probabilities.cpu().numpy(), g[f'sample{j}_prob'])
  for (n, _, _), got in zip(LEAVES, nest.flatten(rb.gather_all())):
    np.testing.assert_array_equal(got.cpu().numpy(), g[f'gather_all_{n}'])


@pytest.mark.gpu
def test_cuda_scans_match_value_ops_fixture(cuda):
  from agents_b200.utils import value_ops
  g = _load('value_ops.npz')
  d = lambda k: torch.as_tensor(g[k], device=cuda)

  def close(got, want):
    scale = max(1.0, float(np.abs(want).max()))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6 * scale)

  close(value_ops.discounted_return(d('rewards'), d('discounts'), d('final_value'), time_major=False),
        g['returns'])
  close(value_ops.discounted_return(d('rewards'), d('discounts'), None, time_major=False),
        g['returns_no_final'])
  close(value_ops.generalized_advantage_estimation(d('values'), d('final_value'), d('discounts'),
                                                   d('rewards'), 0.95, False), g['gae_095'])
  close(value_ops.generalized_advantage_estimation(d('values'), d('final_value'), d('discounts'),
                                                   d('rewards'), 1.0, False), g['gae_100'])
