"""Generates the committed golden fixtures of tests/golden/ from the CPU oracle.

TensorFlow cannot be installed in the build container (no wheel, no network), so fixtures cannot
be produced by importing the reference; the reference's OWN expected values are replayed
literally in tests/test_oracle_goldens.py / test_oracle_ppo.py / test_oracle_sac.py.  The files
written here freeze seeded input/output vectors of the (golden-pinned) oracle so that
  * `-m "not gpu"`: the oracle cannot drift silently (tests/test_golden_fixtures.py), and
  * `-m gpu`: the CUDA path is compared with committed bytes, not only with live oracle code.

Run from the repo root:  python tests/golden/make_goldens.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import philox  # noqa: E402
from oracle import replay as oreplay  # noqa: E402
from oracle import value_ops as ovo  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32

# MuJoCo-shape PPO/SAC row: Trajectory leaves in flatten order
LEAVES = [('step_type', (), np.int32), ('observation', (17,), np.float32), ('action', (6,), np.float32),
          ('next_step_type', (), np.int32), ('reward', (), np.float32), ('discount', (), np.float32)]
B_ENV, L, ADDS, SEED = 5, 16, 21, 7          # 21 adds into 16 slots: the ring wraps
SAMPLES = [(4, 1), (7, 2), (16, 3)]


def replay_fixture():
  rng = np.random.RandomState(0)
  orc = oreplay.UniformReplayOracle([s for _, s, _ in LEAVES], [d for _, _, d in LEAVES], B_ENV, L,
                                    seed=SEED)
  out = {}
  for k in range(ADDS):
    items = []
    for name, shape, dt in LEAVES:
      if np.issubdtype(dt, np.integer):
        items.append(rng.randint(0, 3, size=(B_ENV,) + shape).astype(dt))
      else:
        items.append(rng.rand(*((B_ENV,) + shape)).astype(dt))
      out[f'add{k}_{name}'] = items[-1]
    orc.add_batch(items)
  for j, (B, T) in enumerate(SAMPLES):
    want, ids, rows, prob = orc.get_next(B, T)
    for (name, _, _), w in zip(LEAVES, want):
      out[f'sample{j}_{name}'] = w
    out[f'sample{j}_ids'], out[f'sample{j}_rows'], out[f'sample{j}_prob'] = ids, rows, prob
  for (name, _, _), w in zip(LEAVES, orc.gather_all()):
    out[f'gather_all_{name}'] = w
  out['last_id'] = np.int64(orc.last_id)
  np.savez_compressed(os.path.join(HERE, 'replay_mujoco_shape.npz'), **out)


def value_ops_fixture():
  rng = np.random.RandomState(1)
  B, T = 5, 33
  r = rng.randn(B, T).astype(f32)
  d = (0.99 * (rng.rand(B, T) > 0.1)).astype(f32)
  v = rng.randn(B, T).astype(f32)
  fv = rng.randn(B).astype(f32)
  np.savez_compressed(
      os.path.join(HERE, 'value_ops.npz'), rewards=r, discounts=d, values=v, final_value=fv,
      returns=ovo.discounted_return(r, d, fv, time_major=False),
      returns_no_final=ovo.discounted_return(r, d, None, time_major=False),
      gae_095=ovo.generalized_advantage_estimation(v, fv, d, r, 0.95, False),
      gae_100=ovo.generalized_advantage_estimation(v, fv, d, r, 1.0, False))


def philox_fixture():
  elems = np.arange(64, dtype=np.uint64)
  words = np.array([philox.philox(int(e), 3, 0x5eed0000) for e in elems], dtype=np.uint32)
  np.savez_compressed(os.path.join(HERE, 'philox.npz'), words=words, call=np.uint64(3),
                      seed=np.uint64(0x5eed0000))


if __name__ == '__main__':
  replay_fixture()
  value_ops_fixture()
  philox_fixture()
  print('wrote', sorted(f for f in os.listdir(HERE) if f.endswith('.npz')))
