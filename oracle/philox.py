"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11) in numpy.

Defines the sampling stream of this project (the reference's tf.random.uniform draws at
replay_buffers/tf_uniform_replay_buffer.py:265-272 are unseeded and unpinned).
counter = (elem_lo, elem_hi, call_lo, call_hi), key = (seed_lo, seed_hi); identical to
agents_b200/csrc/common.cuh.  Checked against the Random123 known-answer vectors in
tests/test_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10_raw(c0, c1, c2, c3, k0, k1):
  """Vectorised Philox on uint32 arrays; returns 4 uint32 arrays."""
  c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint32).copy() for x in (c0, c1, c2, c3)]
  k0 = np.uint32(k0)
  k1 = np.uint32(k1)
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = c0.astype(np.uint64) * M0
      p1 = c2.astype(np.uint64) * M1
      hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
      lo0 = (p0 & MASK32).astype(np.uint32)
      hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
      lo1 = (p1 & MASK32).astype(np.uint32)
      n0 = hi1 ^ c1 ^ k0
      n2 = hi0 ^ c3 ^ k1
      c0, c1, c2, c3 = n0, lo1, n2, lo0
      k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
      k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
  return c0, c1, c2, c3


def philox(elem, call, seed):
  """Philox block for element indices `elem` (uint64 array), call index and 64-bit seed."""
  elem = np.asarray(elem, dtype=np.uint64)
  call = int(call) & 0xFFFFFFFFFFFFFFFF
  seed = int(seed) & 0xFFFFFFFFFFFFFFFF
  c0 = (elem & MASK32).astype(np.uint32)
  c1 = (elem >> np.uint64(32)).astype(np.uint32)
  c2 = np.full(elem.shape, call & 0xFFFFFFFF, dtype=np.uint32)
  c3 = np.full(elem.shape, call >> 32, dtype=np.uint32)
  return philox4x32_10_raw(c0, c1, c2, c3, seed & 0xFFFFFFFF, seed >> 32)


def uniform_i64(a, b, lo, hi):
  """lo + ((b<<32)|a) % (hi-lo), int64."""
  u = (b.astype(np.uint64) << np.uint64(32)) | a.astype(np.uint64)
  return (np.int64(lo) + (u % np.uint64(hi - lo)).astype(np.int64)).astype(np.int64)


def uniform_f32(a):
  """[0,1) float32 from the top 24 bits."""
  return (a >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
