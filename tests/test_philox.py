"""Philox4x32-10 oracle against the Random123 known-answer vectors (kat_vectors)."""
import numpy as np

from oracle import philox

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_known_answer_vectors():
  for ctr, key, want in KAT:
    got = philox.philox4x32_10_raw([ctr[0]], [ctr[1]], [ctr[2]], [ctr[3]], key[0], key[1])
    assert tuple(int(g[0]) for g in got) == want


def test_counter_layout_matches_raw():
  elem = np.array([0x85a308d3243f6a88], dtype=np.uint64)
  got = philox.philox(elem, 0x0370734413198a2e, 0x299f31d0a4093822)
  assert tuple(int(g[0]) for g in got) == KAT[2][2]


def test_uniform_helpers():
  a = np.array([0, 0xffffffff, 256], dtype=np.uint32)
  f = philox.uniform_f32(a)
  assert f.dtype == np.float32 and f[0] == 0.0 and f[1] < 1.0 and f[2] == np.float32(2 ** -24)
  x = np.array([5, 7], dtype=np.uint32)
  y = np.array([0, 1], dtype=np.uint32)
  v = philox.uniform_i64(x, y, 10, 13)
  assert v.dtype == np.int64
  assert list(v) == [10 + 5 % 3, 10 + ((1 << 32) + 7) % 3]
