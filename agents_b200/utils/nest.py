"""Minimal nested-structure helpers (the subset of tf.nest the hot path relies on).

Ordering follows tf.nest: tuples/lists/namedtuples in field order, dicts in sorted-key order.
An empty tuple `()` is a valid structure with zero leaves (e.g. `policy_info=()`,
trajectories/trajectory.py:36-72 of the reference).
"""
import collections


def _is_namedtuple(x):
  return isinstance(x, tuple) and hasattr(x, '_fields')


def is_nested(x):
  return isinstance(x, (tuple, list, dict))


def flatten(structure):
  out = []

  def rec(s):
    if isinstance(s, dict):
      for k in sorted(s):
        rec(s[k])
    elif isinstance(s, (tuple, list)):
      for v in s:
        rec(v)
    else:
      out.append(s)

  rec(structure)
  return out


def pack_sequence_as(structure, flat):
  it = iter(flat)

  def rec(s):
    if isinstance(s, dict):
      vals = {k: rec(s[k]) for k in sorted(s)}
      if isinstance(s, collections.OrderedDict):
        return collections.OrderedDict((k, vals[k]) for k in s)
      return {k: vals[k] for k in s}
    if _is_namedtuple(s):
      return type(s)(*[rec(v) for v in s])
    if isinstance(s, tuple):
      return tuple(rec(v) for v in s)
    if isinstance(s, list):
      return [rec(v) for v in s]
    return next(it)

  packed = rec(structure)
  rest = list(it)
  if rest:
    raise ValueError(f'Structure has fewer leaves than the flat sequence ({len(rest)} left over).')
  return packed


def map_structure(fn, *structures):
  flats = [flatten(s) for s in structures]
  n = len(flats[0])
  for f in flats[1:]:
    if len(f) != n:
      raise ValueError('The two structures do not have the same number of leaves: '
                       f'{n} vs {len(f)}.')
  return pack_sequence_as(structures[0], [fn(*xs) for xs in zip(*flats)])


def assert_same_structure(a, b):
  def sig(s):
    if isinstance(s, dict):
      return ('dict', tuple((k, sig(s[k])) for k in sorted(s)))
    if isinstance(s, (tuple, list)):
      return ('seq', tuple(sig(v) for v in s))
    return 'leaf'

  if sig(a) != sig(b):
    raise ValueError("The two structures don't have the same nested structure.\n"
                     f'First structure: {a!r}\nSecond structure: {b!r}')
