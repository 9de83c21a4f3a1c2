"""The C-ABI library loads and exports every symbol include/b200rl.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'b200rl.h')


def _declared():
  src = open(HEADER).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(b200rl_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_symbols():
  names = _declared()
  assert 'b200rl_rb_sample' in names and 'b200rl_dqn_td_loss' in names and len(names) >= 25


def test_library_exports_every_declared_symbol():
  from agents_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    import __graft_entry__
    __graft_entry__.build()
  l = ctypes.CDLL(_lib.LIB_PATH)
  missing = [n for n in _declared() if not hasattr(l, n)]
  assert not missing, missing


def test_ctypes_signatures_cover_header():
  from agents_b200 import _lib
  bound = set(_lib.SIGNATURES) | set(_lib._RESTYPES)
  assert set(_declared()) == bound


def test_no_gpu_calls_fail_loudly_not_silently():
  """Without CUDA tensors the product path raises instead of falling back to the CPU."""
  import torch
  from agents_b200 import _lib
  with pytest.raises(_lib.B200RLError):
    _lib.ptr(torch.zeros(4))


def test_struct_layout_matches_header():
  from agents_b200 import _lib
  assert ctypes.sizeof(_lib.Leaf) == 16
  assert ctypes.sizeof(_lib.Ring) == 8 + 8 + 8 + 8 + 8 + 8 + 16 * _lib.MAX_LEAVES
  assert ctypes.sizeof(_lib.ConvGeom) == 40


def test_integration_md_stub_matches_the_binding():
  """The ctypes stub printed in INTEGRATION.md must stay loadable and layout-compatible with the
  binding the package uses (struct layout, symbol, argument count)."""
  import ctypes
  import os
  import re
  from agents_b200 import _lib
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  text = open(os.path.join(root, 'INTEGRATION.md')).read()
  block = re.search(r"```python\nimport ctypes\n(.*?)```", text, re.S).group(1)
  block = block.replace("ctypes.CDLL('libb200rl.so')", "ctypes.CDLL(%r)" % _lib.LIB_PATH)
  ns = {'ctypes': ctypes}
  exec('import ctypes\n' + block, ns)              # defines lib, Leaf, Ring, fused_read
  assert ctypes.sizeof(ns['Ring']) == ctypes.sizeof(_lib.Ring)
  assert ctypes.sizeof(ns['Leaf']) == ctypes.sizeof(_lib.Leaf)
  for (name, _), (name2, _) in zip(ns['Ring']._fields_, _lib.Ring._fields_):
    assert name == name2
    assert getattr(ns['Ring'], name).offset == getattr(_lib.Ring, name2).offset
  assert len(ns['lib'].b200rl_rb_read_rows.argtypes) == len(_lib.SIGNATURES['b200rl_rb_read_rows'])
  assert callable(ns['fused_read'])


def test_every_entry_point_is_documented_in_integration_md():
  """INTEGRATION.md's entry-point table must mention every symbol the header declares
  (shorthands `a_fwd/bwd`, `a[_ld]` and `a_*` are expanded)."""
  import os
  import re
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  header = open(os.path.join(root, 'include', 'b200rl.h')).read()
  doc = open(os.path.join(root, 'INTEGRATION.md')).read()
  symbols = sorted(set(re.findall(r'\b(b200rl_[a-z0-9_]+)\s*\(', header)))
  covered = set()
  for m in re.finditer(r'`(b200rl_[a-z0-9_]+)/([a-z0-9_/]+)`', doc):
    stem = m.group(1)[:m.group(1).rfind('_') + 1]
    covered.update([m.group(1)] + [stem + a for a in m.group(2).split('/')])
  for m in re.finditer(r'`(b200rl_[a-z0-9_]+)\[(_[a-z0-9]+)\]`', doc):
    covered.update([m.group(1), m.group(1) + m.group(2)])
  prefixes = [m.group(1) + '_' for m in re.finditer(r'`(b200rl_[a-z0-9_]+)_\*`', doc)]
  missing = [s for s in symbols
             if s not in doc and s not in covered and not any(s.startswith(p) for p in prefixes)]
  assert not missing, missing
