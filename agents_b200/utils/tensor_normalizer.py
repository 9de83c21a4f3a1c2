"""StreamingTensorNormalizer on CUDA (tf_agents/utils/tensor_normalizer.py:134-205,288-470).

Running count / mean / second moment per feature with Chan's parallel merge and a Kahan carry;
`normalize` is tf.nn.batch_normalization without offset/scale followed by clipping.  The column
moments, the merge and the normalisation are libb200rl launches (csrc/ppo.cu).
"""
import numpy as np
import torch

from agents_b200 import _lib
from agents_b200.utils import workspace

_EPS = 1e-10


def batch_moments(x, cols, avg, m2, stat_sync=None, replicas=1, rank=0):
  """Per-column mean and centred second moment of `x` ([rows, cols], contiguous fp32) written to
  the device vectors `avg` / `m2`; returns the number of rows they describe.

  With `replicas > 1` the rows are one shard of a global batch: every rank computes its local
  (mean, m2) in two passes, the R pairs are exchanged with ONE collective (a [R, 2*cols] buffer
  that is zero except for the caller's row, SUM-all-reduced by `stat_sync` -- i.e. an
  all-gather through the strategy's only primitive) and merged in rank order with Chan's
  parallel update (utils/tensor_normalizer.py:397-445), so all ranks end with the moments of the
  concatenated batch -- what a single device would have computed -- without the cancellation of a
  sum / sum-of-squares reduction.  Shards must have equal row counts (strategy.shard_range).
  """
  rows = x.shape[0]
  ws, nb = workspace.get(x.device)
  st = _lib.stream()
  _lib.call('b200rl_colsum', _lib.ptr(x), None, 0, rows, cols, 1.0 / rows, _lib.ptr(avg),
            _lib.ptr(ws), nb, st)
  _lib.call('b200rl_colsum', _lib.ptr(x), _lib.ptr(avg), 1, rows, cols, 1.0, _lib.ptr(m2),
            _lib.ptr(ws), nb, st)
  if stat_sync is None or replicas <= 1:
    return rows
  buf = torch.zeros((replicas, 2 * cols), dtype=torch.float32, device=x.device)
  buf[rank, :cols].copy_(avg)
  buf[rank, cols:].copy_(m2)
  stat_sync(buf)
  acc = torch.zeros((3, cols), dtype=torch.float32, device=x.device)   # count, carry, (unused)
  avg.zero_()
  m2.zero_()
  for r in range(replicas):
    _lib.call('b200rl_normalizer_update', _lib.ptr(acc[0]), _lib.ptr(avg), _lib.ptr(m2),
              _lib.ptr(acc[1]), buf[r].data_ptr(), buf[r].data_ptr() + 4 * cols, float(rows),
              cols, st)
  return rows * replicas


class StreamingTensorNormalizer(object):

  def __init__(self, tensor_spec, scope='normalize_tensor', device='cuda'):
    self._tensor_spec = tensor_spec
    shape = tuple(tensor_spec.shape)
    self._cols = int(np.prod(shape)) if shape else 1
    dev = torch.device(device)
    self._count = torch.full((self._cols,), _EPS, dtype=torch.float32, device=dev)
    self._avg = torch.zeros(self._cols, dtype=torch.float32, device=dev)
    self._m2 = torch.zeros(self._cols, dtype=torch.float32, device=dev)
    self._m2_carry = torch.zeros(self._cols, dtype=torch.float32, device=dev)
    self._tmp_avg = torch.zeros(self._cols, dtype=torch.float32, device=dev)
    self._tmp_m2 = torch.zeros(self._cols, dtype=torch.float32, device=dev)

  @property
  def variables(self):
    return (self._count, self._avg, self._m2, self._m2_carry)

  def update(self, tensor, outer_dims=None, stat_sync=None, replicas=1, rank=0):
    """Merges the batch statistics of `tensor` ([..., *spec.shape]) (:330-372).  In a
    data-parallel run (`stat_sync`, `replicas`, `rank` from the strategy) the batch statistics
    are those of the global batch, so every replica's running state stays identical."""
    x = tensor.float().contiguous().reshape(-1, self._cols)
    rows = batch_moments(x, self._cols, self._tmp_avg, self._tmp_m2, stat_sync, replicas, rank)
    _lib.call('b200rl_normalizer_update', _lib.ptr(self._count), _lib.ptr(self._avg),
              _lib.ptr(self._m2), _lib.ptr(self._m2_carry), _lib.ptr(self._tmp_avg),
              _lib.ptr(self._tmp_m2), float(rows), self._cols, _lib.stream())

  def normalize(self, tensor, clip_value=5.0, center_mean=True, variance_epsilon=1e-3):
    """(x - mean) / sqrt(var + eps), clipped to +-clip_value when > 0 (:134-205)."""
    x = tensor.float().contiguous()
    out = torch.empty_like(x)
    rows = x.numel() // self._cols
    _lib.call('b200rl_normalize', _lib.ptr(x), _lib.ptr(out), rows, self._cols,
              _lib.ptr(self._avg) if center_mean else None, _lib.ptr(self._m2),
              _lib.ptr(self._count), float(variance_epsilon), float(clip_value), _lib.stream())
    return out

  def reset(self):
    self._count.fill_(_EPS)
    for t in (self._avg, self._m2, self._m2_carry):
      t.zero_()
