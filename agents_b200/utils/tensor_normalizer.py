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

  def update(self, tensor, outer_dims=None):
    """Merges the batch statistics of `tensor` ([..., *spec.shape]) (:330-372)."""
    x = tensor.float().contiguous().reshape(-1, self._cols)
    rows = x.shape[0]
    ws, nb = workspace.get(x.device)
    _lib.call('b200rl_colsum', _lib.ptr(x), None, 0, rows, self._cols, 1.0 / rows,
              _lib.ptr(self._tmp_avg), _lib.ptr(ws), nb, _lib.stream())
    _lib.call('b200rl_colsum', _lib.ptr(x), _lib.ptr(self._tmp_avg), 1, rows, self._cols, 1.0,
              _lib.ptr(self._tmp_m2), _lib.ptr(ws), nb, _lib.stream())
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
