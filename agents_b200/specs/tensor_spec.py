"""TensorSpec / BoundedTensorSpec for torch tensors.

Host-side mirror of tf_agents/specs/tensor_spec.py (TensorSpec, BoundedTensorSpec :235-311
sampling rules are not reproduced; see environments/random_tf_environment.py here).
"""
import numpy as np
import torch

_NP_TO_TORCH = {
    np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64,
    np.dtype('int32'): torch.int32, np.dtype('int64'): torch.int64,
    np.dtype('uint8'): torch.uint8, np.dtype('int8'): torch.int8,
    np.dtype('int16'): torch.int16, np.dtype('bool'): torch.bool,
    np.dtype('float16'): torch.float16,
}
_TORCH_TO_NP = {v: k for k, v in _NP_TO_TORCH.items()}


def as_torch_dtype(dtype):
  if isinstance(dtype, torch.dtype):
    return dtype
  return _NP_TO_TORCH[np.dtype(dtype)]


def as_numpy_dtype(dtype):
  if isinstance(dtype, torch.dtype):
    return _TORCH_TO_NP[dtype]
  return np.dtype(dtype)


class TensorSpec(object):
  """Shape + dtype (+ name) of one leaf; shape excludes batch/time dimensions."""

  def __init__(self, shape, dtype, name=None):
    self._shape = tuple(int(d) for d in shape)
    self._dtype = as_torch_dtype(dtype)
    self._name = name

  @property
  def shape(self):
    return self._shape

  @property
  def dtype(self):
    return self._dtype

  @property
  def name(self):
    return self._name

  @property
  def itemsize(self):
    return torch.empty((), dtype=self._dtype).element_size()

  @property
  def row_bytes(self):
    return int(np.prod(self._shape, dtype=np.int64)) * self.itemsize

  def is_compatible_with(self, tensor):
    return (tuple(tensor.shape) == self._shape and
            as_torch_dtype(tensor.dtype) == self._dtype)

  def __repr__(self):
    return f'{type(self).__name__}(shape={self._shape}, dtype={self._dtype}, name={self._name!r})'

  def __eq__(self, other):
    return (type(self) is type(other) and self._shape == other._shape and
            self._dtype == other._dtype)

  def __hash__(self):
    return hash((self._shape, self._dtype))


class BoundedTensorSpec(TensorSpec):
  """TensorSpec with inclusive minimum / maximum."""

  def __init__(self, shape, dtype, minimum, maximum, name=None):
    super().__init__(shape, dtype, name)
    self._minimum = np.asarray(minimum, dtype=as_numpy_dtype(self.dtype))
    self._maximum = np.asarray(maximum, dtype=as_numpy_dtype(self.dtype))

  @property
  def minimum(self):
    return self._minimum

  @property
  def maximum(self):
    return self._maximum

  def __repr__(self):
    return (f'BoundedTensorSpec(shape={self.shape}, dtype={self.dtype}, name={self.name!r}, '
            f'minimum={self._minimum}, maximum={self._maximum})')

  def __eq__(self, other):
    return (super().__eq__(other) and np.array_equal(self._minimum, other._minimum) and
            np.array_equal(self._maximum, other._maximum))

  def __hash__(self):
    return hash((self.shape, self.dtype))
