"""Per-(device, stream) scratch buffer handed to libb200rl (split-K partials, dcol, norm partials).

Grows on demand; growth during CUDA-graph capture is an error (warm up eagerly first, as
`utils.common.function` does).
"""
import torch

_WS = {}
_MIN_BYTES = 8 << 20
_SLOT = [0]


class slot(object):
  """Context manager selecting which scratch buffer `get` hands out (0 = main stream).

  Work forked onto a side stream runs under `with workspace.slot(1):` so that its split-K
  partials do not alias those of kernels running concurrently on the main stream.
  """

  def __init__(self, index):
    self._index = int(index)

  def __enter__(self):
    self._prev = _SLOT[0]
    _SLOT[0] = self._index

  def __exit__(self, *exc):
    _SLOT[0] = self._prev
    return False


def get(device, nbytes=0):
  """Returns (tensor, nbytes) of a uint8 scratch buffer of at least `nbytes` on `device`."""
  device = torch.device(device)
  index = device.index if device.index is not None else torch.cuda.current_device()
  # one buffer per slot: kernels on concurrent streams must not share split-K partials
  key = (device.type, index, _SLOT[0])
  cur = _WS.get(key)
  need = max(int(nbytes), _MIN_BYTES)
  if cur is None or cur.numel() < need:
    if torch.cuda.is_current_stream_capturing():
      raise RuntimeError('workspace must be sized before CUDA-graph capture; run the step '
                         'eagerly once first.')
    size = 1 << (need - 1).bit_length()
    cur = torch.empty(size, dtype=torch.uint8, device=device)
    _WS[key] = cur
  return cur, cur.numel()


def mirror(device, index):
  """Sizes slot `index` like slot 0 (call eagerly, before a capture that forks onto it)."""
  device = torch.device(device)
  dev_index = device.index if device.index is not None else torch.cuda.current_device()
  base = _WS.get((device.type, dev_index, 0))
  with slot(index):
    get(device, base.numel() if base is not None else 0)
