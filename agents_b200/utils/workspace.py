"""Per-(device, stream) scratch buffer handed to libb200rl (split-K partials, dcol, norm partials).

Grows on demand; growth during CUDA-graph capture is an error (warm up eagerly first, as
`utils.common.function` does).
"""
import torch

_WS = {}
_MIN_BYTES = 8 << 20


def get(device, nbytes=0):
  """Returns (tensor, nbytes) of a uint8 scratch buffer of at least `nbytes` on `device`."""
  device = torch.device(device)
  index = device.index if device.index is not None else torch.cuda.current_device()
  # one buffer per stream: kernels on concurrent streams must not share split-K partials
  key = (device.type, index, torch.cuda.current_stream(index).cuda_stream)
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
