"""Subset of tf_agents/utils/common.py used by the hot path.

function            :128   -> CUDA-graph capture/replay instead of tf.function
soft_variables_update :250-346, Periodically :450-507 (device-resident counter)
index_with_actions  :367-411, element_wise_squared_loss/huber_loss :1199-1208,
aggregate_losses    :1400-1476 (host composition; the fused DQN path uses csrc/dqn.cu)
"""
import collections

import torch

from agents_b200 import _lib
from agents_b200.utils import nest

# ---- host side effects that must be replayed together with a captured graph ----------------
_CAPTURE_EFFECTS = None


def record_host_effect(fn):
  """Called by objects that keep host mirrors of device counters (replay-buffer last_id,
  train_step).  During graph capture the effect is remembered and re-run on every replay."""
  if _CAPTURE_EFFECTS is not None:
    _CAPTURE_EFFECTS.append(fn)


class _GraphFunction(object):
  """Captures `fn(*args)` into one CUDA graph after `warmup` eager calls and replays it.

  Tensor arguments are copied into static input buffers before each replay; outputs are the
  static tensors produced at capture time (valid until the next call), like a tf.function
  running on a fixed input signature.
  """

  def __init__(self, fn, warmup=2):
    self._fn = fn
    self._warmup = warmup
    self._calls = 0
    self._graph = None
    self._static_in = None
    self._out = None
    self._effects = []

  def __call__(self, *args, **kwargs):
    global _CAPTURE_EFFECTS
    if not torch.cuda.is_available():
      raise _lib.B200RLError('common.function needs a CUDA device (no CPU fallback).')
    if self._graph is None:
      if self._calls < self._warmup:
        self._calls += 1
        return self._fn(*args, **kwargs)
      flat = nest.flatten((args, kwargs))
      self._static_in = [a.clone() if isinstance(a, torch.Tensor) else a for a in flat]
      s_args, s_kwargs = nest.pack_sequence_as((args, kwargs), self._static_in)
      torch.cuda.synchronize()
      self._graph = torch.cuda.CUDAGraph()
      _CAPTURE_EFFECTS = []
      try:
        # thread_local: other threads (the NCCL watchdog polling its events, a data-loader
        # thread) may keep calling CUDA while this thread captures
        with torch.cuda.graph(self._graph, capture_error_mode='thread_local'):
          self._out = self._fn(*s_args, **s_kwargs)
      finally:
        self._effects, _CAPTURE_EFFECTS = _CAPTURE_EFFECTS, None
      # capture does not execute; replay once for this call (inputs are already in place;
      # host effects already ran once during capture).
      self._graph.replay()
      return self._out
    flat = nest.flatten((args, kwargs))
    for dst, src in zip(self._static_in, flat):
      if isinstance(dst, torch.Tensor) and src is not dst:
        dst.copy_(src, non_blocking=True)
    self._graph.replay()
    for e in self._effects:
      e()
    return self._out


def function(fn=None, warmup=2, **unused_tf_function_kwargs):
  """Drop-in for `common.function` (utils/common.py:128): compiles a step into one CUDA graph."""
  if fn is None:
    return lambda f: _GraphFunction(f, warmup)
  return _GraphFunction(fn, warmup)


def soft_variables_update(source, target, tau=1.0, tau_non_trainable=None,
                          sort_variables_by_name=False, period=1, counter=None):
  """target = (1-tau)*target + tau*source over flat buffers (utils/common.py:250-346).

  `source`/`target` are Networks or flat fp32 tensors.  `period`/`counter` fold the
  `Periodically` gate (:450-507) into the same launch.
  """
  if tau < 0 or tau > 1:
    raise ValueError('Input `tau` should be in [0, 1].')
  src = source.flat_params if hasattr(source, 'flat_params') else source
  dst = target.flat_params if hasattr(target, 'flat_params') else target
  if src.numel() != dst.numel():
    raise ValueError('Source and target variable lists have different lengths: '
                     '{} vs. {}'.format(src.numel(), dst.numel()))
  _lib.call('b200rl_soft_update', _lib.ptr(dst), _lib.ptr(src), dst.numel(), float(tau),
            int(period), _lib.ptr(counter), _lib.stream())


class Periodically(object):
  """Runs `body(period, counter)` gated on a device-resident counter (utils/common.py:450-507).

  The gate is evaluated inside the body's kernel (see b200rl_soft_update), so the call is
  graph-capturable; `period=None` is a no-op and `period=1` always fires, as in the reference.
  """

  def __init__(self, body, period, name='periodically', device='cuda'):
    if not callable(body):
      raise TypeError('body must be callable.')
    self._body = body
    self._period = period
    self._counter = torch.zeros(2, dtype=torch.int64, device=device)

  def __call__(self):
    if self._period is None:
      return
    self._body(int(self._period), self._counter)


def index_with_actions(q_values, actions, multi_dim_actions=False):
  """q_values[..., actions] (utils/common.py:367-411)."""
  if multi_dim_actions:
    raise NotImplementedError('multi_dim_actions is not supported on the hot path.')
  return torch.gather(q_values, -1, actions.long().unsqueeze(-1)).squeeze(-1)


def element_wise_squared_loss(x, y):
  return (x - y) ** 2


def element_wise_huber_loss(x, y):
  e = (y - x).abs()
  quad = torch.clamp(e, max=1.0)
  return 0.5 * quad * quad + (e - quad)


AggregatedLosses = collections.namedtuple('AggregatedLosses',
                                          ['total_loss', 'weighted', 'regularization'])
