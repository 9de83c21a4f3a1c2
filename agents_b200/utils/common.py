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


def _object_state(obj):
  """Serializable host state of a checkpointable object."""
  if isinstance(obj, torch.Tensor):
    return {'tensor': obj.detach().cpu()}
  if hasattr(obj, 'train_step_counter') and hasattr(obj, 'collect_policy'):   # a TFAgent
    from agents_b200.train import learner as learner_lib
    return {'agent': {k: v.detach().cpu() for k, v in
                      learner_lib._agent_state_tensors(obj).items()},
            'train_step': int(obj.train_step_counter.item())}
  if hasattr(obj, 'state_dict'):
    return {'state_dict': obj.state_dict()}
  if hasattr(obj, 'variables'):                                               # policies, networks
    v = obj.variables() if callable(obj.variables) else obj.variables
    return {'variables': [t.detach().cpu() for t in v]}
  raise TypeError('Cannot checkpoint object of type {}'.format(type(obj).__name__))


def _restore_object(obj, state):
  if 'tensor' in state:
    obj.copy_(state['tensor'].to(obj.device))
  elif 'agent' in state:
    from agents_b200.train import learner as learner_lib
    cur = learner_lib._agent_state_tensors(obj)
    for k, v in state['agent'].items():
      if k in cur:
        cur[k].copy_(v.to(cur[k].device))
    obj.train_step_counter.fill_(state['train_step'])
    obj._train_step_host = state['train_step']
  elif 'state_dict' in state:
    obj.load_state_dict(state['state_dict'])
  else:
    v = obj.variables() if callable(obj.variables) else obj.variables
    for dst, src in zip(v, state['variables']):
      dst.copy_(src.to(dst.device))


class Checkpointer(object):
  """Checkpoints training state, policy state, and replay_buffer state
  (utils/common.py:1045-1100).

  `Checkpointer(ckpt_dir, max_to_keep, agent=..., replay_buffer=..., global_step=...)`: like the
  reference, the latest checkpoint (if any) is loaded into the objects on construction
  (:1074-1076); `save(global_step)` writes `ckpt-<step>.pt` and keeps the newest
  `max_to_keep`.  Agents (flat parameters, optimiser slots, train_step), replay buffers (ring
  storage, id table, last_id, Philox counters), tensors and anything with
  `state_dict/load_state_dict` or `variables()` can be listed.
  """

  def __init__(self, ckpt_dir, max_to_keep=20, **kwargs):
    import os
    self._dir = ckpt_dir
    self._max_to_keep = max_to_keep
    self._objects = kwargs
    os.makedirs(ckpt_dir, exist_ok=True)
    files = self._files()
    self._checkpoint_exists = bool(files)
    self._restored = False
    if files:
      self._restore(files[-1])

  def _files(self):
    import os
    names = [f for f in os.listdir(self._dir) if f.startswith('ckpt-') and f.endswith('.pt')]
    return [os.path.join(self._dir, f)
            for f in sorted(names, key=lambda f: int(f[5:-3]))]

  def _restore(self, path):
    state = torch.load(path, map_location='cpu', weights_only=False)
    for name, obj in self._objects.items():
      if name in state:
        _restore_object(obj, state[name])
    self._restored = True

  @property
  def checkpoint_exists(self):
    return self._checkpoint_exists

  def initialize_or_restore(self, session=None):
    """Objects were restored on construction; returns whether a checkpoint was loaded."""
    return self._restored

  def save(self, global_step, options=None):
    import os
    step = int(global_step.item()) if isinstance(global_step, torch.Tensor) else int(global_step)
    state = {name: _object_state(obj) for name, obj in self._objects.items()}
    path = os.path.join(self._dir, 'ckpt-{}.pt'.format(step))
    torch.save(state, path + '.tmp')
    os.replace(path + '.tmp', path)
    self._checkpoint_exists = True
    files = self._files()
    for f in files[:-self._max_to_keep] if self._max_to_keep else []:
      os.remove(f)
    return path


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
