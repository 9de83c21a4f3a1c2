"""ActorDistributionNetwork with a NormalProjectionNetwork head for continuous actions.

Reference: networks/actor_distribution_network.py (MLP encoder, default fc_layer_params
(200, 100)) + networks/normal_projection_network.py: means = Dense(A, variance_scaling(0.1)),
squashed to the action spec with tanh (`tanh_squash_to_spec`); standard deviations are a
state-independent bias passed through softplus (std_bias_initializer_value=0.0).
The head transform and its backward run in csrc/ppo.cu (b200rl_normal_proj_fwd/bwd).
"""
import numpy as np
import torch

from agents_b200 import _lib
from agents_b200.networks import layers as L
from agents_b200.networks import network
from agents_b200.utils import nest
from agents_b200.utils import workspace


def _mean_init(shape, fan_in, generator):
  return L.variance_scaling(shape, fan_in, 0.1, generator)


class StdBias(L.Layer):
  """The BiasLayer of NormalProjectionNetwork: a trainable [A] vector; identity in the chain."""
  has_params = True

  def __init__(self, num_actions, init_value=0.0):
    self.num_actions = num_actions
    self.init_value = float(init_value)
    self.l2 = 0.0

  def build(self, input_shape):
    return tuple(input_shape)

  def param_shapes(self):
    return [(self.num_actions,)]

  def bind(self, params, grads):
    self.bias, self.d_bias = params[0], grads[0]

  def init_params(self, generator):
    self.bias.fill_(self.init_value)

  def forward(self, x):
    return x

  def backward(self, x, y, dy, need_dx):
    return dy


class ActorDistributionNetwork(network.Network):
  """Emits the parameters (loc, scale) of a diagonal Normal over bounded continuous actions."""

  def __init__(self, input_tensor_spec, output_tensor_spec, fc_layer_params=(200, 100),
               activation_fn='relu', kernel_initializer=None, std_bias_initializer_value=0.0,
               name='ActorDistributionNetwork', device='cuda'):
    super().__init__(input_tensor_spec=input_tensor_spec, state_spec=(), name=name, device=device)
    spec = nest.flatten(output_tensor_spec)
    if len(spec) != 1 or len(spec[0].shape) != 1:
      raise ValueError('ActorDistributionNetwork here supports one rank-1 continuous action.')
    self._action_spec = spec[0]
    self.num_actions = int(spec[0].shape[0])
    from agents_b200.networks import q_network
    kinit = kernel_initializer or q_network._vs_init  # EncodingNetwork default initializer
    layers = []
    for units in (fc_layer_params or []):
      layers.append(L.Dense(units, activation=activation_fn, kernel_initializer=kinit))
    layers.append(L.Dense(self.num_actions, activation=None, kernel_initializer=_mean_init))
    self._std = StdBias(self.num_actions, std_bias_initializer_value)
    layers.append(self._std)
    self._layers = layers
    amin = np.broadcast_to(np.asarray(spec[0].minimum, np.float32), (self.num_actions,)).copy()
    amax = np.broadcast_to(np.asarray(spec[0].maximum, np.float32), (self.num_actions,)).copy()
    self._amin = torch.as_tensor(amin, device=self._device)
    self._amax = torch.as_tensor(amax, device=self._device)

  @property
  def action_min(self):
    return self._amin

  @property
  def action_max(self):
    return self._amax

  def _head(self, m_raw):
    n = m_raw.shape[0]
    loc = torch.empty_like(m_raw)
    scale = torch.empty_like(m_raw)
    _lib.call('b200rl_normal_proj_fwd', _lib.ptr(m_raw), _lib.ptr(self._std.bias),
              _lib.ptr(self._amin), _lib.ptr(self._amax), n, self.num_actions, _lib.ptr(loc),
              _lib.ptr(scale), _lib.stream())
    return loc, scale

  def distribution_params(self, observation):
    """(loc, scale) each [N, A], no tape."""
    m_raw, _ = self._run(observation, keep=False)
    return self._head(m_raw)

  def forward_train(self, observation):
    m_raw, tape = self._run(observation, keep=True)
    loc, scale = self._head(m_raw)
    return (loc, scale), (tape, m_raw)

  def backward(self, ctx, dparams):
    """dparams = (dloc, dscale), each [N, A] contiguous."""
    tape, m_raw = ctx
    dloc, dscale = dparams
    n = m_raw.shape[0]
    dm = torch.empty_like(m_raw)
    ds_part = torch.empty_like(m_raw)
    _lib.call('b200rl_normal_proj_bwd', _lib.ptr(m_raw), _lib.ptr(self._std.bias),
              _lib.ptr(self._amin), _lib.ptr(self._amax), _lib.ptr(dloc), _lib.ptr(dscale), n,
              self.num_actions, _lib.ptr(dm), _lib.ptr(ds_part), _lib.stream())
    grads = super().backward(tape, dm)      # zeroes the flat gradient, then accumulates
    ws, nb = workspace.get(m_raw.device)
    _lib.call('b200rl_colsum', _lib.ptr(ds_part), None, 0, n, self.num_actions, 1.0,
              _lib.ptr(self._std.d_bias), _lib.ptr(ws), nb, _lib.stream())
    return grads

  def __call__(self, observation, step_type=None, network_state=(), training=False):
    return self.distribution_params(observation), network_state
