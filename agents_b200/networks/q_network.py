"""QNetwork (tf_agents/networks/q_network.py:60-158): encoder stack + a Dense Q head.

conv_layer_params items are (filters, kernel_size, stride); fc_layer_params are unit counts
(networks/encoding_network.py:240-300).  Kernels default to
variance_scaling(2.0, fan_in, truncated_normal) (encoding_network.py:224-226); the Q head uses
uniform(-0.03, 0.03) kernels and a -0.2 bias (q_network.py:126-135).
`preprocessing_layers=CastScale(255.)` restates the cast+/255 Lambda of the Atari example
(examples/dqn/mnih15/dqn_train_eval_atari.py:104).
"""
import torch

from agents_b200.networks import layers as L
from agents_b200.networks import network
from agents_b200.utils import nest


def validate_specs(action_spec, observation_spec):
  """q_network.py:36-57."""
  del observation_spec
  flat_action_spec = nest.flatten(action_spec)
  if len(flat_action_spec) > 1:
    raise ValueError('Network only supports action_specs with a single action.')
  if flat_action_spec[0].shape not in [(), (1,)]:
    raise ValueError('Network only supports action_specs with shape in [(), (1,)])')


def _vs_init(shape, fan_in, generator):
  return L.variance_scaling(shape, fan_in, 2.0, generator)


def _q_head_init(shape, fan_in, generator):
  return torch.rand(shape, dtype=torch.float32, generator=generator) * 0.06 - 0.03


class QNetwork(network.Network):

  def __init__(self, input_tensor_spec, action_spec, preprocessing_layers=None,
               conv_layer_params=None, fc_layer_params=(75, 40), activation_fn='relu',
               kernel_initializer=None, q_layer_activation_fn=None, name='QNetwork',
               device='cuda'):
    validate_specs(action_spec, input_tensor_spec)
    action_spec = nest.flatten(action_spec)[0]
    num_actions = int(action_spec.maximum - action_spec.minimum + 1)
    super(QNetwork, self).__init__(input_tensor_spec=input_tensor_spec, state_spec=(),
                                   name=name, device=device)
    kinit = kernel_initializer or _vs_init
    layers = []
    if preprocessing_layers is not None:
      layers.append(preprocessing_layers)
    for (filters, kernel_size, strides) in (conv_layer_params or []):
      layers.append(L.Conv2D(filters, kernel_size, strides, activation=activation_fn,
                             kernel_initializer=kinit))
    layers.append(L.Flatten())
    for units in (fc_layer_params or []):
      layers.append(L.Dense(units, activation=activation_fn, kernel_initializer=kinit))
    layers.append(L.Dense(num_actions, activation=q_layer_activation_fn,
                          kernel_initializer=_q_head_init,
                          bias_initializer=-0.2))
    self._layers = layers
