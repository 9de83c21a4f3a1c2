"""Sequential network (tf_agents/networks/sequential.py): a plain stack of layers."""
from agents_b200.networks import network


class Sequential(network.Network):

  def __init__(self, layers, input_spec=None, name=None, device='cuda'):
    super(Sequential, self).__init__(input_tensor_spec=input_spec, state_spec=(), name=name,
                                     device=device)
    if not layers:
      raise ValueError('`layers` must not be empty.')
    self._layers = list(layers)
