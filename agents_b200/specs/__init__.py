from agents_b200.specs.tensor_spec import BoundedTensorSpec
from agents_b200.specs.tensor_spec import TensorSpec
from agents_b200.specs import tensor_spec
