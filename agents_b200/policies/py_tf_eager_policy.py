"""Module-path parity with tf_agents/policies/py_tf_eager_policy.py: `PyTFEagerPolicy` lives in
`agents_b200.policies.py_policy`."""
from agents_b200.policies.py_policy import PyTFEagerPolicy  # noqa: F401
