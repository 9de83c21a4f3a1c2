"""Module alias matching tf_agents/policies/greedy_policy.py; see q_policy.py in this package."""
from agents_b200.policies.q_policy import *  # noqa: F401,F403
