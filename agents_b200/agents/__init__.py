from agents_b200.agents import tf_agent
from agents_b200.agents.tf_agent import LossInfo
from agents_b200.agents.tf_agent import TFAgent
