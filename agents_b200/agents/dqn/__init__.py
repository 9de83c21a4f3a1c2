from agents_b200.agents.dqn import dqn_agent
