from agents_b200.trajectories import policy_step
from agents_b200.trajectories import time_step
from agents_b200.trajectories import trajectory
from agents_b200.trajectories.policy_step import PolicyStep
from agents_b200.trajectories.time_step import StepType
from agents_b200.trajectories.time_step import TimeStep
from agents_b200.trajectories.trajectory import Trajectory
from agents_b200.trajectories.trajectory import Transition
