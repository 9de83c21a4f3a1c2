from agents_b200.train import learner
from agents_b200.train.learner import Learner
