"""create_train_step (tf_agents/train/utils/train_utils.py:39-46): a device-resident int64
counter shared by agent and learner."""
import torch


def create_train_step(device='cuda'):
  return torch.zeros((), dtype=torch.int64, device=device)
