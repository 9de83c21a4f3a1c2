"""discounted_return / generalized_advantage_estimation on CUDA tensors.

Same signatures as tf_agents/utils/value_ops.py:21 and :102; the reverse scans run in
csrc/scans.cu (warp-shuffle affine scan for batch-major inputs).
"""
import torch

from agents_b200 import _lib


def _bt(t, time_major):
  if t.dim() == 1:
    return (1, t.shape[0]), True
  return ((t.shape[1], t.shape[0]) if time_major else (t.shape[0], t.shape[1])), False


def discounted_return(rewards, discounts, final_value=None, time_major=True,
                      provide_all_returns=True):
  rewards = rewards.float().contiguous()
  discounts = discounts.float().contiguous()
  (B, T), squeeze = _bt(rewards, time_major)
  if final_value is not None:
    final_value = torch.as_tensor(final_value, dtype=torch.float32,
                                  device=rewards.device).reshape(B).contiguous()
  out = torch.empty_like(rewards) if provide_all_returns else torch.empty(
      B, dtype=torch.float32, device=rewards.device)
  _lib.call('b200rl_discounted_return', _lib.ptr(rewards), _lib.ptr(discounts),
            _lib.ptr(final_value), _lib.ptr(out), B, T, int(bool(time_major) and not squeeze),
            int(bool(provide_all_returns)), _lib.stream())
  if not provide_all_returns and squeeze:
    out = out.reshape(())
  return out


def generalized_advantage_estimation(values, final_value, discounts, rewards, td_lambda=1.0,
                                     time_major=True):
  values = values.float().contiguous()
  discounts = discounts.float().contiguous()
  rewards = rewards.float().contiguous()
  (B, T), squeeze = _bt(values, time_major)
  final_value = torch.as_tensor(final_value, dtype=torch.float32,
                                device=values.device).reshape(B).contiguous()
  out = torch.empty_like(values)
  _lib.call('b200rl_gae', _lib.ptr(values), _lib.ptr(final_value), _lib.ptr(discounts),
            _lib.ptr(rewards), float(td_lambda), _lib.ptr(out), B, T,
            int(bool(time_major) and not squeeze), _lib.stream())
  return out
