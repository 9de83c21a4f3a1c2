"""numpy restatement of utils/value_ops.py and the n-step reduction (TEST INFRASTRUCTURE).

discounted_return: utils/value_ops.py:21-99; generalized_advantage_estimation: :102-164;
n_step_reduce: trajectories/trajectory.py:815-832.  Serial float32 loops in the reference's
own association order (acc*disc + r; td + wd*acc).
"""
import numpy as np

f32 = np.float32


def discounted_return(rewards, discounts, final_value=None, time_major=True,
                      provide_all_returns=True):
  rewards = np.asarray(rewards, dtype=f32)
  discounts = np.asarray(discounts, dtype=f32)
  if not time_major:                                       # :61-64
    rewards, discounts = rewards.T, discounts.T
  if final_value is None:                                  # :66-67
    final_value = np.zeros_like(rewards[-1])
  acc = np.asarray(final_value, dtype=f32)
  out = np.zeros_like(rewards)
  for t in range(rewards.shape[0] - 1, -1, -1):            # tf.scan(reverse=True) / foldr
    acc = (acc * discounts[t]).astype(f32) + rewards[t]    # :73-75
    acc = acc.astype(f32)
    out[t] = acc
  if provide_all_returns:
    return out if time_major else out.T
  return acc


def generalized_advantage_estimation(values, final_value, discounts, rewards, td_lambda=1.0,
                                     time_major=True):
  values = np.asarray(values, dtype=f32)
  discounts = np.asarray(discounts, dtype=f32)
  rewards = np.asarray(rewards, dtype=f32)
  final_value = np.asarray(final_value, dtype=f32)
  if not time_major:                                       # :133-137
    values, discounts, rewards = values.T, discounts.T, rewards.T
  next_values = np.concatenate([values[1:], final_value[None]], axis=0)   # :140-142
  delta = ((rewards + (discounts * next_values).astype(f32)).astype(f32) - values).astype(f32)
  weighted_discounts = (discounts * f32(td_lambda)).astype(f32)           # :144
  acc = np.zeros_like(final_value)
  adv = np.zeros_like(values)
  for t in range(values.shape[0] - 1, -1, -1):                            # :146-158
    acc = (delta[t] + (weighted_discounts[t] * acc).astype(f32)).astype(f32)
    adv[t] = acc
  return adv if time_major else adv.T


def n_step_reduce(reward, discount, gamma):
  """reward/discount [B, N+1] -> (n-step reward [B], final discount [B])."""
  reward = np.asarray(reward, dtype=f32)
  discount = np.asarray(discount, dtype=f32)
  n = reward.shape[1] - 1
  r = reward[:, :-1]                                       # trajectory.py:815-816
  d = discount[:, :-1]
  discounted_reward = discounted_return(r, (f32(gamma) * d).astype(f32), time_major=False,
                                        provide_all_returns=False)        # :822-827
  prod = np.ones(d.shape[0], dtype=f32)
  for t in range(n):                                       # reduce_prod, left to right
    prod = (prod * d[:, t]).astype(f32)
  final_discount = (f32(gamma ** (n - 1)) * prod).astype(f32)             # :832
  return discounted_reward.astype(f32), final_discount
