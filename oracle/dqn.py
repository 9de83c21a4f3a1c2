"""numpy restatement of DqnAgent / DdqnAgent loss and train step (TEST INFRASTRUCTURE).

Follows agents/dqn/dqn_agent.py: compute_td_targets :75-78, _train :412-449, _td_loss :451-460,
_loss :462-579, _compute_q_values :581-602, _compute_next_q_values :604-645 (DdqnAgent
:659-700), target updater :385-409; utils/common.py: index_with_actions :367-411,
element_wise_squared/huber_loss :1199-1208, aggregate_losses :1400-1476;
policies/greedy_policy.py:70-89 + q_policy.py:150-194 for the (masked) greedy action.
"""
import copy

import numpy as np

from oracle import nn
from oracle import optim
from oracle import value_ops

f32 = np.float32
STEP_LAST = 2  # trajectories/time_step.py:113-121


def index_with_actions(q_values, actions):
  # utils/common.py:367-411 (single-dim actions): gather_nd over (batch_index, action)
  b = np.arange(q_values.shape[0])
  return q_values[b, np.asarray(actions, dtype=np.int64)]


def huber(targets, predictions, delta=1.0):
  # tf.compat.v1.losses.huber_loss, reduction NONE
  error = (predictions - targets).astype(f32)
  abs_error = np.abs(error)
  quadratic = np.minimum(abs_error, f32(delta))
  linear = (abs_error - quadratic).astype(f32)
  return (f32(0.5) * (quadratic * quadratic).astype(f32) + f32(delta) * linear).astype(f32)


def squared(targets, predictions):
  e = (predictions - targets).astype(f32)
  return (e * e).astype(f32)


def greedy_action(q, mask=None):
  # q_policy.py:183-191: masked logits -> dtype.min; greedy_policy.py:73: mode == first argmax
  if mask is not None:
    q = np.where(np.asarray(mask).astype(bool), q, np.finfo(np.float32).min)
  return np.argmax(q, axis=1)


def compute_td_targets(next_q_values, rewards, discounts):
  # dqn_agent.py:75-78
  return (rewards + (discounts * next_q_values).astype(f32)).astype(f32)


def dqn_loss(q_s0, next_q_target, next_q_select, actions, step_type0, traj_reward,
             traj_discount, gamma=1.0, reward_scale=1.0, loss_fn='huber', weights=None,
             next_mask=None, global_batch=None, reg_loss=None):
  """Everything after the network evaluations in DqnAgent._loss (:494-579).

  Returns dict(loss, td_loss, td_error, dq) where dq = dLoss/dq_s0.
  """
  q_s0 = np.asarray(q_s0, dtype=f32)
  B, A = q_s0.shape
  # AsNStepTransition -> to_n_step_transition (trajectory.py:815-832)
  R, D = value_ops.n_step_reduce(traj_reward, traj_discount, gamma)
  q_values = index_with_actions(q_s0, actions)                               # :501
  best = greedy_action(next_q_select, next_mask)                             # :634 / :688
  next_q = index_with_actions(np.asarray(next_q_target, dtype=f32), best)    # :641-645
  rewards = (f32(reward_scale) * R).astype(f32)                              # :507
  discounts = (f32(gamma) * D).astype(f32)                                   # :508
  td_targets = compute_td_targets(next_q, rewards, discounts)
  td_error = (td_targets - q_values).astype(f32)                             # :457
  fn = huber if loss_fn == 'huber' else squared
  td_loss = fn(td_targets, q_values)                                         # :458
  valid_mask = (np.asarray(step_type0) != STEP_LAST).astype(f32)             # :514
  td_error = valid_mask * td_error
  td_loss = valid_mask * td_loss
  per_example = td_loss
  w = np.ones(B, dtype=f32)
  if weights is not None:                                                    # common.py:1427-1441
    w = np.broadcast_to(np.asarray(weights, dtype=f32), (B,)).copy()
    per_example = np.where(w == 0, f32(0), per_example * w).astype(f32)     # multiply_no_nan
  gb = f32(global_batch if global_batch is not None else B)
  loss = f32(np.sum(per_example, dtype=f32) / gb)                            # common.py:1465-1468
  total = loss if reg_loss is None else f32(loss + f32(reg_loss))
  # backward of the epilogue
  e = (td_targets - q_values).astype(f32)
  if loss_fn == 'huber':
    dl_dq = -np.clip(e, -1.0, 1.0).astype(f32)
  else:
    dl_dq = (-2.0 * e).astype(f32)
  dq = np.zeros((B, A), dtype=f32)
  dq[np.arange(B), np.asarray(actions, dtype=np.int64)] = valid_mask * w * dl_dq / gb
  return dict(loss=total, td_loss=td_loss, td_error=td_error, dq=dq, weighted=loss)


class DqnOracle(object):
  """One-replica DQN/DDQN learner on numpy Sequentials (dqn_agent.py:82-645)."""

  def __init__(self, q_net, optimizer, gamma=1.0, reward_scale=1.0, n_step_update=1,
               loss_fn='huber', target_update_tau=1.0, target_update_period=1,
               gradient_clipping=None, ddqn=False):
    self.q_net = q_net
    self.target_net = copy.deepcopy(q_net)            # maybe_copy_target_network_with_checks
    self.optimizer = optimizer
    self.gamma, self.reward_scale, self.n = gamma, reward_scale, n_step_update
    self.loss_fn = loss_fn
    self.gradient_clipping = gradient_clipping
    self.ddqn = ddqn
    self.train_step_counter = 0
    self._update_target = optim.Periodically(                                 # :385-409
        lambda: optim.soft_variables_update(self.q_net.params(), self.target_net.params(),
                                            target_update_tau), target_update_period)

  def loss(self, exp, weights=None, keep=False):
    """exp: dict of [B,T,...] arrays with keys step_type, observation, action, reward, discount."""
    obs0 = exp['observation'][:, 0]
    obsn = exp['observation'][:, -1]
    q_s0, tape = self.q_net.forward(obs0, keep=True)
    next_t = self.target_net.forward(obsn)
    next_sel = self.q_net.forward(obsn) if self.ddqn else next_t
    out = dqn_loss(q_s0, next_t, next_sel, exp['action'][:, 0], exp['step_type'][:, 0],
                   exp['reward'], exp['discount'], self.gamma, self.reward_scale, self.loss_fn,
                   weights)
    if keep:
      out['tape'] = tape
    return out

  def train(self, exp, weights=None):
    out = self.loss(exp, weights, keep=True)                                  # :413-421
    grads = self.q_net.backward(out['tape'], out['dq'])                       # :426
    if self.gradient_clipping is not None:                                    # :429-432
      grads = [optim.clip_by_norm(g, self.gradient_clipping) for g in grads]
    self.optimizer.apply(self.q_net.params(), grads)                          # :444
    self.train_step_counter += 1                                              # :445
    self._update_target()                                                     # :447
    out['grads'] = grads
    return out
