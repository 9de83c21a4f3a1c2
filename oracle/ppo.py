"""numpy restatement of PPOAgent / PPOClipAgent update math (TEST INFRASTRUCTURE).

Follows agents/ppo/ppo_agent.py: _normalize_advantages :100-110, compute_advantages :440-479,
get_loss :481-615, compute_return_and_advantage :617-719, _preprocess :721-807, _train :834-1076,
entropy_regularization_loss :1159-1201, value_estimation_loss :1203-1327,
policy_gradient_loss :1329-1512, kl_cutoff_loss :1514-1539, adaptive_kl_loss :1541-1558,
kl_penalty_loss :1586-1630, update_adaptive_kl_beta :1632-1675;
agents/ppo/ppo_utils.py:35-59 (make_trajectory_mask), :194-227 (nested_kl_divergence);
utils/common.py:682-755 (log_probability / entropy summed over action dims), :883-895
(get_episode_mask), :1400-1476 (aggregate_losses); utils/tensor_normalizer.py:134-205,288-470.
The Normal log-prob / entropy / KL formulas are TFP's closed forms (tensorflow_probability is an
un-vendored dependency; Normal.kl_divergence is `_kl_normal_normal` in
tfp/distributions/normal.py: 0.5*sqdiff(mu_a/s_b, mu_b/s_b) + 0.5*expm1(2d) - d, d = log s_a - log s_b).
"""
import numpy as np

from oracle import nn
from oracle import optim
from oracle import value_ops

f32 = np.float32
LOG2PI = f32(np.log(2 * np.pi))
STEP_LAST = 2


def normal_log_prob(loc, scale, action):
  """sum_k log N(a_k; loc_k, scale_k)  (common.log_probability sums over the action dims)."""
  z = (action - loc) / scale
  return np.sum(-0.5 * z * z - np.log(scale) - 0.5 * LOG2PI, axis=-1).astype(f32)


def normal_entropy(scale):
  return np.sum(0.5 + 0.5 * LOG2PI + np.log(scale), axis=-1).astype(f32)


def normal_kl(loc_a, scale_a, loc_b, scale_b):
  """sum_k KL(N(loc_a, scale_a) || N(loc_b, scale_b)) over the action dims
  (ppo_utils.nested_kl_divergence :194-227 reduces the non-batch dims)."""
  d = (np.log(scale_a) - np.log(scale_b)).astype(f32)
  z = (loc_a / scale_b - loc_b / scale_b).astype(f32)
  return np.sum(f32(0.5) * z * z + f32(0.5) * np.expm1(f32(2) * d) - d, axis=-1).astype(f32)


def kl_cutoff_loss(kl, kl_cutoff_factor, adaptive_kl_target, kl_cutoff_coef):
  """:1514-1539."""
  if kl_cutoff_factor <= 0:
    return f32(0)
  cutoff = f32(kl_cutoff_factor * adaptive_kl_target)
  over = np.maximum(np.mean(kl, dtype=f32) - cutoff, f32(0))
  return f32(f32(kl_cutoff_coef) * over * over)


def adaptive_kl_loss(kl, beta):
  """:1541-1558 (beta None -> 0)."""
  if beta is None:
    return f32(0)
  return f32(f32(beta) * np.mean(kl, dtype=f32))


def update_adaptive_kl_beta(beta, kl, adaptive_kl_target, adaptive_kl_tolerance):
  """:1632-1675; returns the new beta."""
  if beta is None:
    return None
  mean_kl = np.mean(kl, dtype=f32)
  factor = f32(1)
  if mean_kl < f32(adaptive_kl_target) * f32(1.0 - adaptive_kl_tolerance):
    factor = f32(1.0 / 1.5)
  elif mean_kl > f32(adaptive_kl_target) * f32(1.0 + adaptive_kl_tolerance):
    factor = f32(1.5)
  return f32(np.clip(f32(beta) * factor, f32(10e-16), f32(10e16)))


def aggregate(per_example, weights, global_batch=None):
  """common.aggregate_losses (:1400-1476): weight, mean over non-batch dims, sum / global B."""
  x = per_example * weights
  x = np.where(weights == 0, f32(0), x)                    # multiply_no_nan
  if x.ndim > 1:
    x = x.reshape(x.shape[0], -1).mean(axis=1, dtype=f32)
  gb = f32(global_batch if global_batch is not None else x.shape[0])
  return f32(np.sum(x, dtype=f32) / gb)


def policy_gradient_loss(logp, sample_logp, advantages, weights, clip_eps, logp_clip=0.0,
                         global_batch=None):
  """:1364-1417. Returns (loss, clip_fraction)."""
  if logp_clip > 0:
    logp = np.clip(logp, -logp_clip, logp_clip)
  ratio = np.exp(logp - sample_logp).astype(f32)
  ratio_c = np.clip(ratio, 1 - clip_eps, 1 + clip_eps).astype(f32)
  obj, obj_c = ratio * advantages, ratio_c * advantages
  per = -np.minimum(obj, obj_c) if clip_eps > 0 else -obj
  clip_frac = f32(np.mean((np.abs(ratio - 1.0) > clip_eps).astype(f32))) if clip_eps > 0 else f32(0)
  return aggregate(per.astype(f32), weights, global_batch), clip_frac


def value_estimation_loss(value_preds, returns, weights, vf_coef, value_clip=0.0,
                          old_value_preds=None, global_batch=None):
  """:1262-1295."""
  err = (returns - value_preds) ** 2
  if value_clip > 0:
    vc = old_value_preds + np.clip(value_preds - old_value_preds, -value_clip, value_clip)
    err = np.maximum(err, (returns - vc) ** 2)
  return f32(aggregate(err.astype(f32), weights, global_batch) * f32(vf_coef))


def entropy_regularization_loss(entropy, weights, ent_coef, global_batch=None):
  """:1167-1180."""
  if ent_coef <= 0:
    return f32(0)
  return f32(aggregate((-entropy).astype(f32), weights, global_batch) * f32(ent_coef))


def normalize_advantages(adv, eps=1e-8):
  """tf.nn.moments over axes (0,1) + tf.nn.batch_normalization (:100-110), unmasked."""
  mean = np.mean(adv, dtype=f32)
  var = np.mean((adv - mean) ** 2, dtype=f32)
  inv = f32(1.0) / np.sqrt(var + f32(eps), dtype=f32)
  return (adv * inv + (-mean * inv)).astype(f32)


def make_trajectory_mask(step_type, returns, advantages):
  """ppo_utils.py:35-59."""
  return ((step_type != STEP_LAST) & ~((returns == 0) & (advantages == 0))).astype(f32)


def compute_return_and_advantage(reward, discount, next_step_type, value_preds, gamma, lam,
                                 use_gae=True, use_td_lambda_return=False):
  """reward/discount/next_step_type are the [B, T-1] next_time_steps fields; value_preds [B, T].
  Returns (returns, advantages) of shape [B, T-1] (:617-719)."""
  discounts = (discount * f32(gamma)).astype(f32)                      # :632-634
  episode_mask = (next_step_type != STEP_LAST).astype(f32)             # common.py:883-895
  discounts = (discounts * episode_mask).astype(f32)                   # :660
  final_value_bootstrapped = value_preds[:, -1]                        # :662
  returns = value_ops.discounted_return(reward, discounts, final_value_bootstrapped,
                                        time_major=False)              # :663-668
  vp = value_preds[:, :-1]                                             # :464
  if use_gae:                                                          # :465-473 (final_value sic)
    advantages = value_ops.generalized_advantage_estimation(
        values=vp, final_value=vp[:, -1], rewards=reward, discounts=discounts, td_lambda=lam,
        time_major=False)
  else:
    advantages = (returns - vp).astype(f32)
  if use_td_lambda_return and use_gae:                                 # :708-717
    returns = (advantages + vp).astype(f32)
  return returns.astype(f32), advantages.astype(f32)


class StreamingNormalizer(object):
  """utils/tensor_normalizer.py:288-470 (count starts at _EPS = 1e-10)."""

  def __init__(self, shape):
    self.count = np.full(shape, 1e-10, f32)
    self.avg = np.zeros(shape, f32)
    self.m2 = np.zeros(shape, f32)
    self.carry = np.zeros(shape, f32)

  def update(self, x):
    x = np.asarray(x, f32).reshape((-1,) + self.avg.shape)
    n_a = f32(x.shape[0])
    avg_a = x.mean(axis=0, dtype=f32)
    m2_a = ((x - avg_a) ** 2).sum(axis=0, dtype=f32)
    n_b, avg_b, m2_b, c = self.count, self.avg, self.m2, self.carry
    n_ab = n_a + n_b
    delta = avg_b - avg_a
    s_delta = delta * n_b / n_ab
    avg_ab = avg_a + s_delta
    value = m2_a + delta * n_a * s_delta
    y = value - c                                                       # kahan_summation
    t = m2_b + y
    self.carry = ((t - m2_b) - y).astype(f32)
    self.m2 = t.astype(f32)
    self.count, self.avg = n_ab.astype(f32), avg_ab.astype(f32)

  def normalize(self, x, clip_value=5.0, center_mean=True, variance_epsilon=1e-3):
    var = self.m2 / self.count
    mean = self.avg if center_mean else np.zeros_like(self.avg)
    inv = f32(1.0) / np.sqrt(var + f32(variance_epsilon), dtype=f32)
    out = (x * inv + (-mean * inv)).astype(f32)
    if clip_value > 0:
      out = np.clip(out, -clip_value, clip_value)
    return out.astype(f32)


def softplus(x):
  return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20)))).astype(f32)


class PPOOracle(object):
  """PPOClipAgent-style learner on numpy nets.

  actor trunk: oracle.nn.Sequential ending in a linear Dense(A) (mean layer); the
  NormalProjectionNetwork head gives loc = shift + half_range * tanh(m_raw),
  scale = softplus(std_bias).  value net: Sequential ending in Dense(1).
  """

  def __init__(self, actor, std_bias, value, amin, amax, optimizer, num_epochs=25,
               clip_eps=0.2, vf_coef=0.5, ent_coef=0.0, gamma=0.99, lam=0.95, value_clip=0.0,
               logp_clip=0.0, gradient_clipping=None, normalize_rewards=False,
               reward_norm_clipping=10.0, use_gae=True, use_td_lambda_return=False,
               kl_cutoff_factor=0.0, kl_cutoff_coef=0.0, initial_adaptive_kl_beta=0.0,
               adaptive_kl_target=0.0, adaptive_kl_tolerance=0.0, normalize_observations=False,
               obs_dim=None):
    self.actor, self.value, self.std_bias = actor, value, np.asarray(std_bias, f32)
    self.amin, self.amax = np.asarray(amin, f32), np.asarray(amax, f32)
    self.opt = optimizer
    self.num_epochs, self.clip_eps, self.vf_coef, self.ent_coef = num_epochs, clip_eps, vf_coef, ent_coef
    self.gamma, self.lam, self.value_clip, self.logp_clip = gamma, lam, value_clip, logp_clip
    self.gradient_clipping = gradient_clipping
    self.use_gae, self.use_td = use_gae, use_td_lambda_return
    self.reward_normalizer = StreamingNormalizer(()) if normalize_rewards else None
    self.reward_norm_clipping = reward_norm_clipping
    self.train_step_counter = 0
    self.kl_cutoff_factor, self.kl_cutoff_coef = kl_cutoff_factor, kl_cutoff_coef
    self.initial_beta = initial_adaptive_kl_beta
    self.beta = f32(initial_adaptive_kl_beta) if initial_adaptive_kl_beta > 0 else None   # :339-345
    self.kl_target, self.kl_tol = adaptive_kl_target, adaptive_kl_tolerance
    # observation normaliser: applied by the policy before both networks (ppo_policy.py)
    self.obs_normalizer = StreamingNormalizer((obs_dim,)) if normalize_observations else None

  def norm_obs(self, obs):
    if self.obs_normalizer is None:
      return obs
    return self.obs_normalizer.normalize(obs)       # defaults: clip 5, centred (:134-205)

  # -- policy head -----------------------------------------------------------------------------
  def dist(self, obs, keep=False):
    m_raw, tape = self.actor.forward(obs, keep=True)
    half, shift = (self.amax - self.amin) / 2, (self.amax + self.amin) / 2
    loc = (shift + half * np.tanh(m_raw)).astype(f32)
    scale = np.broadcast_to(softplus(self.std_bias), loc.shape).astype(f32)
    return (loc, scale, m_raw, tape) if keep else (loc, scale)

  def params(self):
    return self.actor.params() + [self.std_bias] + self.value.params()

  def preprocess(self, exp):
    """_preprocess (:721-807): value preds on all T steps, returns / advantages padded with 0."""
    B, T = exp['reward'].shape
    obs = exp['observation']
    vp = self.value.forward(self.norm_obs(obs.reshape(B * T, -1))).reshape(B, T)
    reward = exp['reward'][:, :-1]
    if self.reward_normalizer is not None:                                     # :651-654
      reward = self.reward_normalizer.normalize(reward, center_mean=False,
                                                clip_value=self.reward_norm_clipping)
    ret, adv = compute_return_and_advantage(reward, exp['discount'][:, :-1],
                                            exp['next_step_type'][:, :-1], vp, self.gamma,
                                            self.lam, self.use_gae, self.use_td)
    pad = np.zeros((B, 1), f32)
    return vp, np.concatenate([ret, pad], 1), np.concatenate([adv, pad], 1)

  def loss_and_grads(self, obs, action, old_logp, ret, adv_n, v_old, w, B, T, global_batch=None,
                     old_loc=None, old_scale=None):
    """get_loss (:481-615) + hand-written backward. All inputs flattened to N = B*T."""
    N = B * T
    gb = f32(global_batch or B)
    denom = f32(T) * gb
    loc, scale, m_raw, atape = self.dist(obs, keep=True)
    v, vtape = self.value.forward(obs, keep=True)
    v = v[:, 0]
    logp = normal_log_prob(loc, scale, action)
    ent = normal_entropy(scale)
    sh = (B, T)
    pg, clip_frac = policy_gradient_loss(logp.reshape(sh), old_logp.reshape(sh), adv_n.reshape(sh),
                                         w.reshape(sh), self.clip_eps, self.logp_clip, gb)
    ve = value_estimation_loss(v.reshape(sh), ret.reshape(sh), w.reshape(sh), self.vf_coef,
                               self.value_clip, None if v_old is None else v_old.reshape(sh), gb)
    en = entropy_regularization_loss(ent.reshape(sh), w.reshape(sh), self.ent_coef, gb)
    use_kl = not (self.initial_beta == 0 and self.kl_cutoff_factor == 0)           # :586
    klp, g_kl = f32(0), np.zeros(N, f32)
    if use_kl:                                                                     # :1586-1630
      kl = (normal_kl(old_loc, old_scale, loc, scale) * w).astype(f32)
      cut = kl_cutoff_loss(kl, self.kl_cutoff_factor, self.kl_target, self.kl_cutoff_coef)
      ada = adaptive_kl_loss(kl, self.beta)
      klp = f32(cut + ada)
      mean_kl = np.mean(kl, dtype=f32)
      dmean = f32(0)
      if self.kl_cutoff_factor > 0:
        dmean += f32(2 * self.kl_cutoff_coef) * np.maximum(
            mean_kl - f32(self.kl_cutoff_factor * self.kl_target), f32(0))
      if self.beta is not None:
        dmean += f32(self.beta)
      g_kl = (dmean * w / f32(N)).astype(f32)
    total = f32(pg + ve + en + klp)
    # ---- backward (derived independently of csrc/ppo.cu; checked against autograd in tests)
    lp, dlp = logp, np.ones(N, f32)
    if self.logp_clip > 0:
      dlp = ((logp >= -self.logp_clip) & (logp <= self.logp_clip)).astype(f32)
      lp = np.clip(logp, -self.logp_clip, self.logp_clip)
    ratio = np.exp(lp - old_logp)
    lo, hi = 1 - self.clip_eps, 1 + self.clip_eps
    ratio_c = np.clip(ratio, lo, hi)
    if self.clip_eps > 0:
      unclipped_is_min = ratio * adv_n <= ratio_c * adv_n
      inside = (ratio >= lo) & (ratio <= hi)
      dobj = np.where(unclipped_is_min, adv_n, np.where(inside, adv_n, 0.0))
    else:
      dobj = adv_n
    g_logp = (-dobj * ratio * dlp * w / denom).astype(f32)
    g_ent = (-f32(self.ent_coef) * w / denom).astype(f32) if self.ent_coef > 0 else np.zeros(N, f32)
    d = action - loc
    inv = 1.0 / scale
    dloc = (g_logp[:, None] * d * inv * inv).astype(f32)
    dscale = (g_logp[:, None] * (d * d * inv ** 3 - inv) + g_ent[:, None] * inv).astype(f32)
    if use_kl:       # d KL(old || new) / d(new loc, new scale)
      dm = old_loc - loc
      dloc = (dloc + g_kl[:, None] * (-dm * inv * inv)).astype(f32)
      dscale = (dscale + g_kl[:, None] * (inv - (dm * dm + old_scale ** 2) * inv ** 3)).astype(f32)
    err = (ret - v) ** 2
    derr = -2 * (ret - v)
    if self.value_clip > 0:
      dv_ = v - v_old
      vc = v_old + np.clip(dv_, -self.value_clip, self.value_clip)
      err_c = (ret - vc) ** 2
      use_c = err_c > err
      derr = np.where(use_c, np.where(np.abs(dv_) <= self.value_clip, -2 * (ret - vc), 0.0), derr)
    dv = (f32(self.vf_coef) * derr * w / denom).astype(f32)
    half = (self.amax - self.amin) / 2
    dm_raw = (dloc * half * (1 - np.tanh(m_raw) ** 2)).astype(f32)
    sig = 1.0 / (1.0 + np.exp(-self.std_bias))
    dstd = (dscale * sig).sum(axis=0).astype(f32)
    grads = self.actor.backward(atape, dm_raw) + [dstd] + self.value.backward(vtape, dv[:, None])
    return dict(loss=total, pg=pg, ve=ve, ent=en, clip_fraction=clip_frac, kl=klp), grads

  def preprocess_sequence(self, exp):
    """_preprocess_sequence (:809-832) for compute_value_and_advantage_in_train=False: value
    predictions come from policy_info (:775-776); returns / advantages are stored next to them
    (:789-800).  Returns a copy of `exp` with 'return' and 'advantage' added."""
    B, T = exp['reward'].shape
    vp = np.asarray(exp['value_prediction'], f32)
    reward = exp['reward'][:, :-1]
    if self.reward_normalizer is not None:
      reward = self.reward_normalizer.normalize(reward, center_mean=False,
                                                clip_value=self.reward_norm_clipping)
    ret, adv = compute_return_and_advantage(reward, exp['discount'][:, :-1],
                                            exp['next_step_type'][:, :-1], vp, self.gamma,
                                            self.lam, self.use_gae, self.use_td)
    pad = np.zeros((B, 1), f32)
    out = dict(exp)
    out['return'] = np.concatenate([ret, pad], 1)
    out['advantage'] = np.concatenate([adv, pad], 1)
    return out

  def train(self, exp, weights=None, preprocessed=False, update_normalizers=True):
    """_train (:834-1076): preprocess once (or take the stored return/advantage/value_prediction
    when compute_value_and_advantage_in_train=False, :843-846), then num_epochs steps."""
    B, T = exp['reward'].shape
    if preprocessed:
      vp, ret, adv = (np.asarray(exp[k], f32) for k in ('value_prediction', 'return', 'advantage'))
    else:
      vp, ret, adv = self.preprocess(exp)
    mask = make_trajectory_mask(exp['step_type'], ret, adv)                     # :845
    w = mask if weights is None else (np.asarray(weights, f32) * mask).astype(f32)
    A = exp['action'].shape[-1]
    action = exp['action'].reshape(B * T, A)
    old_loc, old_scale = exp['loc'].reshape(B * T, A), exp['scale'].reshape(B * T, A)
    old_logp = normal_log_prob(old_loc, old_scale, action)                      # :867-869
    adv_n = normalize_advantages(adv)                                           # :893-895
    raw_obs = exp['observation'].reshape(B * T, -1)
    obs = self.norm_obs(raw_obs)
    infos = []
    for _ in range(self.num_epochs):                                            # :925-967
      info, grads = self.loss_and_grads(obs, action, old_logp, ret.reshape(-1), adv_n.reshape(-1),
                                        vp.reshape(-1), w.reshape(-1), B, T,
                                        old_loc=old_loc, old_scale=old_scale)
      if self.gradient_clipping:
        grads, _ = optim.clip_by_global_norm(grads, self.gradient_clipping)
      self.opt.apply(self.params(), grads)
      self.train_step_counter += 1
      infos.append(info)
    if self.initial_beta > 0:                                                   # :978-989
      loc, scale = self.dist(obs)
      kl = (normal_kl(old_loc, old_scale, loc, scale) * w.reshape(-1)).astype(f32)
      self.beta = update_adaptive_kl_beta(self.beta, kl, self.kl_target, self.kl_tol)
    if update_normalizers:                                                      # :991-993
      if self.obs_normalizer is not None:
        self.obs_normalizer.update(raw_obs)
      if self.reward_normalizer is not None:
        self.reward_normalizer.update(exp['reward'])
    return infos
