"""numpy restatement of the device-resident environments and epsilon-greedy (TEST INFRASTRUCTURE).

random_env_step: environments/random_tf_environment.py:96-129 with per-env termination and the
auto-reset contract of environments/tf_environment.py:211-241 / trajectories/time_step.py:135-195.
cartpole_step: gym classic_control CartPole-v1 formulae (the reference loads it through
suite_gym, agents/dqn/examples/v2/train_eval.py:151) + TimeLimit truncation.
epsilon_greedy: policies/epsilon_greedy_policy.py:120-145.
Random numbers: Philox stream of oracle/philox.py, identical indexing to csrc/env.cu.
"""
import numpy as np

from oracle import philox

f32 = np.float32
FIRST, MID, LAST = 0, 1, 2
STATE_DOMAIN = 1 << 62


def epsilon_greedy(q, eps, seed, call, mask=None):
  B, A = q.shape
  x, y, _, _ = philox.philox(np.arange(B, dtype=np.uint64), call, seed)
  u = philox.uniform_f32(x)
  if mask is None:
    rnd = (y % np.uint32(A)).astype(np.int32)
    qm = q
  else:
    m = np.asarray(mask).astype(bool)
    qm = np.where(m, q, np.finfo(np.float32).min)
    rnd = np.zeros(B, dtype=np.int32)
    for b in range(B):
      allowed = np.flatnonzero(m[b])
      rnd[b] = allowed[int(y[b]) % len(allowed)] if len(allowed) else 0
  greedy = np.argmax(qm, axis=1).astype(np.int32)
  return np.where(u >= f32(eps), greedy, rnd).astype(np.int32)     # :126-145


def random_env_step(step_type, obs_elems, obs_is_u8, p_term, seed, call):
  """Returns (new_step_type, obs, reward, discount) for a batch of envs."""
  B = step_type.shape[0]
  if obs_is_u8:
    vec = (obs_elems + 15) // 16
    x, y, z, w = philox.philox(np.arange(B * vec, dtype=np.uint64), call, seed)
    words = np.stack([x, y, z, w], axis=1).astype('<u4')
    obs = words.view(np.uint8).reshape(B, vec * 16)[:, :obs_elems].copy()
  else:
    vec = (obs_elems + 3) // 4
    x, y, z, w = philox.philox(np.arange(B * vec, dtype=np.uint64), call, seed)
    words = np.stack([x, y, z, w], axis=1)
    vals = (words >> np.uint32(8)).astype(f32) * f32(1.0 / 8388608.0) - f32(1.0)
    obs = vals.reshape(B, vec * 4)[:, :obs_elems].copy()
  sx, sy, _, _ = philox.philox(np.uint64(STATE_DOMAIN) + np.arange(B, dtype=np.uint64), call, seed)
  was_last = step_type == LAST
  term = philox.uniform_f32(sy) < f32(p_term)
  reward = np.where(was_last, f32(0), philox.uniform_f32(sx)).astype(f32)
  new_type = np.where(was_last, FIRST, np.where(term, LAST, MID)).astype(np.int32)
  discount = np.where(was_last, f32(1), np.where(term, f32(0), f32(1))).astype(f32)
  return new_type, obs, reward, discount


def cartpole_step(state, steps, step_type, action, max_steps, seed, call):
  B = state.shape[0]
  state = state.astype(f32).copy()
  steps = steps.copy()
  x, xd, th, thd = [state[:, i] for i in range(4)]
  r0, r1, r2, r3 = philox.philox(np.arange(B, dtype=np.uint64), call, seed)
  reset = step_type == LAST
  gravity, masspole, total_mass, length = f32(9.8), f32(0.1), f32(1.1), f32(0.5)
  polemass_length, force_mag, tau = f32(0.05), f32(10.0), f32(0.02)
  force = np.where(action == 1, force_mag, -force_mag).astype(f32)
  c, s = np.cos(th).astype(f32), np.sin(th).astype(f32)
  temp = (force + polemass_length * thd * thd * s) / total_mass
  thacc = (gravity * s - c * temp) / (length * (f32(4.0 / 3.0) - masspole * c * c / total_mass))
  xacc = temp - polemass_length * thacc * c / total_mass
  nx, nxd = x + tau * xd, xd + tau * xacc
  nth, nthd = th + tau * thd, thd + tau * thacc
  n = steps + 1
  fell = (nx < -2.4) | (nx > 2.4) | (nth < f32(-0.20943951)) | (nth > f32(0.20943951))
  trunc = n >= max_steps
  u = [philox.uniform_f32(r) * f32(0.1) - f32(0.05) for r in (r0, r1, r2, r3)]
  new_state = np.stack([np.where(reset, u[0], nx), np.where(reset, u[1], nxd),
                        np.where(reset, u[2], nth), np.where(reset, u[3], nthd)], axis=1).astype(f32)
  new_steps = np.where(reset, 0, n).astype(np.int32)
  new_type = np.where(reset, FIRST, np.where(fell | trunc, LAST, MID)).astype(np.int32)
  reward = np.where(reset, f32(0), f32(1)).astype(f32)
  discount = np.where(reset, f32(1), np.where(fell, f32(0), f32(1))).astype(f32)
  return new_state, new_steps, new_type, new_state.copy(), reward, discount
