"""discounted_return / GAE / n-step kernels vs the oracle (fp32, 1e-5 relative)."""
import numpy as np
import pytest
import torch

from agents_b200.trajectories import trajectory
from agents_b200.utils import value_ops
from oracle import value_ops as ovo

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6
f32 = np.float32


def _close(got, want):
  """1e-5 relative; the absolute floor scales with the magnitude of the returns in the batch
  because the warp-shuffle scan re-associates acc*d+r (SURVEY.md §7 "scan association order"):
  a return that cancels to ~0 keeps the rounding error of its O(scale) partial sums."""
  want = np.asarray(want)
  scale = max(1.0, float(np.abs(want).max())) if want.size else 1.0
  np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=ATOL * scale)


def _inputs(rng, B, T, p_end=0.1):
  r = rng.randn(B, T).astype(f32)
  d = (0.99 * (rng.rand(B, T) > p_end)).astype(f32)      # zeros at episode ends
  v = rng.randn(B, T).astype(f32)
  fv = rng.randn(B).astype(f32)
  return r, d, v, fv


@pytest.mark.parametrize('B,T', [(1, 1), (3, 5), (7, 32), (5, 33), (64, 128), (9, 1000), (4096, 128)])
def test_discounted_return_parity(cuda, B, T):
  rng = np.random.RandomState(B * 1000 + T)
  r, d, _, fv = _inputs(rng, B, T)
  for time_major in (False, True):
    rr, dd = (r.T.copy(), d.T.copy()) if time_major else (r, d)
    for final in (None, fv):
      want = ovo.discounted_return(rr, dd, final, time_major=time_major)
      got = value_ops.discounted_return(torch.as_tensor(rr, device=cuda), torch.as_tensor(dd, device=cuda),
                                        None if final is None else torch.as_tensor(final, device=cuda),
                                        time_major=time_major)
      _close(got, want)
    want = ovo.discounted_return(rr, dd, fv, time_major=time_major, provide_all_returns=False)
    got = value_ops.discounted_return(torch.as_tensor(rr, device=cuda), torch.as_tensor(dd, device=cuda),
                                      torch.as_tensor(fv, device=cuda), time_major=time_major,
                                      provide_all_returns=False)
    _close(got, want)


@pytest.mark.parametrize('B,T', [(1, 1), (2, 9), (7, 31), (5, 64), (33, 257), (4096, 128)])
@pytest.mark.parametrize('lam', [0.0, 0.95, 1.0])
def test_gae_parity(cuda, B, T, lam):
  rng = np.random.RandomState(B + T)
  r, d, v, fv = _inputs(rng, B, T)
  for time_major in (False, True):
    args = [x.T.copy() if time_major else x for x in (v, d, r)]
    want = ovo.generalized_advantage_estimation(args[0], fv, args[1], args[2], lam, time_major)
    got = value_ops.generalized_advantage_estimation(
        torch.as_tensor(args[0], device=cuda), torch.as_tensor(fv, device=cuda),
        torch.as_tensor(args[1], device=cuda), torch.as_tensor(args[2], device=cuda), lam, time_major)
    _close(got, want)


def test_reference_goldens_through_cuda(cuda):
  # utils/value_ops_test.py:179-206 and :239-278
  got = value_ops.discounted_return(
      torch.ones(9, device=cuda), torch.tensor([1, 1, 1, 1, 0, .9, .9, .9, .9], device=cuda),
      final_value=torch.tensor(8., device=cuda))
  want = [5, 4, 3, 2, 1, 8 * 0.9**4 + 3.439, 8 * 0.9**3 + 2.71, 8 * 0.9**2 + 1.9, 8 * 0.9 + 1]
  np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-6)
  d = torch.tensor([[1, 1, 1, 1, 0, .9, .9, .9, 0]] * 2, device=cuda)
  adv = value_ops.generalized_advantage_estimation(
      values=torch.full((2, 9), 3., device=cuda), final_value=torch.full((2,), 3., device=cuda),
      discounts=d, rewards=torch.ones(2, 9, device=cuda), td_lambda=0.95, time_major=False)
  want = [2.0808625, 1.13775, 0.145, -0.9, -2.0, 0.56016475, -0.16355, -1.01, -2.0]
  np.testing.assert_allclose(adv.cpu().numpy(), [want, want], rtol=1e-6, atol=1e-6)


def test_linearity_property_full_size(cuda):
  """Size-independent property at config-3 size: returns are linear in the rewards."""
  B, T = 4096, 128
  g = torch.Generator(device='cuda').manual_seed(0)
  r1 = torch.randn(B, T, device=cuda, generator=g)
  r2 = torch.randn(B, T, device=cuda, generator=g)
  d = (torch.rand(B, T, device=cuda, generator=g) > 0.05).float() * 0.99
  f = lambda r: value_ops.discounted_return(r, d, time_major=False)
  lhs, rhs = f(r1 + 2 * r2), f(r1) + 2 * f(r2)
  assert torch.allclose(lhs, rhs, rtol=1e-4, atol=1e-4)
  # where the discount is 0 the return equals the reward
  assert torch.equal(f(r1)[d == 0], r1[d == 0])


@pytest.mark.parametrize('n', [1, 2, 3, 5])
def test_n_step_transition_parity(cuda, n):
  rng = np.random.RandomState(n)
  B, T = 17, n + 1
  rew = rng.randn(B, T).astype(f32)
  disc = (rng.rand(B, T) > 0.2).astype(f32) * f32(0.9)
  gamma = 0.97
  wr, wd = ovo.n_step_reduce(rew, disc, gamma)
  traj = trajectory.Trajectory(
      step_type=torch.zeros(B, T, dtype=torch.int32, device=cuda),
      observation=torch.arange(B * T, dtype=torch.float32, device=cuda).reshape(B, T),
      action=torch.zeros(B, T, dtype=torch.int32, device=cuda), policy_info=(),
      next_step_type=torch.ones(B, T, dtype=torch.int32, device=cuda),
      reward=torch.as_tensor(rew, device=cuda), discount=torch.as_tensor(disc, device=cuda))
  time_steps, policy_steps, next_time_steps = trajectory.to_n_step_transition(traj, gamma)
  np.testing.assert_allclose(next_time_steps.reward.cpu().numpy(), wr, rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(next_time_steps.discount.cpu().numpy(), wd, rtol=1e-6, atol=0)
  assert torch.isnan(time_steps.reward).all() and torch.isnan(time_steps.discount).all()
  assert time_steps.observation.cpu().tolist() == [float(b * T) for b in range(B)]
  assert next_time_steps.observation.cpu().tolist() == [float(b * T + T - 1) for b in range(B)]


def test_n_step_goldens(cuda):  # trajectories/trajectory_test.py:241-317
  g = 0.5
  traj = trajectory.Trajectory(
      step_type=torch.tensor([[0, 1, 1, 2]], dtype=torch.int32, device=cuda),
      observation=torch.tensor([[10., 20., 30., 40.]], device=cuda),
      action=torch.tensor([[11., 22., 33., 44.]], device=cuda),
      policy_info=torch.tensor([[10., 20., 30., 40.]], device=cuda),
      next_step_type=torch.tensor([[1, 1, 2, 0]], dtype=torch.int32, device=cuda),
      reward=torch.tensor([[-1., 1., 2., 0.]], device=cuda),
      discount=torch.tensor([[.9, .95, 1., 0.]], device=cuda))
  time_steps, policy_steps, next_time_steps = trajectory.to_n_step_transition(traj, gamma=g)
  np.testing.assert_allclose(next_time_steps.reward.cpu().numpy(),
                             [-1.0 + 1.0 * g * 0.9 + 2.0 * g**2 * 0.9 * 0.95], rtol=1e-6)
  np.testing.assert_allclose(next_time_steps.discount.cpu().numpy(), [g**2 * 0.9 * 0.95], rtol=1e-6)
  assert next_time_steps.step_type.item() == 2 and next_time_steps.observation.item() == 40.0
  assert policy_steps.action.item() == 11.0 and policy_steps.info.item() == 10.0
  with pytest.raises(ValueError, match='at least 2'):
    trajectory.to_n_step_transition(nest_slice(traj), gamma=g)


def nest_slice(traj):
  from agents_b200.utils import nest
  return nest.map_structure(lambda t: t[:, :1], traj)
