"""Late-sorting GPU tests added after the round-1 GPU budget was spent (written and traced on CPU,
first executed by the round-end run): reference PPOLearner cases through the real minibatch
pipeline, and PyDriver -> PyTFEagerPolicy -> PinnedAddBatch -> ring.  Kept in a file that sorts
last so that `pytest -x` reaches every validated test first."""
import numpy as np
import pytest
import torch

from agents_b200.train import ppo_learner
from agents_b200.trajectories import trajectory

pytestmark = pytest.mark.gpu
f32 = np.float32


# ---- train/ppo_learner_test.py replayed through the real pipeline ---------------------------------
class _FakePPOAgent(object):
  """ppo_learner_test.py:41-79 FakePPOAgent: counts train() calls and keeps what it was fed."""

  def __init__(self, device):
    from agents_b200.agents import tf_agent
    self._loss_info = tf_agent.LossInfo(torch.zeros((), device=device), ())
    self._compute_value_and_advantage_in_train = False
    self.update_normalizers_in_train = False
    self.train_step_counter = torch.zeros((), dtype=torch.int64, device=device)
    self._train_step_host = 0
    self.experiences = []

  def train(self, experience, weights=None):
    self.experiences.append(experience)
    self._train_step_host += 1
    return self._loss_info

  def update_observation_normalizer(self, batched_observations):
    pass

  def update_reward_normalizer(self, batched_rewards):
    pass


def _reference_trajectories(cuda, n_time_steps, batch_size):
  """ppo_learner_test.py:82-130 `_create_trajectories`: obs[b, t, 0] = 10 b + t, the rest ones."""
  obs = torch.as_tensor(np.asarray([np.arange(n_time_steps) + 10 * i for i in range(batch_size)], f32)[..., None],
                        device=cuda)
  ones = torch.ones(batch_size, n_time_steps, device=cuda)
  mid = torch.ones(batch_size, n_time_steps, dtype=torch.int32, device=cuda)
  info = {'dist_params': {'loc': ones[..., None].clone(), 'scale': ones[..., None].clone()},
          'value_prediction': ones.clone(), 'return': ones.clone(), 'advantage': ones.clone()}
  return trajectory.Trajectory(mid, obs, ones[..., None].clone(), info, mid.clone(), ones.clone(), ones.clone())


@pytest.mark.parametrize('episodes,steps,num_epochs,envs,mb,expected', [
    (1, 100, 1, 1, 10, 10), (1, 100, 2, 1, 10, 20), (1, 100, 2, 3, 10, 60),       # :193-253
    (1, 100, 1, 1, None, 1), (1, 100, 2, 3, None, 2),
    (3, 40, 1, 1, 10, 12), (3, 40, 2, 3, 10, 72), (3, 40, 2, 1, None, 6),          # :255-329
    (3, 40, 4, 1, 10, 48)])                                                          # :331-376
def test_reference_ppo_learner_cases(cuda, tmp_path, episodes, steps, num_epochs, envs, mb, expected):
  traj = _reference_trajectories(cuda, steps, envs)
  dataset_fn = lambda: [(traj, ())] * episodes
  agent = _FakePPOAgent(cuda)
  lrn = ppo_learner.PPOLearner(str(tmp_path), agent.train_step_counter, agent, dataset_fn, dataset_fn,
                               num_samples=episodes, num_epochs=num_epochs, minibatch_size=mb,
                               shuffle_buffer_size=1, checkpoint_interval=0)    # buffer 1 = no shuffling
  loss = lrn.run()
  assert len(agent.experiences) == expected and float(loss.loss.item()) == 0.0
  want_obs = traj.observation.reshape(-1).cpu().numpy()
  if mb:
    stream = np.concatenate([want_obs] * (episodes * num_epochs))   # _concat_and_flatten (:133-152)
    for i, got in enumerate(agent.experiences):                     # _get_expected_minibatch (:155-178)
      assert tuple(got.observation.shape) == (mb, 1, 1) and tuple(got.reward.shape) == (mb, 1)
      np.testing.assert_array_equal(got.observation.reshape(-1).cpu().numpy(), stream[mb * i:mb * (i + 1)])
      assert set(got.policy_info) == {'dist_params', 'value_prediction', 'return', 'advantage'}
  else:
    for got in agent.experiences:
      assert got is traj


def test_py_driver_feeds_the_ring(cuda):
  """drivers/py_driver.py host loop with a device policy behind PyTFEagerPolicy and the pinned
  add_batch observer: ring contents equal the reference's driver golden
  (dynamic_step_driver_test.py:121-166)."""
  from agents_b200.drivers import py_driver
  from agents_b200.environments import batched_py_environment
  from agents_b200.policies import py_tf_eager_policy
  from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
  from py_env_mocks import PyEnvironmentMock
  from test_driver_gpu import PolicyMock
  env = batched_py_environment.BatchedPyEnvironment([PyEnvironmentMock()], multithreading=False)
  device_policy = PolicyMock(env.time_step_spec(), env.action_spec(), cuda)
  policy = py_tf_eager_policy.PyTFEagerPolicy(device_policy, device=cuda)
  rb = rb_mod.TFUniformReplayBuffer(device_policy.trajectory_spec, batch_size=1, max_length=1000, device=cuda)
  driver = py_driver.PyDriver(env, policy, observers=[py_driver.PinnedAddBatch(rb)], max_steps=6)
  driver.run(env.reset(), policy.get_initial_state(1))
  tr = rb.gather_all()
  assert tr.step_type.cpu().tolist() == [[0, 1, 2, 0, 1, 2, 0, 1]]
  assert tr.observation.cpu().tolist() == [[0, 1, 3, 0, 1, 3, 0, 1]]
  assert tr.action.cpu().tolist() == [[1, 2, 1, 1, 2, 1, 1, 2]]
  assert tr.policy_info.cpu().tolist() == [[2, 4, 2, 2, 4, 2, 2, 4]]
  assert tr.next_step_type.cpu().tolist() == [[1, 2, 0, 1, 2, 0, 1, 2]]
  assert tr.reward.cpu().tolist() == [[1., 1., 0., 1., 1., 0., 1., 1.]]
  assert tr.discount.cpu().tolist() == [[1., 0., 1., 1., 0., 1., 1., 0.]]
  env.close()


@pytest.mark.parametrize('K,L,adds', [(4, 16, 11), (4, 8, 29), (3, 8, 20), (2, 8, 9)])
def test_frame_stack_buffer_rebuilds_the_stored_stacks(cuda, K, L, adds):
  """FrameStackReplayBuffer (one frame per slot) returns, for the ids it samples, exactly the
  stacks a FrameStack-K producer emitted (oracle/frame_stack.py), across episode starts and after
  the ring wrapped; the other leaves come back as in TFUniformReplayBuffer."""
  from agents_b200.replay_buffers import frame_stack_replay_buffer as fsrb
  from agents_b200.specs import tensor_spec
  from oracle import frame_stack as ofs
  rng = np.random.RandomState(K * 100 + adds)
  B_env, H, W = 3, 6, 6
  frames = rng.randint(0, 256, size=(B_env, adds, H, W)).astype(np.uint8)
  step_types = rng.choice([0, 1, 1, 1, 2], size=(B_env, adds)).astype(np.int32)
  step_types[:, 0] = 0
  for b in range(B_env):
    for t in range(1, adds):
      if step_types[b, t - 1] == 2:
        step_types[b, t] = 0
  stacks = np.stack([ofs.stack_rule(frames[b], step_types[b], K) for b in range(B_env)])   # [B_env, adds, H, W, K]
  spec = trajectory.Trajectory(
      step_type=tensor_spec.TensorSpec([], torch.int32, 'step_type'),
      observation=tensor_spec.TensorSpec((H, W, K), torch.uint8, 'observation'),
      action=tensor_spec.TensorSpec([], torch.int32, 'action'), policy_info=(),
      next_step_type=tensor_spec.TensorSpec([], torch.int32, 'next_step_type'),
      reward=tensor_spec.TensorSpec([], torch.float32, 'reward'),
      discount=tensor_spec.TensorSpec([], torch.float32, 'discount'))
  rb = fsrb.FrameStackReplayBuffer(spec, batch_size=B_env, max_length=L, device=cuda, seed=5)
  assert rb.stack_depth == K
  d = lambda a: torch.as_tensor(a, device=cuda)
  for t in range(adds):
    obs = d(stacks[:, t]) if t % 2 == 0 else d(frames[:, t])       # stacked and single-frame producers
    rb.add_batch(trajectory.Trajectory(
        d(step_types[:, t]), obs, d(np.arange(B_env, dtype=np.int32)), (), d(step_types[:, t]),
        d(np.full(B_env, t, f32)), d(np.ones(B_env, f32))))
  oldest = max(0, adds - L)
  for B, T in [(8, 1), (5, 2), (16, 3)]:
    data, info = rb.get_next(sample_batch_size=B, num_steps=T)
    ids = info.ids.cpu().numpy()
    env = data.action.cpu().numpy()                                  # action leaf = segment index
    got = data.observation.cpu().numpy()
    assert got.shape == (B, T, H, W, K) and ids.shape == (B, T)
    lo, hi = ofs.valid_range_ids(adds - 1, L, T, K)
    assert (ids[:, 0] >= lo).all() and (ids[:, 0] < hi).all() and (ids[:, -1] <= adds - 1).all()
    if adds <= L:
      assert lo == 0                                                  # nothing excluded before the wrap
    np.testing.assert_allclose(info.probabilities.cpu().numpy(), 1.0 / ((hi - lo) * B_env), rtol=1e-6)
    np.testing.assert_array_equal(data.reward.cpu().numpy(), ids.astype(f32))   # reward leaf = id
    for b in range(B):
      for t in range(T):
        np.testing.assert_array_equal(got[b, t], stacks[env[b, t], ids[b, t]])
  one, info1 = rb.get_next()                                           # unbatched, single step
  assert tuple(one.observation.shape) == (H, W, K) and info1.ids.dim() == 0
