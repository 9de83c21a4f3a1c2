"""EpisodicReplayBuffer on the GPU: cases of replay_buffers/episodic_replay_buffer_test.py replayed
through the product class (ids, stale-id protection, over-capacity eviction, gather_all,
add_sequence / _get_episode, extract / extend, datasets)."""
import numpy as np
import pytest
import torch

from agents_b200.replay_buffers import episodic_replay_buffer as erb
from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
from agents_b200.specs import tensor_spec

pytestmark = pytest.mark.gpu


def _nested_spec():
  return (tensor_spec.TensorSpec([3], torch.float32, 'action'),
          (tensor_spec.TensorSpec([5], torch.float32, 'lidar'),
           tensor_spec.TensorSpec([3, 2], torch.float32, 'camera')))


def test_create_episode_ids(cuda):  # :94-144
  rb = erb.EpisodicReplayBuffer(_nested_spec(), capacity=2, device=cuda)
  assert rb.create_episode_ids().item() == -1
  assert rb.create_episode_ids(2).cpu().tolist() == [-1, -1]
  with pytest.raises(ValueError):
    rb.create_episode_ids(3)


def test_get_episode_id(cuda):  # :146-185
  rb = erb.EpisodicReplayBuffer(_nested_spec(), capacity=2, device=cuda)
  e0 = rb._get_episode_id(rb.create_episode_ids(), begin_episode=True)
  e1 = rb._get_episode_id(e0, begin_episode=False)
  e2 = rb._get_episode_id(e1, begin_episode=True)
  assert (e0.item(), e1.item(), e2.item()) == (0, 0, 1)


def test_get_batch_episode_ids(cuda):  # :187-249
  rb = erb.EpisodicReplayBuffer(_nested_spec(), capacity=5, device=cuda)
  ids = rb.create_episode_ids(num_episodes=3)
  got = [ids.cpu().tolist()]
  for begin in (False, False, [False, False, False], [True, False, False], [False, True, False],
                [False, True, True]):
    ids = rb._get_batch_episode_ids(ids, begin_episode=begin)
    got.append(ids.cpu().tolist())
  assert got == [[-1, -1, -1], [0, 1, 2], [0, 1, 2], [0, 1, 2], [3, 1, 2], [3, 4, 2], [3, 5, 6]]


def test_maybe_end_episode(cuda):  # :340-385
  rb = erb.EpisodicReplayBuffer(_nested_spec(), capacity=2, device=cuda)
  e0 = rb._get_episode_id(rb.create_episode_ids(), begin_episode=True)
  assert not rb._maybe_end_episode(e0, end_episode=False).item()
  assert rb._maybe_end_episode(e0, end_episode=True).item()
  assert rb._completed_episodes().cpu().tolist() == [0]
  # a new episode in the same slot resets the flag
  e1 = rb._get_episode_id(e0, begin_episode=True)
  e2 = rb._get_episode_id(e1, begin_episode=True)          # id 2 -> slot 0
  assert e2.item() == 2 and rb._completed_episodes().cpu().tolist() == []


def _scalar_rb(cuda, capacity, begin, end, **kw):
  return erb.EpisodicReplayBuffer(tensor_spec.TensorSpec([], torch.int32, 'action'), capacity=capacity,
                                  begin_episode_fn=lambda _: begin, end_episode_fn=lambda _: end,
                                  device=cuda, **kw)


def test_stateful_add_batch(cuda):  # :1426-1458
  spec = tensor_spec.TensorSpec([3], torch.int32, 'lidar')
  rb = erb.EpisodicReplayBuffer(spec, capacity=3, begin_episode_fn=lambda _: False,
                                end_episode_fn=lambda _: False, device=cuda)
  srb = erb.StatefulEpisodicReplayBuffer(rb, num_episodes=3)
  values = np.stack([np.ones(3, np.int32), 10 * np.ones(3, np.int32), 100 * np.ones(3, np.int32)])
  new_ids = srb.add_batch(torch.as_tensor(values, device=cuda))
  assert new_ids.cpu().tolist() == [0, 1, 2] and rb._get_last_episode_id() == 2
  items = torch.cat([rb._get_episode(i) for i in range(3)], 0)
  np.testing.assert_array_equal(items.cpu().numpy(), values)
  assert int(rb.num_frames()) == 3


def test_stateful_gather_all(cuda):  # :1532-1561
  rb = _scalar_rb(cuda, 1000, False, False, max_episode_length=16)
  srb = erb.StatefulEpisodicReplayBuffer(rb, num_episodes=1)
  for i in range(10):
    srb.add_batch(torch.tensor([i], dtype=torch.int32, device=cuda))
  assert rb.gather_all().cpu().tolist() == [list(range(10))]


@pytest.mark.parametrize('check_ids', [False, True])
def test_add_over_capacity_overwrites_old_episodes(cuda, check_ids):  # :1563-1637
  rb = _scalar_rb(cuda, 3, True, False, max_episode_length=4)
  srb = erb.StatefulEpisodicReplayBuffer(rb, num_episodes=1)
  for i in range(5):
    srb.add_batch(torch.tensor([i], dtype=torch.int32, device=cuda))
    if check_ids:
      assert srb.episode_ids.cpu().tolist() == [i]
  assert rb.gather_all()[0].cpu().tolist() == [2, 3, 4]


def test_add_to_stale_episode_id_is_avoided(cuda):  # :1639-1680
  rb = _scalar_rb(cuda, 1, False, False, max_episode_length=4)
  s0 = erb.StatefulEpisodicReplayBuffer(rb, num_episodes=1)
  s1 = erb.StatefulEpisodicReplayBuffer(rb, num_episodes=1)
  t = lambda v: torch.tensor([v], dtype=torch.int32, device=cuda)
  s0.add_batch(t(0))
  assert s0.episode_ids.cpu().tolist() == [0] and rb.gather_all()[0].cpu().tolist() == [0]
  s1.add_batch(t(1))
  assert s1.episode_ids.cpu().tolist() == [1] and rb.gather_all()[0].cpu().tolist() == [1]
  s0.add_batch(t(2))                                   # episode 0 is gone: the step is dropped
  assert s0.episode_ids.cpu().tolist() == [0] and rb.gather_all()[0].cpu().tolist() == [1]


def test_add_sequence_get_episode_and_num_frames(cuda):  # :613-684, :502-541
  spec = (tensor_spec.TensorSpec([2], torch.float32, 'obs'), tensor_spec.TensorSpec([], torch.int64, 'a'))
  rb = erb.EpisodicReplayBuffer(spec, capacity=4, begin_episode_fn=lambda _: False,
                                end_episode_fn=lambda _: False, device=cuda, max_episode_length=32)
  eid = rb.create_episode_ids()
  seq1 = (torch.arange(10, dtype=torch.float32, device=cuda).reshape(5, 2), torch.arange(5, device=cuda))
  seq2 = (torch.arange(10, 16, dtype=torch.float32, device=cuda).reshape(3, 2), torch.arange(5, 8, device=cuda))
  eid = rb.add_sequence(seq1, eid)
  eid = rb.add_sequence(seq2, eid)
  assert eid.item() == 0 and int(rb.num_frames()) == 8
  obs, a = rb._get_episode(eid)
  np.testing.assert_array_equal(obs.cpu().numpy(), np.arange(16, dtype=np.float32).reshape(8, 2))
  assert a.cpu().tolist() == list(range(8))
  other = rb.add_sequence(seq2, rb.create_episode_ids())
  assert other.item() == 1 and int(rb.num_frames()) == 11
  with pytest.raises(rb_mod.InvalidArgumentError):
    rb._get_episode(7)
  data, info = rb.get_next()
  assert info.ids.item() in (0, 1) and data[0].shape[0] in (8, 3)
  # overflow: steps beyond max_episode_length are dropped and flagged
  assert not rb.overflowed()
  big = (torch.zeros(40, 2, device=cuda), torch.zeros(40, dtype=torch.int64, device=cuda))
  rb.add_sequence(big, other)
  assert rb.overflowed() and int(rb.num_frames()) == 11


def test_get_next_empty_raises(cuda):  # :543-564
  rb = _scalar_rb(cuda, 2, False, False)
  with pytest.raises(rb_mod.InvalidArgumentError):
    rb.get_next()
  assert tuple(rb.gather_all().shape) == (0,)           # :801-811


def test_extract_and_extend(cuda):  # :838-1022
  rb = _scalar_rb(cuda, 4, False, False, max_episode_length=8)
  ids = rb.create_episode_ids(2)
  for step in range(3):
    ids = rb.add_batch(torch.tensor([step, 10 + step], dtype=torch.int32, device=cuda), ids)
  ep = rb.extract([0, 1])
  assert ep.length.cpu().tolist() == [3, 3] and ep.completed.cpu().tolist() == [0, 0]
  assert ep.tensor_lists[:, :3].cpu().tolist() == [[0, 1, 2], [10, 11, 12]]
  # extend a second buffer with the extracted episodes
  rb2 = _scalar_rb(cuda, 4, False, False, max_episode_length=8)
  ids2 = rb2.extend_episodes(rb2.create_episode_ids(3), [0, 2], ep)
  assert ids2.cpu().tolist() == [0, -1, 1]
  assert rb2._get_episode(0).cpu().tolist() == [0, 1, 2] and rb2._get_episode(1).cpu().tolist() == [10, 11, 12]
  ids2 = rb2.extend_episodes(ids2, [2], erb.Episodes(length=torch.tensor([2]), completed=torch.tensor([1]),
                                                     tensor_lists=torch.tensor([[7, 8, 0, 0, 0, 0, 0, 0]], dtype=torch.int32)))
  assert rb2._get_episode(1).cpu().tolist() == [10, 11, 12, 7, 8]
  assert rb2._completed_episodes().cpu().tolist() == [1]
  cleared = rb.extract([0], clear_data=True)
  assert cleared.length.cpu().tolist() == [3] and int(rb.num_frames()) == 3


def test_datasets(cuda):  # :1104-1267 (single deterministic pass), :509-691 (random slices)
  rb = _scalar_rb(cuda, 8, False, False, max_episode_length=16)
  ids = rb.create_episode_ids(2)
  for step in range(5):
    ids = rb.add_batch(torch.tensor([step, 100 + step], dtype=torch.int32, device=cuda), ids)
  eps = [e.cpu().tolist() for e in rb.as_dataset(single_deterministic_pass=True)]
  assert eps == [[0, 1, 2, 3, 4], [100, 101, 102, 103, 104]]
  win = [w.cpu().tolist() for w in rb.as_dataset(num_steps=3, single_deterministic_pass=True)]
  assert win == [[0, 1, 2], [3, 4, 100], [101, 102, 103], [104]]
  ds = rb.as_dataset(num_steps=2, sample_batch_size=4)
  data, info = next(ds)
  assert tuple(data.shape) == (4, 2) and tuple(info.ids.shape) == (4,)
  d = data.cpu().numpy()
  assert ((d[:, 1] - d[:, 0]) == 1).all()                 # contiguous slices of one episode
  with pytest.raises(ValueError):
    next(rb.as_dataset(sample_batch_size=2))


def test_clear(cuda):  # :1024-1102
  rb = _scalar_rb(cuda, 3, False, False, max_episode_length=4)
  ids = rb.add_batch(torch.tensor([1, 2], dtype=torch.int32, device=cuda), rb.create_episode_ids(2))
  rb.clear()
  assert int(rb.num_frames()) == 0 and rb._get_last_episode_id() == 1
  ids = rb.add_batch(torch.tensor([3, 4], dtype=torch.int32, device=cuda), ids)   # ids in flight survive
  assert ids.cpu().tolist() == [0, 1] and rb.gather_all()[0].cpu().tolist() == [3, 4]
  rb.clear(clear_all_variables=True)
  assert rb._get_last_episode_id() == -1
