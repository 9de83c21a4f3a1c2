"""CPU restatement of PPOLearner's data pipeline and run loop (TEST INFRASTRUCTURE ONLY).

Follows tf_agents/train/ppo_learner.py:
  * `_create_datasets` (:220-250): `take(num_samples).cache().repeat(num_epochs)`, then for
    minibatching `map(BatchSquash(2).flatten).unbatch().shuffle(S).batch(1).batch(mb,
    drop_remainder=True)`;
  * `run` (:270-304): update normalisers over `num_samples` batches, then
    `int(num_frames / mb) * num_epochs / num_replicas` train iterations (or
    `num_samples * num_epochs / num_replicas` without minibatching);
  * `_update_normalizers` (:306-337).
Pinned against the reference's own expectations (train/ppo_learner_test.py:193-376: train-call
counts 10/20/60/1/2/2, 12/24/72/3/6/6, 48 and the exact minibatch contents with
shuffle_buffer_size=1) in tests/test_ppo_learner_host.py.  The shuffle order for buffer > 1 is
unpinned in the reference (tf.data RNG); ours is defined in include/b200rl.h
(b200rl_shuffle_order) and restated here in pure Python.
"""
import numpy as np

from oracle import philox


def shuffle_order(n, buffer, seed, call):
  """Emission order of shuffle(buffer) over a stream of n elements (include/b200rl.h)."""
  cap = min(int(buffer), int(n))
  slots = list(range(cap))
  nxt, fill = cap, cap
  out = np.empty(n, np.int64)
  for i in range(n):
    r = philox.philox(i, call, seed)
    j = int(philox.uniform_i64(r[0], r[1], 0, fill))
    out[i] = slots[j]
    if nxt < n:
      slots[j] = nxt
      nxt += 1
    else:
      fill -= 1
      slots[j] = slots[fill]
  return out


def minibatch_rows(num_frames, num_epochs, minibatch_size, shuffle_buffer_size, seed, call):
  """Row indices (into the flattened [num_frames] cache) of every minibatch one pass of the
  inner dataset yields: stream element i is cache element i % num_frames (:226), the shuffled
  stream is cut into minibatches and the remainder is dropped (:239-242)."""
  n = num_frames * num_epochs
  order = shuffle_order(n, shuffle_buffer_size, seed, call) % num_frames
  nb = n // minibatch_size
  return order[:nb * minibatch_size].reshape(nb, minibatch_size)


def iterations_per_run(num_frames, num_samples, num_epochs, minibatch_size, num_replicas):
  """`run` (:283-300)."""
  if minibatch_size:
    total = int(num_frames / minibatch_size) * num_epochs
  else:
    total = num_samples * num_epochs
  it = int(total / num_replicas)
  if it == 0:
    raise ValueError('Cannot distribute {} batches across {} replicas.'.format(total, num_replicas))
  return it
