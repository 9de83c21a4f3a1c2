// Episode bookkeeping of the EpisodicReplayBuffer (replay_buffers/episodic_replay_buffer.py).
//
// The reference keeps, per episode slot, an id (`_episodes_loc_to_id_map`), a length, a completed
// flag and a TensorList of unbounded length, and serialises every update through a
// tf.CriticalSection.  Here a slot owns a fixed window of `max_len` rows of the leaf storage
// (row = slot * max_len + position), the counters are device-resident, and ONE single-CTA launch
// does what `_get_batch_episode_ids` (:1109-1187), `_maybe_end_batch_episodes` (:1015-1045) and
// the length bump of `add_batch` / `add_sequence` (:332-463) do: stream order replaces the
// critical section.  The rows it returns are then written with b200rl_rb_write_rows (the same
// bulk-copy kernel as the uniform buffer's add_batch).
#include "common.cuh"

namespace b200rl {

struct EpArgs {
  int64_t* episode_ids;        // [N] in/out
  const uint8_t* begin;        // [N] or NULL (all false)
  const uint8_t* end;          // [N] or NULL
  const uint8_t* mask;         // [N] or NULL (all true): which ids may be renewed
  const int64_t* steps;        // [N] rows to append per item, or NULL -> steps_all
  int64_t steps_all;           // 0: only renew / end episodes (no append)
  int64_t N, capacity, max_len;
  int64_t* last_episode;       // scalar, -1 when empty
  int64_t* loc_to_id;          // [capacity], -1 = never used
  int64_t* lengths;            // [capacity]
  uint8_t* completed;          // [capacity]
  int64_t* num_writes;         // scalar (add_batch bumps it, :446)
  int bump_writes;
  int set_completed_from_end;  // extend_episodes: completed[loc] = end[i] instead of |= (:1381-1385)
  int64_t* out_rows;           // [N] first storage row of the appended steps, or the trash row
  int32_t* overflow;           // set to 1 when an append would exceed max_len (those steps are dropped)
};

// One CTA.  Id renewal is order dependent (new ids are consecutive in item order, :1147-1152), so
// it runs as a block-wide exclusive scan; everything else is per item.
__global__ void __launch_bounds__(1024) ep_assign_kernel(const EpArgs a) {
  pdl_prologue();
  __shared__ int64_t s_scan[1024];
  __shared__ int64_t s_base, s_total;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t trash = a.capacity * a.max_len;
  if (tid == 0) {
    s_base = *a.last_episode + 1;
    s_total = 0;
  }
  __syncthreads();
  // ---- 1. renew ids: (id < 0 | begin) & mask, consecutive new ids in item order -------------
  for (int64_t i0 = 0; i0 < a.N; i0 += nt) {
    const int64_t i = i0 + tid;
    int64_t flag = 0;
    if (i < a.N) {
      const bool upd = (a.episode_ids[i] < 0 || (a.begin && a.begin[i])) && (!a.mask || a.mask[i]);
      flag = upd ? 1 : 0;
    }
    s_scan[tid] = flag;
    __syncthreads();
    for (int off = 1; off < nt; off <<= 1) {          // Hillis-Steele inclusive scan
      const int64_t v = tid >= off ? s_scan[tid - off] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int64_t before = s_total;
    if (i < a.N && flag) {
      const int64_t id = s_base + before + s_scan[tid] - 1;
      const int64_t loc = id % a.capacity;
      a.episode_ids[i] = id;
      a.loc_to_id[loc] = id;                          // the slot's previous episode is gone
      a.completed[loc] = 0;
      a.lengths[loc] = 0;
    }
    __syncthreads();
    if (tid == nt - 1) s_total = before + s_scan[tid];
    __syncthreads();
  }
  if (tid == 0) {
    *a.last_episode = s_base + s_total - 1;
    if (a.bump_writes && a.num_writes) *a.num_writes += 1;
  }
  __syncthreads();
  // ---- 2. end episodes, 3. reserve rows ----------------------------------------------------
  for (int64_t i = tid; i < a.N; i += nt) {
    const int64_t id = a.episode_ids[i];
    const int64_t loc = id >= 0 ? id % a.capacity : 0;
    const bool valid = id >= 0 && a.loc_to_id[loc] == id;
    if (valid && a.end) {
      if (a.set_completed_from_end) a.completed[loc] = a.end[i] ? 1 : 0;
      else if (a.end[i]) a.completed[loc] = 1;
    }
    if (a.out_rows) {
      int64_t row = trash;
      const int64_t n = a.steps ? a.steps[i] : a.steps_all;
      if (valid && n > 0) {
        // items of one launch belong to different episodes (one env step per episode, or one
        // sequence): the read-modify-write of the length needs no atomic
        const int64_t pos = a.lengths[loc];
        if (pos + n <= a.max_len) {
          row = loc * a.max_len + pos;
          a.lengths[loc] = pos + n;
        } else if (a.overflow) {
          *a.overflow = 1;
        }
      }
      a.out_rows[i] = row;
    }
  }
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_ep_assign(int64_t* episode_ids, const uint8_t* begin, const uint8_t* end,
                                const uint8_t* mask, const int64_t* steps, int64_t steps_all,
                                int64_t N, int64_t capacity, int64_t max_len,
                                int64_t* last_episode, int64_t* loc_to_id, int64_t* lengths,
                                uint8_t* completed, int64_t* num_writes, int bump_writes,
                                int set_completed_from_end, int64_t* out_rows, int32_t* overflow,
                                void* stream) {
  B200RL_CHECK_ARG(episode_ids && last_episode && loc_to_id && lengths && completed,
                   "ep_assign: NULL argument");
  B200RL_CHECK_ARG(N >= 0 && capacity >= 1 && max_len >= 1, "ep_assign: bad sizes");
  if (N == 0) return B200RL_OK;
  EpArgs a{episode_ids, begin, end, mask, steps, steps_all, N, capacity, max_len, last_episode,
           loc_to_id, lengths, completed, num_writes, bump_writes, set_completed_from_end,
           out_rows, overflow};
  B200RL_LAUNCH(ep_assign_kernel, 1, 1024, 0, (cudaStream_t)stream, a);
  B200RL_CHECK_LAUNCH("ep_assign");
  return B200RL_OK;
}
