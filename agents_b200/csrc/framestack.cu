// Frame-dedup gather: the ring stores ONE uint8 frame per slot and the sampled observation
// [B, T, H, W, K] is rebuilt from the K most recent frames of the same episode.
//
// The reference only de-duplicates frames in the host-side PyHashedReplayBuffer
// (replay_buffers/py_hashed_replay_buffer.py:37-181); TFUniformReplayBuffer stores the 4-stack
// the Atari wrappers emit (environments/atari_wrappers.py:82-126), i.e. every frame 4 times.
// Semantics (oracle/frame_stack.py): channel c of the item with id i is the stored frame of id
// max(i - (K-1-c), id of the episode's FIRST step) in the same segment.
//
// One CTA per sampled (b, t) row and half frame: thread 0 walks the step_type leaf back (<= K-1
// dependent 4-byte loads), then all threads read 4 pixels of each of the K frames (coalesced 4 B
// loads), transpose bytes in registers and write 4 pixels x K channels (16 B stores for K = 4).
// HBM-bound: reads K * HW and writes K * HW bytes per row, the same traffic as gathering the
// stored stack, with 1/K of the storage.
#include "common.cuh"

namespace b200rl {

template <int K>
__global__ void __launch_bounds__(256) frame_stack_gather_kernel(
    const uint8_t* __restrict__ frames, const int32_t* __restrict__ step_type, int64_t frame_bytes,
    int64_t L, const int64_t* __restrict__ ids, const int64_t* __restrict__ offs, int64_t T,
    uint8_t* __restrict__ out) {
  pdl_prologue();
  __shared__ int64_t src_row[K];
  const int64_t row = blockIdx.x;                    // b * T + t
  const int64_t b = row / T, t = row - b * T;
  if (threadIdx.x == 0) {
    const int64_t base = offs[b] * L;
    int64_t cur = ids[b] + t;
#pragma unroll
    for (int c = K - 1; c >= 0; --c) {               // newest -> oldest, stop at the episode's FIRST
      const int64_t r = base + cur % L;
      src_row[c] = r;
      if (c > 0 && step_type[r] != 0 && cur > 0) cur -= 1;
    }
  }
  __syncthreads();
  const uint32_t* f[K];
#pragma unroll
  for (int c = 0; c < K; ++c)
    f[c] = reinterpret_cast<const uint32_t*>(frames + src_row[c] * frame_bytes);
  uint8_t* o = out + row * frame_bytes * K;
  const int64_t nw = frame_bytes >> 2;               // 4 pixels per 32-bit word
  for (int64_t p = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; p < nw;
       p += (int64_t)gridDim.y * blockDim.x) {
    uint32_t w[K];
#pragma unroll
    for (int c = 0; c < K; ++c) w[c] = __ldg(f[c] + p);
    if (K == 4) {
      uint4 v;
      v.x = (w[0] & 0xFFu) | ((w[1] & 0xFFu) << 8) | ((w[2] & 0xFFu) << 16) | ((w[3] & 0xFFu) << 24);
      v.y = ((w[0] >> 8) & 0xFFu) | (((w[1] >> 8) & 0xFFu) << 8) | (((w[2] >> 8) & 0xFFu) << 16) |
            (((w[3] >> 8) & 0xFFu) << 24);
      v.z = ((w[0] >> 16) & 0xFFu) | (((w[1] >> 16) & 0xFFu) << 8) | (((w[2] >> 16) & 0xFFu) << 16) |
            (((w[3] >> 16) & 0xFFu) << 24);
      v.w = (w[0] >> 24) | ((w[1] >> 24) << 8) | ((w[2] >> 24) << 16) | ((w[3] >> 24) << 24);
      *reinterpret_cast<uint4*>(o + p * 16) = v;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < K; ++c) o[(p * 4 + j) * K + c] = (uint8_t)((w[c] >> (8 * j)) & 0xFFu);
    }
  }
}

}  // namespace b200rl

extern "C" int b200rl_rb_gather_frame_stack(const void* frames, const int32_t* step_type,
                                            int64_t frame_bytes, int64_t max_length,
                                            const int64_t* ids_dev, const int64_t* offsets_dev,
                                            int64_t B, int64_t T, int32_t K, void* out,
                                            void* stream) {
  using namespace b200rl;
  B200RL_CHECK_ARG(frames && step_type && ids_dev && offsets_dev && out,
                   "gather_frame_stack: NULL argument");
  B200RL_CHECK_ARG(B >= 0 && T >= 1 && max_length >= 1, "gather_frame_stack: bad sizes");
  B200RL_CHECK_ARG(K >= 1 && K <= 4, "gather_frame_stack: stack depth must be 1..4, got %d", K);
  B200RL_CHECK_ARG(frame_bytes > 0 && (frame_bytes & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(frames) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   "gather_frame_stack: frames must be a multiple of 4 bytes and 4 B / 16 B aligned");
  if (B == 0) return B200RL_OK;
  const dim3 grid((unsigned)(B * T), 2);
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* fr = (const uint8_t*)frames;
  uint8_t* o = (uint8_t*)out;
  switch (K) {
    case 1: B200RL_LAUNCH(frame_stack_gather_kernel<1>, grid, 256, 0, st, fr, step_type, frame_bytes, max_length, ids_dev, offsets_dev, T, o); break;
    case 2: B200RL_LAUNCH(frame_stack_gather_kernel<2>, grid, 256, 0, st, fr, step_type, frame_bytes, max_length, ids_dev, offsets_dev, T, o); break;
    case 3: B200RL_LAUNCH(frame_stack_gather_kernel<3>, grid, 256, 0, st, fr, step_type, frame_bytes, max_length, ids_dev, offsets_dev, T, o); break;
    default: B200RL_LAUNCH(frame_stack_gather_kernel<4>, grid, 256, 0, st, fr, step_type, frame_bytes, max_length, ids_dev, offsets_dev, T, o); break;
  }
  B200RL_CHECK_LAUNCH("gather_frame_stack");
  return B200RL_OK;
}
