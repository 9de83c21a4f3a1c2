// Tensor-core GEMM for the Q / actor / critic networks: tcgen05.mma (kind::tf32) with TMEM
// accumulators, fp32-grade accuracy through the 3xTF32 split.
//
//   C[M,N] (+)= act(A[M,K] @ B[K,N] + bias)      A, B given as operand views (see nn.cu)
//
// Why 3xTF32: north_star parity is 1e-5 relative on the loss; one TF32 pass keeps ~11 mantissa
// bits (1e-3).  Each fp32 operand x is split in registers into hi = x with the low 13 mantissa
// bits cleared and lo = x - hi (exact); the tensor core accumulates lo_a*hi_b + hi_a*lo_b +
// hi_a*hi_b into the same fp32 TMEM accumulator (the dropped lo*lo term is 2^-22 relative).
//
// Structure (one CTA = one 128 x BN output tile, cta_group::1):
//   warps 0-3  producers: gather a (128 x 32) A tile and a (BN x 32) B^T tile of fp32 through
//              the operand views (implicit im2col, u8->f32 cast, transposes all happen here),
//              split hi/lo, store them K-major into 128B-swizzled shared memory (the layout a
//              TMA SWIZZLE_128B load would produce), fence.proxy.async, arrive on full[s];
//              after the main loop the same 4 warps are the epilogue (tcgen05.ld 32x32b:
//              warp w owns TMEM lanes 32w..32w+31), apply bias/activation, store C or the
//              split-K partial.
//   warp 4     allocates TMEM (BN fp32 columns), lane 0 issues tcgen05.mma for every K step
//              (K = 8 per instruction for tf32; 4 steps x 3 passes per 32-wide K block),
//              releases stages with tcgen05.commit -> empty[s], signals the epilogue with a
//              final commit.
// Descriptors follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor: K-major SWIZZLE_128B,
// SBO = 1024 B, version 1; InstrDescriptor: F32 accumulate, TF32 x TF32, M = 128, N = BN).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rl {
namespace tc {

constexpr int kBM = 128;
constexpr int kBK = 32;                    // fp32 elements = 128 bytes = one swizzle row
constexpr int kProducerThreads = 128;
constexpr int kThreads = 160;

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset
  d |= (uint64_t)1 << 46;                          // version = 1 (sm_100)
  d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// byte offset of 16-byte chunk j (0..7) of row r inside a [rows x 128 B] swizzled tile
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t j) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((j ^ (r & 7u)) << 4);
}

__device__ __forceinline__ void split_store(unsigned char* hi_tile, unsigned char* lo_tile,
                                            uint32_t off, float4 v, bool with_lo) {
  float4 h;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  if (with_lo) {
    float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

template <int BN, int STAGES, int PASSES>
struct SmemLayout {
  static constexpr int kATile = kBM * 128;
  static constexpr int kBTile = BN * 128;
  static constexpr int kNumA = PASSES == 3 ? 2 : 1;
  static constexpr int kStage = kNumA * (kATile + kBTile);
  static constexpr int kBytes = STAGES * kStage + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// AL::at(m, k) / BL::at(k, n) are the fp32 operand views of nn.cu.
template <int BN, int STAGES, int PASSES, class AL, class BL>
__global__ void __launch_bounds__(kThreads) tc_gemm_kernel(const AL a, const BL b,
                                                           float* __restrict__ C,
                                                           const float* __restrict__ bias,
                                                           int64_t M, int64_t N, int64_t K, int act,
                                                           int beta, int splits,
                                                           int64_t k_per_split,
                                                           float* __restrict__ ws) {
  using L = SmemLayout<BN, STAGES, PASSES>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* bars = smem + STAGES * L::kStage;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(bars);
  unsigned long long* empty = full + STAGES;
  unsigned long long* accum = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.y * kBM, n0 = (int64_t)blockIdx.x * BN;
  const int split = blockIdx.z;
  const int64_t kb = (int64_t)split * k_per_split;
  const int64_t ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  const int nkb = (int)((ke - kb + kBK - 1) / kBK);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_addr(&full[s]), 4);    // one arrive per producer warp
      mbar_init(smem_addr(&empty[s]), 1);   // tcgen05.commit
    }
    mbar_init(smem_addr(accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    constexpr int kCols = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(tmem_slot)),
                 "n"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== producers =====================
    for (int kbi = 0; kbi < nkb; ++kbi) {
      const int s = kbi % STAGES;
      const uint32_t ph = (uint32_t)((kbi / STAGES) & 1);
      if (kbi >= STAGES) mbar_wait(smem_addr(&empty[s]), ph ^ 1u);
      unsigned char* st = smem + s * L::kStage;
      unsigned char* a_hi = st;
      unsigned char* a_lo = st + L::kATile;
      unsigned char* b_hi = st + L::kNumA * L::kATile;
      unsigned char* b_lo = b_hi + L::kBTile;
      const int64_t k0 = kb + (int64_t)kbi * kBK;
      // A tile: 128 rows x 8 chunks; 8 consecutive threads cover one row (128 contiguous bytes
      // for K-contiguous views)
#pragma unroll 4
      for (int i = tid; i < kBM * 8; i += kProducerThreads) {
        const int r = i >> 3, j = i & 7;
        const int64_t m = m0 + r, k = k0 + j * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M) {
          if (k + 0 < ke) v.x = a.at(m, k + 0);
          if (k + 1 < ke) v.y = a.at(m, k + 1);
          if (k + 2 < ke) v.z = a.at(m, k + 2);
          if (k + 3 < ke) v.w = a.at(m, k + 3);
        }
        split_store(a_hi, a_lo, sw128(r, j), v, PASSES == 3);
      }
#pragma unroll 2
      for (int i = tid; i < BN * 8; i += kProducerThreads) {
        const int r = i >> 3, j = i & 7;
        const int64_t n = n0 + r, k = k0 + j * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N) {
          if (k + 0 < ke) v.x = b.at(k + 0, n);
          if (k + 1 < ke) v.y = b.at(k + 1, n);
          if (k + 2 < ke) v.z = b.at(k + 2, n);
          if (k + 3 < ke) v.w = b.at(k + 3, n);
        }
        split_store(b_hi, b_lo, sw128(r, j), v, PASSES == 3);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&full[s]));
    }
    // ===================== epilogue =====================
    mbar_wait(smem_addr(accum), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t m = m0 + warp * 32 + lane;
    float* out = (splits > 1) ? ws + (int64_t)split * M * N : C;
#pragma unroll
    for (int c = 0; c < BN; c += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
            "=r"(r[7])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < M && nkb > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t n = n0 + c + j;
          if (n >= N) continue;
          float v = __uint_as_float(r[j]);
          if (splits == 1) {
            if (bias) v += bias[n];
            v = apply_act(v, act);
            if (beta) v += out[m * N + n];
          }
          out[m * N + n] = v;
        }
      } else if (m < M && nkb == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t n = n0 + c + j;
          if (n < N) out[m * N + n] = 0.f;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kbi = 0; kbi < nkb; ++kbi) {
        const int s = kbi % STAGES;
        const uint32_t ph = (uint32_t)((kbi / STAGES) & 1);
        mbar_wait(smem_addr(&full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        unsigned char* st = smem + s * L::kStage;
        const uint32_t a_hi = smem_addr(st), a_lo = a_hi + L::kATile;
        const uint32_t b_hi = a_hi + L::kNumA * L::kATile, b_lo = b_hi + L::kBTile;
#pragma unroll
        for (int ks = 0; ks < kBK / 8; ++ks) {
          const uint32_t koff = (uint32_t)ks * 32u;  // 8 tf32 = 32 bytes inside the 128 B row
          const uint32_t first = (kbi == 0 && ks == 0) ? 0u : 1u;
          if (PASSES == 3) {
            tc_mma_tf32(tmem_base, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, first);
            tc_mma_tf32(tmem_base, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1u);
            tc_mma_tf32(tmem_base, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, 1u);
          } else {
            tc_mma_tf32(tmem_base, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
          }
        }
        tc_commit(smem_addr(&empty[s]));
      }
      tc_commit(smem_addr(accum));
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    constexpr int kCols = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(kCols));
  }
}

}  // namespace tc
}  // namespace b200rl
