// tcgen05 cross-attention + distance kernel for tuple counts N <= 128 (T=16 pairs: N=120).
//
// Reference semantics (modules/ar/utils/model.py:95-135), per (query window b, class c):
//   S = Kq.Kc^T / sqrt(D);  P = softmax(S, dim=-2)  (over the QUERY-tuple axis, per support tuple);
//   proto = P.Vc;  logit = -||Vq - proto||_F^2 / N.
//
// Mapping to the hardware (one persistent CTA per SM, 12 warps, warp-specialised):
//   MMA1  S^T[s,q]   = Kc[s,:] . Kq'[q,:]      M=128 (support tuples on TMEM lanes), N=128, K=128
//         Kq' is pre-scaled by log2(e)/sqrt(D) so that exp(S) = exp2(S^T) with no multiply.
//   softmax warps: thread == TMEM lane == support tuple s; the normaliser over the query axis is a
//         thread-local sum over the 128 columns.  P^T[s,:] (fp16) goes to shared memory as the B operand.
//   MMA2  proto^T[d,q] = Vc^T[d,:] . P[q,:]     M=128 (d on TMEM lanes), N=128, K=128 (support tuples)
//   epilogue warps: thread == lane d; Vq[q][d] = a[i][d] + b[j][d] is rebuilt from the per-frame V
//         projections held in registers (T=16 pair specialisation) -- tuple features never exist in HBM.
// fp16 operands, fp32 accumulation in TMEM (precision study: tools/precision_study.py).
// Class operands are reused by GROUP=2 windows per load; everything is double-buffered through mbarriers.
#include "arx_internal.cuh"
#include "arx_ptx.cuh"
#include <cuda_bf16.h>
#include <utility>

namespace {
using namespace ptx;

constexpr int TILE = 128;
constexpr int DD = 128;
constexpr uint32_t IMG_BYTES = TILE * DD * 2;      // 32 KB fp16 operand image (two 16 KB SW128 sub-tiles)
constexpr uint32_t SUB_BYTES = TILE * 64 * 2;
constexpr int GROUP = 2;
constexpr int NTHREADS = 384;

// shared memory carve-up (offsets from a 1024-aligned base)
constexpr uint32_t OFF_KQ = 0;                      // 2 x 32 KB
constexpr uint32_t OFF_KC = 2 * IMG_BYTES;          // 2 x 32 KB
constexpr uint32_t OFF_VCT = 4 * IMG_BYTES;         // 2 x 32 KB
constexpr uint32_t OFF_P = 6 * IMG_BYTES;           // 32 KB
constexpr uint32_t OFF_BAR = 7 * IMG_BYTES;
enum { B_FULL_KQ = 0, B_EMPTY_KQ = 2, B_FULL_C = 4, B_EMPTY_C = 6, B_S_FULL = 8, B_S_EMPTY = 10, B_P_FULL = 12, B_P_EMPTY = 13,
       B_O_FULL = 14, B_O_EMPTY = 16, B_COUNT = 18 };
constexpr uint32_t SMEM_BYTES = OFF_BAR + B_COUNT * 8 + 16 + 1024;

struct AttnParams {
  const __half *kq_img;    // [n_win] images, K-major SW128, rows = query tuples (pre-scaled)
  const __half *kc_img;    // [way]   images, rows = support tuples
  const __half *vct_img;   // [way]   images, rows = d, cols = support tuples
  const float *G;          // [n_win*T][ldg] per-frame projections; V part p at column (c+p)*D   (MODE 0)
  const float *Vq;         // [n_win][N][D] fp32 tuple values                                     (MODE 1)
  float *partial;          // [n_win][way][4]
  int n_win, way, N, T, ldg, voff;
  long long *trace;        // optional per-role timestamps of CTA 0 (bring-up tool), [role][tile][8]
};

#define ARX_TRACE_TILES 64
#define TRACE(role, tile, k) do { if (p.trace && blockIdx.x == 0 && (tile) < ARX_TRACE_TILES) p.trace[(((role) * ARX_TRACE_TILES) + (tile)) * 8 + (k)] = clock64(); } while (0)

struct TileIter {
  int n_win, way, n_groups, gstride;
  int group, gi, c, w, nw;
  bool valid;
  __device__ void init(int n_win_, int way_, int first, int stride) {
    n_win = n_win_; way = way_; n_groups = (n_win + GROUP - 1) / GROUP; gstride = stride;
    group = first; gi = 0; c = 0; w = 0;
    valid = group < n_groups;
    nw = valid ? min(GROUP, n_win - group * GROUP) : 0;
  }
  __device__ void next() {
    if (++w == nw) {
      w = 0;
      if (++c == way) {
        c = 0; group += gstride; ++gi;
        valid = group < n_groups;
        nw = valid ? min(GROUP, n_win - group * GROUP) : 0;
      }
    }
  }
  __device__ int window() const { return group * GROUP + w; }
  __device__ int cls_counter() const { return gi * way + c; }
};

// lexicographic rank -> (i, j) of itertools.combinations(range(T), 2)
__host__ __device__ constexpr int pair_i(int q, int T) {
  int i = 0, start = 0;
  while (q >= start + (T - 1 - i)) { start += T - 1 - i; ++i; }
  return i;
}
__host__ __device__ constexpr int pair_j(int q, int T) {
  int i = 0, start = 0;
  while (q >= start + (T - 1 - i)) { start += T - 1 - i; ++i; }
  return i + 1 + (q - start);
}

template <int Q> __device__ __forceinline__ void acc_one16(const float (&a)[16], const float (&b)[16], const uint32_t (&r)[32], float &acc) {
  if constexpr (Q < 120) {
    constexpr int I = pair_i(Q, 16), J = pair_j(Q, 16);
    const float diff = (a[I] + b[J]) - __uint_as_float(r[Q & 31]);
    acc = fmaf(diff, diff, acc);
  }
}
template <int CH, int... Js>
__device__ __forceinline__ void acc_chunk16(const float (&a)[16], const float (&b)[16], const uint32_t (&r)[32], float &acc,
                                            std::integer_sequence<int, Js...>) {
  (acc_one16<CH * 32 + Js>(a, b, r, acc), ...);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// volatile: keeps the MUFU ops of a row in one back-to-back batch
__device__ __forceinline__ uint32_t ex2_bits(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}

// MODE 0: T=16 pair tuples, Vq rebuilt from per-frame V projections in registers.  MODE 1: Vq read from HBM.
template <int MODE, bool P_MN>
__global__ void __launch_bounds__(NTHREADS, 1) k_attn_tc(const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + B_COUNT * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_FULL_KQ + i], 1); mbar_init(&bars[B_EMPTY_KQ + i], 1);
      mbar_init(&bars[B_FULL_C + i], 1); mbar_init(&bars[B_EMPTY_C + i], 1);
      mbar_init(&bars[B_S_FULL + i], 1); mbar_init(&bars[B_S_EMPTY + i], 128);
      mbar_init(&bars[B_O_FULL + i], 1); mbar_init(&bars[B_O_EMPTY + i], 128);
    }
    mbar_init(&bars[B_P_FULL], 128); mbar_init(&bars[B_P_EMPTY], 1);
    mbar_init_fence();
  }
  if (warp == 3) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;                 // lane 0, column 0 of the allocation
  const uint32_t TM_S = tmem, TM_O = tmem + 256;    // S^T buffers: cols [0,256); proto^T buffers: cols [256,512)

  if (warp == 0) {
    // ---------------- producer: class operands (Kc, Vc^T), one stage per class, reused by the group's windows
    if (elect_one()) {
      TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
      int cl = 0;
      while (it.valid) {
        for (int c = 0; c < p.way; ++c, ++cl) {
          const int st = cl & 1;
          mbar_wait(&bars[B_EMPTY_C + st], ((cl >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[B_FULL_C + st], 2 * IMG_BYTES);
          const uint8_t *kc = reinterpret_cast<const uint8_t *>(p.kc_img) + (size_t)c * IMG_BYTES;
          const uint8_t *vc = reinterpret_cast<const uint8_t *>(p.vct_img) + (size_t)c * IMG_BYTES;
          bulk_g2s(smem + OFF_KC + st * IMG_BYTES, kc, SUB_BYTES, &bars[B_FULL_C + st]);
          bulk_g2s(smem + OFF_KC + st * IMG_BYTES + SUB_BYTES, kc + SUB_BYTES, SUB_BYTES, &bars[B_FULL_C + st]);
          bulk_g2s(smem + OFF_VCT + st * IMG_BYTES, vc, SUB_BYTES, &bars[B_FULL_C + st]);
          bulk_g2s(smem + OFF_VCT + st * IMG_BYTES + SUB_BYTES, vc + SUB_BYTES, SUB_BYTES, &bars[B_FULL_C + st]);
        }
        // skip to the next group of this CTA
        it.c = p.way - 1; it.w = it.nw - 1; it.next();
      }
    }
  } else if (warp == 2) {
    // ---------------- producer: query-side Kq images, one slot per window of the group
    if (elect_one()) {
      TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
      while (it.valid) {
        for (int w = 0; w < it.nw; ++w) {
          mbar_wait(&bars[B_EMPTY_KQ + w], (it.gi & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[B_FULL_KQ + w], IMG_BYTES);
          const uint8_t *src = reinterpret_cast<const uint8_t *>(p.kq_img) + (size_t)(it.group * GROUP + w) * IMG_BYTES;
          bulk_g2s(smem + OFF_KQ + w * IMG_BYTES, src, SUB_BYTES, &bars[B_FULL_KQ + w]);
          bulk_g2s(smem + OFF_KQ + w * IMG_BYTES + SUB_BYTES, src + SUB_BYTES, SUB_BYTES, &bars[B_FULL_KQ + w]);
        }
        it.c = p.way - 1; it.w = it.nw - 1; it.next();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (one thread): MMA1 of tile f+1 is issued before MMA2 of tile f
    if (elect_one()) {
      constexpr uint64_t DESC_K = smem_desc_sw128(16, 1024);          // K-major SW128
      constexpr uint64_t DESC_MN = smem_desc_sw128(16384, 1024);      // MN-major SW128 (P)
      constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);
      constexpr uint32_t IDESC2 = idesc_f16(128, 128, 0, P_MN ? 1 : 0);
      const uint32_t sbase = smem_u32(smem);
      TileIter it1, it2;
      it1.init(p.n_win, p.way, blockIdx.x, gridDim.x);
      it2 = it1;
      int f1 = 0, f2 = 0;
      auto mma1 = [&]() {
        const int cc = it1.cls_counter(), st = cc & 1, buf = f1 & 1;
        if (it1.c == 0) mbar_wait(&bars[B_FULL_KQ + it1.w], it1.gi & 1);
        if (it1.w == 0) mbar_wait(&bars[B_FULL_C + st], (cc >> 1) & 1);
        mbar_wait(&bars[B_S_EMPTY + buf], ((f1 >> 1) & 1) ^ 1);
        tc_fence_after();
        TRACE(0, f1, 0);
        const uint32_t a0 = sbase + OFF_KC + st * IMG_BYTES, b0 = sbase + OFF_KQ + it1.w * IMG_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
          mma_f16_ss(TM_S + buf * 128, smem_desc_at(DESC_K, a0 + off), smem_desc_at(DESC_K, b0 + off), IDESC1, kk > 0);
        }
        mma_commit(&bars[B_S_FULL + buf]);
        if (it1.c == p.way - 1) mma_commit(&bars[B_EMPTY_KQ + it1.w]);
        ++f1; it1.next();
      };
      auto mma2 = [&]() {
        const int cc = it2.cls_counter(), st = cc & 1, buf = f2 & 1;
        TRACE(0, f2, 1);
        mbar_wait(&bars[B_P_FULL], f2 & 1);
        TRACE(0, f2, 2);
        mbar_wait(&bars[B_O_EMPTY + buf], ((f2 >> 1) & 1) ^ 1);
        tc_fence_after();
        TRACE(0, f2, 3);
        const uint32_t a0 = sbase + OFF_VCT + st * IMG_BYTES, b0 = sbase + OFF_P;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
          const uint64_t bd = P_MN ? smem_desc_at(DESC_MN, b0 + kk * 2048) : smem_desc_at(DESC_K, b0 + off);
          mma_f16_ss(TM_O + buf * 128, smem_desc_at(DESC_K, a0 + off), bd, IDESC2, kk > 0);
        }
        mma_commit(&bars[B_O_FULL + buf]);
        mma_commit(&bars[B_P_EMPTY]);
        if (it2.w == it2.nw - 1) mma_commit(&bars[B_EMPTY_C + st]);
        ++f2; it2.next();
      };
      if (it1.valid) mma1();
      while (it2.valid) {
        if (it1.valid) mma1();
        mma2();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------- softmax warps: thread == support tuple s == TMEM lane
    const int quad = warp - 4;
    const int s = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint8_t *pbuf = smem + OFF_P;
    TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
    int f = 0;
    while (it.valid) {
      const int buf = f & 1;
      if (threadIdx.x == 128) TRACE(1, f, 0);
      mbar_wait(&bars[B_S_FULL + buf], (f >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 128) TRACE(1, f, 1);
      uint32_t r[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld32(TM_S + lane_base + buf * 128 + ch * 32, r[ch]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars[B_S_EMPTY + buf]);
      if (threadIdx.x == 128) TRACE(1, f, 2);
      // exp2 of the whole row first (128 independent MUFU ops, 8 clk each: the MUFU floor), the sums afterwards on
      // four independent accumulators -- keeps the in-order issue from stalling on the MUFU->FADD latency
      float zp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[ch][j] = ex2_bits(r[ch][j]);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if ((ch + 1) * 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j) zp[j & 3] += __uint_as_float(r[ch][j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) zp[j & 3] += (ch * 32 + j < p.N) ? __uint_as_float(r[ch][j]) : 0.f;
        }
      }
      const float z = (zp[0] + zp[1]) + (zp[2] + zp[3]);
      const float zinv = __frcp_rn(z);
      if (threadIdx.x == 128) TRACE(1, f, 3);
      mbar_wait(&bars[B_P_EMPTY], (f & 1) ^ 1);
      if (threadIdx.x == 128) TRACE(1, f, 4);
      if constexpr (P_MN) {
        // B operand, MN-major SW128: memory row = support tuple s (K index), 64 query tuples per 128-byte row
        uint8_t *row = pbuf + (s >> 3) * 1024 + (s & 7) * 128;
#pragma unroll
        for (int c16 = 0; c16 < 16; ++c16) {
          const int ch = c16 >> 2, j0 = (c16 & 3) * 8;
          uint4 v;
          v.x = pack_half2(__uint_as_float(r[ch][j0 + 0]) * zinv, __uint_as_float(r[ch][j0 + 1]) * zinv);
          v.y = pack_half2(__uint_as_float(r[ch][j0 + 2]) * zinv, __uint_as_float(r[ch][j0 + 3]) * zinv);
          v.z = pack_half2(__uint_as_float(r[ch][j0 + 4]) * zinv, __uint_as_float(r[ch][j0 + 5]) * zinv);
          v.w = pack_half2(__uint_as_float(r[ch][j0 + 6]) * zinv, __uint_as_float(r[ch][j0 + 7]) * zinv);
          *reinterpret_cast<uint4 *>(row + (c16 >> 3) * 16384 + (((c16 & 7) ^ (s & 7)) << 4)) = v;
        }
      } else {
        // B operand, K-major SW128: memory row = query tuple q, support tuples contiguous
        uint8_t *col = pbuf + (s >> 6) * SUB_BYTES + (s & 7) * 2;
        const int sc = (s & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 128; ++q) {
          const __half hv = __float2half_rn(__uint_as_float(r[q >> 5][q & 31]) * zinv);
          *reinterpret_cast<__half *>(col + q * 128 + ((sc ^ (q & 7)) << 4)) = hv;
        }
      }
      if (threadIdx.x == 128) TRACE(1, f, 5);
      fence_proxy_async_smem();
      mbar_arrive(&bars[B_P_FULL]);
      if (threadIdx.x == 128) TRACE(1, f, 6);
      ++f; it.next();
    }
  } else if (warp >= 8) {
    // ---------------- epilogue warps: thread == output dimension d == TMEM lane
    const int quad = warp - 8;
    const int d = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
    int f = 0;
    float a0[16], b0[16], a1[16], b1[16];
    while (it.valid) {
      if constexpr (MODE == 0) {
        if (it.c == 0 && it.w == 0) {
          // per-frame V projections of the group's windows (bias already folded into part 0)
          const float *g0 = p.G + (size_t)(it.group * GROUP) * 16 * p.ldg + p.voff + d;
#pragma unroll
          for (int i = 0; i < 16; ++i) { a0[i] = __ldg(g0 + (size_t)i * p.ldg); b0[i] = __ldg(g0 + (size_t)i * p.ldg + DD); }
          if (it.nw > 1) {
            const float *g1 = g0 + (size_t)16 * p.ldg;
#pragma unroll
            for (int i = 0; i < 16; ++i) { a1[i] = __ldg(g1 + (size_t)i * p.ldg); b1[i] = __ldg(g1 + (size_t)i * p.ldg + DD); }
          }
        }
      }
      const int buf = f & 1;
      if (threadIdx.x == 256) TRACE(2, f, 0);
      mbar_wait(&bars[B_O_FULL + buf], (f >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 256) TRACE(2, f, 1);
      float acc = 0.f;
      uint32_t r[32];
      if constexpr (MODE == 0) {
        auto run = [&](const float (&a)[16], const float (&b)[16]) {
          uint32_t r2[32];
          tmem_ld32(TM_O + lane_base + buf * 128 + 0, r);
          tmem_ld32(TM_O + lane_base + buf * 128 + 32, r2);
          tmem_ld_wait();
          acc_chunk16<0>(a, b, r, acc, std::make_integer_sequence<int, 32>{});
          tmem_ld32(TM_O + lane_base + buf * 128 + 64, r);
          acc_chunk16<1>(a, b, r2, acc, std::make_integer_sequence<int, 32>{});
          tmem_ld_wait();
          tmem_ld32(TM_O + lane_base + buf * 128 + 96, r2);
          acc_chunk16<2>(a, b, r, acc, std::make_integer_sequence<int, 32>{});
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&bars[B_O_EMPTY + buf]);
          acc_chunk16<3>(a, b, r2, acc, std::make_integer_sequence<int, 32>{});
        };
        if (it.w == 0) run(a0, b0); else run(a1, b1);
      } else {
        const float *vq = p.Vq + (size_t)it.window() * p.N * DD + d;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          tmem_ld32(TM_O + lane_base + buf * 128 + ch * 32, r); tmem_ld_wait();
          if (ch == 3) { tc_fence_before(); mbar_arrive(&bars[B_O_EMPTY + buf]); }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = ch * 32 + j;
            if (q < p.N) {
              const float diff = __ldg(vq + (size_t)q * DD) - __uint_as_float(r[j]);
              acc = fmaf(diff, diff, acc);
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) p.partial[((size_t)it.window() * p.way + it.c) * 4 + quad] = acc;
      if (threadIdx.x == 256) TRACE(2, f, 2);
      ++f; it.next();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- operand image builders -------------------------------------------------------------------------
// Query/support K image: one warp per tuple row; K = LayerNorm(sum_p Gk_p[frame_p]) * alpha -> fp16, K-major SW128.
__global__ void __launch_bounds__(256) k_prep_k_img(const float *__restrict__ G, const int32_t *__restrict__ tuples,
                                                    const float *__restrict__ ln_g, const float *__restrict__ ln_b,
                                                    __half *__restrict__ img, int T, int c, int N, int ldg, float alpha) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t seq = blockIdx.x;
  uint8_t *out = reinterpret_cast<uint8_t *>(img) + seq * IMG_BYTES;
  const int d0 = lane * 4;
  const float4 g = *reinterpret_cast<const float4 *>(ln_g + d0);
  const float4 be = *reinterpret_cast<const float4 *>(ln_b + d0);
  for (int r = warp; r < TILE; r += 8) {
    uint2 packed = make_uint2(0u, 0u);
    if (r < N && tuples[r * c] >= 0) {
      float4 k = make_float4(0, 0, 0, 0);
      for (int pp = 0; pp < c; ++pp) {
        const int fr = tuples[r * c + pp];
        const float4 a = *reinterpret_cast<const float4 *>(G + (seq * T + fr) * (size_t)ldg + pp * DD + d0);
        k.x += a.x; k.y += a.y; k.z += a.z; k.w += a.w;
      }
      float s = k.x + k.y + k.z + k.w;
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / DD;
      const float4 dl = make_float4(k.x - mean, k.y - mean, k.z - mean, k.w - mean);
      float q = dl.x * dl.x + dl.y * dl.y + dl.z * dl.z + dl.w * dl.w;
#pragma unroll
      for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = 1.0f / sqrtf(q / DD + 1e-5f);
      packed.x = pack_half2((dl.x * rstd * g.x + be.x) * alpha, (dl.y * rstd * g.y + be.y) * alpha);
      packed.y = pack_half2((dl.z * rstd * g.z + be.z) * alpha, (dl.w * rstd * g.w + be.w) * alpha);
    }
    *reinterpret_cast<uint2 *>(out + (d0 >> 6) * SUB_BYTES + sw128_offset(r, d0 & 63)) = packed;
  }
}

// Support V^T image: rows = d, cols = support tuple s (zero for s >= N); from fp32 vs (way, N, D).
__device__ __forceinline__ uint32_t pack_bf162(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <bool BF16>
__global__ void __launch_bounds__(256) k_prep_vct_img(const float *__restrict__ vs, __half *__restrict__ img, int N) {
  const size_t cls = blockIdx.x;
  const float *v = vs + cls * (size_t)N * DD;
  uint8_t *out = reinterpret_cast<uint8_t *>(img) + cls * IMG_BYTES;
  for (int e = threadIdx.x; e < DD * 16; e += 256) {
    const int d = e & 127, sc = e >> 7;          // consecutive threads -> consecutive d (coalesced reads)
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s = sc * 8 + i;
      x[i] = s < N ? v[(size_t)s * DD + d] : 0.f;
    }
    uint4 pk;
    if constexpr (BF16) {
      pk.x = pack_bf162(x[0], x[1]); pk.y = pack_bf162(x[2], x[3]); pk.z = pack_bf162(x[4], x[5]); pk.w = pack_bf162(x[6], x[7]);
    } else {
      pk.x = pack_half2(x[0], x[1]); pk.y = pack_half2(x[2], x[3]); pk.z = pack_half2(x[4], x[5]); pk.w = pack_half2(x[6], x[7]);
    }
    const int s0 = sc * 8;
    *reinterpret_cast<uint4 *>(out + (s0 >> 6) * SUB_BYTES + sw128_offset(d, s0 & 63)) = pk;
  }
}

// Support K image from fp32 ks (way, N, D): rows = s, cols = d.
__global__ void __launch_bounds__(256) k_prep_kc_img(const float *__restrict__ ks, __half *__restrict__ img, int N) {
  const size_t cls = blockIdx.x;
  const float *k = ks + cls * (size_t)N * DD;
  uint8_t *out = reinterpret_cast<uint8_t *>(img) + cls * IMG_BYTES;
  for (int e = threadIdx.x; e < TILE * 16; e += 256) {
    const int dc = e & 15, s = e >> 4;
    uint4 pk = make_uint4(0, 0, 0, 0);
    if (s < N) {
      const float4 x0 = *reinterpret_cast<const float4 *>(k + (size_t)s * DD + dc * 8);
      const float4 x1 = *reinterpret_cast<const float4 *>(k + (size_t)s * DD + dc * 8 + 4);
      pk.x = pack_half2(x0.x, x0.y); pk.y = pack_half2(x0.z, x0.w); pk.z = pack_half2(x1.x, x1.y); pk.w = pack_half2(x1.z, x1.w);
    }
    const int d0 = dc * 8;
    *reinterpret_cast<uint4 *>(out + (d0 >> 6) * SUB_BYTES + sw128_offset(s, d0 & 63)) = pk;
  }
}

// Support side in ONE launch (model.py:69-84 for the support set): tuple gather, LayerNorm, fp32 K/V tuple tensors
// (general path, export blob) and the three tcgen05 operand images (Kc fp16, Vc^T fp16 and bf16).  One block per class.
__global__ void __launch_bounds__(256) k_support_build(const float *__restrict__ G, const int32_t *__restrict__ tuples,
                                                       const float *__restrict__ ln_g, const float *__restrict__ ln_b,
                                                       float *__restrict__ ks, float *__restrict__ vs, __half *__restrict__ kc_img,
                                                       __half *__restrict__ vct_img, __half *__restrict__ vct_img_bf, int T, int c, int N,
                                                       int ldg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t cls = blockIdx.x;
  const int d0 = lane * 4;
  const float4 g = *reinterpret_cast<const float4 *>(ln_g + d0);
  const float4 be = *reinterpret_cast<const float4 *>(ln_b + d0);
  uint8_t *kc = kc_img ? reinterpret_cast<uint8_t *>(kc_img) + cls * IMG_BYTES : nullptr;
  uint8_t *vt = vct_img ? reinterpret_cast<uint8_t *>(vct_img) + cls * IMG_BYTES : nullptr;
  uint8_t *vb = vct_img_bf ? reinterpret_cast<uint8_t *>(vct_img_bf) + cls * IMG_BYTES : nullptr;
  const int rows = kc ? TILE : N;
  for (int r = blockIdx.y * 8 + warp; r < rows; r += 8 * gridDim.y) {
    float4 k = make_float4(0, 0, 0, 0), v = make_float4(0, 0, 0, 0);
    uint2 kpk = make_uint2(0u, 0u);
    if (r < N) {
      for (int pp = 0; pp < c; ++pp) {
        const int fr = tuples[r * c + pp];
        const float *row = G + (cls * T + fr) * (size_t)ldg;
        const float4 a = *reinterpret_cast<const float4 *>(row + pp * DD + d0);
        const float4 b = *reinterpret_cast<const float4 *>(row + (c + pp) * DD + d0);
        k.x += a.x; k.y += a.y; k.z += a.z; k.w += a.w;
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      float sum = k.x + k.y + k.z + k.w;
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / DD;
      const float4 dl = make_float4(k.x - mean, k.y - mean, k.z - mean, k.w - mean);
      float q = dl.x * dl.x + dl.y * dl.y + dl.z * dl.z + dl.w * dl.w;
#pragma unroll
      for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = 1.0f / sqrtf(q / DD + 1e-5f);
      k = make_float4(dl.x * rstd * g.x + be.x, dl.y * rstd * g.y + be.y, dl.z * rstd * g.z + be.z, dl.w * rstd * g.w + be.w);
      *reinterpret_cast<float4 *>(ks + (cls * N + r) * DD + d0) = k;
      *reinterpret_cast<float4 *>(vs + (cls * N + r) * DD + d0) = v;
      kpk.x = pack_half2(k.x, k.y);
      kpk.y = pack_half2(k.z, k.w);
    }
    if (kc) {
      *reinterpret_cast<uint2 *>(kc + (d0 >> 6) * SUB_BYTES + sw128_offset(r, d0 & 63)) = kpk;      // row r (zero beyond N)
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {                                                                 // column r of Vc^T (zero beyond N)
        const uint32_t off = (r >> 6) * SUB_BYTES + sw128_offset(d0 + e, r & 63);
        *reinterpret_cast<__half *>(vt + off) = __float2half_rn(vv[e]);
        *reinterpret_cast<__nv_bfloat16 *>(vb + off) = __float2bfloat16_rn(vv[e]);
      }
    }
  }
}

__global__ void k_finish_tc(const float *__restrict__ partial, float *__restrict__ logits, int32_t *__restrict__ chosen, int64_t n_win,
                            int way, int N) {
  pdl_trigger();
  pdl_wait();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_win) return;
  float best = -INFINITY;
  int bi = 0;
  for (int c = 0; c < way; ++c) {
    const float4 t = *reinterpret_cast<const float4 *>(partial + (b * way + c) * 4);
    const float lg = -(((t.x + t.y) + (t.z + t.w)) / (float)N);
    logits[b * way + c] = lg;
    if (lg > best) { best = lg; bi = c; }
  }
  if (chosen) chosen[b] = bi;
}


// =====================================================================================================
// Open-set head pass (model.py:323-324,196): attention for the WINNING class of every window only, then
//   y[q,l] = diff[q,:] . Wdr[l,:] + bdr[l]   as a third MMA:  Y[q,l]  M=128 (q on TMEM lanes), N=L, K=128 (d)
// with A = diff^T written by the epilogue warps in the same MN-major layout as P (thread == d == K index)
// and B = Wdr (L x 128, K-major).  One tile per window; operands: Kq[b], Kc[chosen[b]], Vc^T[chosen[b]].
constexpr uint32_t H_OFF_KQ = 0;                       // 2 x 32 KB
constexpr uint32_t H_OFF_KC = 2 * IMG_BYTES;           // 2 x 32 KB
constexpr uint32_t H_OFF_VCT = 4 * IMG_BYTES;          // 32 KB
constexpr uint32_t H_OFF_P = 5 * IMG_BYTES;            // P, then diff^T (shared), 32 KB
constexpr uint32_t H_OFF_WDR = 6 * IMG_BYTES;          // up to 8 KB
constexpr uint32_t H_OFF_BAR = 6 * IMG_BYTES + 8192;
enum { HB_FULL_A = 0, HB_EMPTY_A = 2, HB_FULL_V = 4, HB_EMPTY_V = 5, HB_S_FULL = 6, HB_S_EMPTY = 8, HB_P_FULL = 10, HB_P_EMPTY = 11,
       HB_O_FULL = 12, HB_O_EMPTY = 13, HB_DF_FULL = 14, HB_Y_FULL = 15, HB_Y_EMPTY = 16, HB_WDR = 17, HB_COUNT = 18 };
constexpr uint32_t H_SMEM_BYTES = H_OFF_BAR + HB_COUNT * 8 + 16 + 1024;

struct HeadParams {
  const __half *kq_img, *kc_img, *vct_img, *wdr_img;
  const float *G, *Vq, *dr_b;
  const int32_t *chosen;
  float *y;                // [n_win][N*L] fp32, or
  __half *y_img;           // fp16 activation image [ceil(n_win/128)][y_nk][128 x 64] for the tcgen05 GEMM of fc1
  int n_win, way, N, T, ldg, voff, L, y_nk;
};

// lexicographic rank of every slot of the padded-triangular query order (-1 = pad); filled by arx_tc_head_features
__constant__ short c_slot_rank[128];

template <int... Is> __device__ __forceinline__ void zero_pad_slots(uint32_t (&r)[4][32], std::integer_sequence<int, Is...>) {
  ((r[arx_slot_row_start(2 * Is) >> 5][arx_slot_row_start(2 * Is) & 31] = 0u), ...);   // the pad slot of every even row
}

template <bool SLOT, int Q> struct TupleOf {   // (i, j) of column Q in the kernel's query-tuple order; valid == not a pad
  static constexpr int i = SLOT ? arx_slot_i(Q) : pair_i(Q < 120 ? Q : 0, 16);
  static constexpr int j = SLOT ? arx_slot_j(Q) : pair_j(Q < 120 ? Q : 0, 16);
  static constexpr bool valid = SLOT ? (arx_slot_j(Q) != arx_slot_i(Q)) : (Q < 120);
};
template <bool SLOT, int Q2> __device__ __forceinline__ void diff_pair16(const float (&a)[16], const float (&b)[16], const uint32_t (&r)[32], uint32_t (&pk)[64]) {
  constexpr int Q = 2 * Q2;
  float d0 = 0.f, d1 = 0.f;
  if constexpr (TupleOf<SLOT, Q>::valid) d0 = (a[TupleOf<SLOT, Q>::i] + b[TupleOf<SLOT, Q>::j]) - __uint_as_float(r[Q & 31]);
  if constexpr (TupleOf<SLOT, Q + 1>::valid) d1 = (a[TupleOf<SLOT, Q + 1>::i] + b[TupleOf<SLOT, Q + 1>::j]) - __uint_as_float(r[(Q + 1) & 31]);
  pk[Q2] = pack_half2(d0, d1);
}
template <bool SLOT, int CH, int... Js>
__device__ __forceinline__ void diff_chunk16(const float (&a)[16], const float (&b)[16], const uint32_t (&r)[32], uint32_t (&pk)[64],
                                             std::integer_sequence<int, Js...>) {
  (diff_pair16<SLOT, CH * 16 + Js>(a, b, r, pk), ...);
}

template <int MODE, int L, bool SLOT>
__global__ void __launch_bounds__(NTHREADS, 1) k_head_tc(const HeadParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + H_OFF_BAR);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + H_OFF_BAR + HB_COUNT * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.n_win > (int)blockIdx.x ? (p.n_win - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[HB_FULL_A + i], 1); mbar_init(&bars[HB_EMPTY_A + i], 1);
      mbar_init(&bars[HB_S_FULL + i], 1); mbar_init(&bars[HB_S_EMPTY + i], 128);
    }
    mbar_init(&bars[HB_FULL_V], 1); mbar_init(&bars[HB_EMPTY_V], 1);
    mbar_init(&bars[HB_P_FULL], 128); mbar_init(&bars[HB_P_EMPTY], 1);
    mbar_init(&bars[HB_O_FULL], 1); mbar_init(&bars[HB_O_EMPTY], 128);
    mbar_init(&bars[HB_DF_FULL], 128); mbar_init(&bars[HB_Y_FULL], 1); mbar_init(&bars[HB_Y_EMPTY], 128);
    mbar_init(&bars[HB_WDR], 1);
    mbar_init_fence();
  }
  if (warp == 3) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S = tmem, TM_O = tmem + 256, TM_Y = tmem + 384;
  constexpr uint32_t WDR_BYTES = L * DD * 2;

  if (warp == 0) {
    if (elect_one()) {                 // producer A: Wdr once, then {Kq[b], Kc[chosen[b]]} per tile
      mbar_arrive_expect_tx(&bars[HB_WDR], WDR_BYTES);
      bulk_g2s(smem + H_OFF_WDR, p.wdr_img, WDR_BYTES, &bars[HB_WDR]);
      for (int t = 0; t < ntiles; ++t) {
        const int b = blockIdx.x + t * gridDim.x, c = p.chosen[b], sa = t & 1;
        mbar_wait(&bars[HB_EMPTY_A + sa], ((t >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[HB_FULL_A + sa], 2 * IMG_BYTES);
        const uint8_t *kq = reinterpret_cast<const uint8_t *>(p.kq_img) + (size_t)b * IMG_BYTES;
        const uint8_t *kc = reinterpret_cast<const uint8_t *>(p.kc_img) + (size_t)c * IMG_BYTES;
        bulk_g2s(smem + H_OFF_KQ + sa * IMG_BYTES, kq, SUB_BYTES, &bars[HB_FULL_A + sa]);
        bulk_g2s(smem + H_OFF_KQ + sa * IMG_BYTES + SUB_BYTES, kq + SUB_BYTES, SUB_BYTES, &bars[HB_FULL_A + sa]);
        bulk_g2s(smem + H_OFF_KC + sa * IMG_BYTES, kc, SUB_BYTES, &bars[HB_FULL_A + sa]);
        bulk_g2s(smem + H_OFF_KC + sa * IMG_BYTES + SUB_BYTES, kc + SUB_BYTES, SUB_BYTES, &bars[HB_FULL_A + sa]);
      }
    }
  } else if (warp == 2) {
    if (elect_one()) {                 // producer V: Vc^T[chosen[b]] per tile (single stage)
      for (int t = 0; t < ntiles; ++t) {
        const int b = blockIdx.x + t * gridDim.x, c = p.chosen[b];
        mbar_wait(&bars[HB_EMPTY_V], (t & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[HB_FULL_V], IMG_BYTES);
        const uint8_t *vc = reinterpret_cast<const uint8_t *>(p.vct_img) + (size_t)c * IMG_BYTES;
        bulk_g2s(smem + H_OFF_VCT, vc, SUB_BYTES, &bars[HB_FULL_V]);
        bulk_g2s(smem + H_OFF_VCT + SUB_BYTES, vc + SUB_BYTES, SUB_BYTES, &bars[HB_FULL_V]);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint64_t DESC_K = smem_desc_sw128(16, 1024);
      constexpr uint64_t DESC_MN = smem_desc_sw128(16384, 1024);
      constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);
      constexpr uint32_t IDESC2 = idesc_f16(128, 128, 0, 1);
      constexpr uint32_t IDESC3 = idesc_f16(128, L, 1, 0);
      const uint32_t sbase = smem_u32(smem);
      mbar_wait(&bars[HB_WDR], 0);
      auto mma1 = [&](int t) {
        const int sa = t & 1;
        mbar_wait(&bars[HB_FULL_A + sa], (t >> 1) & 1);
        mbar_wait(&bars[HB_S_EMPTY + sa], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a0 = sbase + H_OFF_KC + sa * IMG_BYTES, b0 = sbase + H_OFF_KQ + sa * IMG_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
          mma_f16_ss(TM_S + sa * 128, smem_desc_at(DESC_K, a0 + off), smem_desc_at(DESC_K, b0 + off), IDESC1, kk > 0);
        }
        mma_commit(&bars[HB_S_FULL + sa]);
        mma_commit(&bars[HB_EMPTY_A + sa]);
      };
      if (ntiles > 0) mma1(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) mma1(t + 1);
        // MMA2: proto^T = Vc^T . P
        mbar_wait(&bars[HB_P_FULL], t & 1);
        mbar_wait(&bars[HB_FULL_V], t & 1);
        mbar_wait(&bars[HB_O_EMPTY], (t & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
          mma_f16_ss(TM_O, smem_desc_at(DESC_K, sbase + H_OFF_VCT + off), smem_desc_at(DESC_MN, sbase + H_OFF_P + kk * 2048), IDESC2, kk > 0);
        }
        mma_commit(&bars[HB_O_FULL]);
        mma_commit(&bars[HB_EMPTY_V]);
        // MMA3: Y = diff . Wdr^T   (A = diff^T image in the P buffer, MN-major; B = Wdr, K-major)
        mbar_wait(&bars[HB_DF_FULL], t & 1);
        mbar_wait(&bars[HB_Y_EMPTY], (t & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t boff = (kk >> 2) * (L * 128) + (kk & 3) * 32;
          mma_f16_ss(TM_Y, smem_desc_at(DESC_MN, sbase + H_OFF_P + kk * 2048), smem_desc_at(DESC_K, sbase + H_OFF_WDR + boff), IDESC3, kk > 0);
        }
        mma_commit(&bars[HB_Y_FULL]);
        mma_commit(&bars[HB_P_EMPTY]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int quad = warp - 4;
    const int s = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint8_t *row = smem + H_OFF_P + (s >> 3) * 1024 + (s & 7) * 128;
    for (int t = 0; t < ntiles; ++t) {
      const int sa = t & 1;
      mbar_wait(&bars[HB_S_FULL + sa], (t >> 1) & 1);
      tc_fence_after();
      uint32_t r[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld32(TM_S + lane_base + sa * 128 + ch * 32, r[ch]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars[HB_S_EMPTY + sa]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[ch][j] = ex2_bits(r[ch][j]);
      }
      if constexpr (SLOT) {
        zero_pad_slots(r, std::make_integer_sequence<int, 8>{});
      } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[ch][j] = (ch * 32 + j < p.N) ? r[ch][j] : 0u;
        }
      }
      float zp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
        for (int j = 0; j < 32; ++j) zp[j & 3] += __uint_as_float(r[ch][j]);
      }
      const float z = (zp[0] + zp[1]) + (zp[2] + zp[3]);
      const float zinv = 1.0f / z;
      mbar_wait(&bars[HB_P_EMPTY], (t & 1) ^ 1);
#pragma unroll
      for (int c16 = 0; c16 < 16; ++c16) {
        const int ch = c16 >> 2, j0 = (c16 & 3) * 8;
        uint4 v;
        v.x = pack_half2(__uint_as_float(r[ch][j0 + 0]) * zinv, __uint_as_float(r[ch][j0 + 1]) * zinv);
        v.y = pack_half2(__uint_as_float(r[ch][j0 + 2]) * zinv, __uint_as_float(r[ch][j0 + 3]) * zinv);
        v.z = pack_half2(__uint_as_float(r[ch][j0 + 4]) * zinv, __uint_as_float(r[ch][j0 + 5]) * zinv);
        v.w = pack_half2(__uint_as_float(r[ch][j0 + 6]) * zinv, __uint_as_float(r[ch][j0 + 7]) * zinv);
        *reinterpret_cast<uint4 *>(row + (c16 >> 3) * 16384 + (((c16 & 7) ^ (s & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      mbar_arrive(&bars[HB_P_FULL]);
    }
  } else if (warp >= 8) {
    const int quad = warp - 8;
    const int d = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint8_t *row = smem + H_OFF_P + (d >> 3) * 1024 + (d & 7) * 128;
    float bias[L];
#pragma unroll
    for (int l = 0; l < L; ++l) bias[l] = __ldg(p.dr_b + l);
    for (int t = 0; t < ntiles; ++t) {
      const int b = blockIdx.x + t * gridDim.x;
      uint32_t pk[64];
      uint32_t r[32];
      if constexpr (MODE == 0) {
        float a[16], bb[16];
        const float *g0 = p.G + (size_t)b * 16 * p.ldg + p.voff + d;
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = __ldg(g0 + (size_t)i * p.ldg); bb[i] = __ldg(g0 + (size_t)i * p.ldg + DD); }
        mbar_wait(&bars[HB_O_FULL], t & 1);
        tc_fence_after();
        tmem_ld32(TM_O + lane_base + 0, r); tmem_ld_wait();
        diff_chunk16<SLOT, 0>(a, bb, r, pk, std::make_integer_sequence<int, 16>{});
        tmem_ld32(TM_O + lane_base + 32, r); tmem_ld_wait();
        diff_chunk16<SLOT, 1>(a, bb, r, pk, std::make_integer_sequence<int, 16>{});
        tmem_ld32(TM_O + lane_base + 64, r); tmem_ld_wait();
        diff_chunk16<SLOT, 2>(a, bb, r, pk, std::make_integer_sequence<int, 16>{});
        tmem_ld32(TM_O + lane_base + 96, r); tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars[HB_O_EMPTY]);
        diff_chunk16<SLOT, 3>(a, bb, r, pk, std::make_integer_sequence<int, 16>{});
      } else {
        const float *vq = p.Vq + (size_t)b * p.N * DD + d;
        mbar_wait(&bars[HB_O_FULL], t & 1);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          tmem_ld32(TM_O + lane_base + ch * 32, r); tmem_ld_wait();
          if (ch == 3) { tc_fence_before(); mbar_arrive(&bars[HB_O_EMPTY]); }
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const int q = ch * 32 + j;
            const float d0 = q < p.N ? __ldg(vq + (size_t)q * DD) - __uint_as_float(r[j]) : 0.f;
            const float d1 = q + 1 < p.N ? __ldg(vq + (size_t)(q + 1) * DD) - __uint_as_float(r[j + 1]) : 0.f;
            pk[q >> 1] = pack_half2(d0, d1);
          }
        }
      }
      // diff^T[d, :] -> A operand of MMA3 (MN-major: memory row = d, 64 query tuples per 128-byte row); the P
      // buffer is free: o_full implies MMA2 has finished reading it
#pragma unroll
      for (int c16 = 0; c16 < 16; ++c16) {
        uint4 v = make_uint4(pk[c16 * 4 + 0], pk[c16 * 4 + 1], pk[c16 * 4 + 2], pk[c16 * 4 + 3]);
        *reinterpret_cast<uint4 *>(row + (c16 >> 3) * 16384 + (((c16 & 7) ^ (d & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      mbar_arrive(&bars[HB_DF_FULL]);
      // Y[q, 0..L) for q = this thread's TMEM lane
      mbar_wait(&bars[HB_Y_FULL], t & 1);
      tc_fence_after();
      uint32_t yv[L];
      if constexpr (L == 16) {
        tmem_ld16(TM_Y + lane_base, yv);
      } else {
        tmem_ld32(TM_Y + lane_base, yv);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars[HB_Y_EMPTY]);
      // the lane index doubles as the query-tuple slot of the Y tile; y is indexed by the LEXICOGRAPHIC rank
      // (the order of the reference's reshape, model.py:197, that fc1's weights expect)
      int q = d;
      bool qok = q < p.N;
      if constexpr (SLOT) {
        q = c_slot_rank[d];          // lexicographic rank of this slot, -1 for a pad
        qok = q >= 0;
      }
      if (qok && p.y_img) {
        const int col = q * L;
        uint8_t *dst = reinterpret_cast<uint8_t *>(p.y_img) + ((size_t)(b >> 7) * p.y_nk + (col >> 6)) * (128 * 128);
#pragma unroll
        for (int l = 0; l < L; l += 8) {
          uint4 pk;
          pk.x = pack_half2(__uint_as_float(yv[l + 0]) + bias[l + 0], __uint_as_float(yv[l + 1]) + bias[l + 1]);
          pk.y = pack_half2(__uint_as_float(yv[l + 2]) + bias[l + 2], __uint_as_float(yv[l + 3]) + bias[l + 3]);
          pk.z = pack_half2(__uint_as_float(yv[l + 4]) + bias[l + 4], __uint_as_float(yv[l + 5]) + bias[l + 5]);
          pk.w = pack_half2(__uint_as_float(yv[l + 6]) + bias[l + 6], __uint_as_float(yv[l + 7]) + bias[l + 7]);
          *reinterpret_cast<uint4 *>(dst + sw128_offset(b & 127, (col & 63) + l)) = pk;
        }
      } else if (qok) {
        float *dst = p.y + ((size_t)b * p.N + q) * L;
#pragma unroll
        for (int l = 0; l < L; l += 4)
          *reinterpret_cast<float4 *>(dst + l) = make_float4(__uint_as_float(yv[l]) + bias[l], __uint_as_float(yv[l + 1]) + bias[l + 1],
                                                             __uint_as_float(yv[l + 2]) + bias[l + 2], __uint_as_float(yv[l + 3]) + bias[l + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// Wdr image: B operand of MMA3, K-major SW128: [L rows x 64 d] sub-tiles (L*128 bytes each), from fp32 (L, 128)
__global__ void k_prep_wdr_img(const float *__restrict__ w, __half *__restrict__ img, int L) {
  uint8_t *out = reinterpret_cast<uint8_t *>(img);
  for (int e = threadIdx.x; e < L * 16; e += blockDim.x) {
    const int dc = e & 15, l = e >> 4;
    const float4 x0 = *reinterpret_cast<const float4 *>(w + (size_t)l * DD + dc * 8);
    const float4 x1 = *reinterpret_cast<const float4 *>(w + (size_t)l * DD + dc * 8 + 4);
    uint4 pk;
    pk.x = pack_half2(x0.x, x0.y); pk.y = pack_half2(x0.z, x0.w); pk.z = pack_half2(x1.x, x1.y); pk.w = pack_half2(x1.z, x1.w);
    const int d0 = dc * 8;
    *reinterpret_cast<uint4 *>(out + (d0 >> 6) * (L * 128) + sw128_offset(l, d0 & 63)) = pk;
  }
}

}  // namespace

bool arx_tc_supported(const arx_handle *h, const ArxTransformer &tr) {
  // single-tile kernel: N <= 128; exp2 without max-subtraction needs the static LayerNorm bound to stay in fp32 range
  return tr.N <= TILE && h->D == DD && tr.softmax_bound * ARX_SOFTMAX_LOG2E < 100.0f;
}

int arx_tc_prep_support(arx_handle *h, ArxTransformer &tr, int way, cudaStream_t st) {
  if (!tr.ks_img) {
    ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tr.ks_img), (size_t)h->way_cap * IMG_BYTES));
    ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tr.vs_img), (size_t)h->way_cap * IMG_BYTES));
  }
  k_prep_kc_img<<<way, 256, 0, st>>>(tr.ks, tr.ks_img, tr.N);
  ARX_LAUNCH_CHECK(h);
  k_prep_vct_img<false><<<way, 256, 0, st>>>(tr.vs, tr.vs_img, tr.N);
  ARX_LAUNCH_CHECK(h);
  if (!tr.vs_img_bf) ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tr.vs_img_bf), (size_t)h->way_cap * IMG_BYTES));
  k_prep_vct_img<true><<<way, 256, 0, st>>>(tr.vs, tr.vs_img_bf, tr.N);      // bf16 copy for the second-generation kernel
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}

bool arx_tc_slot_order(const arx_handle *h, const ArxTransformer &tr) {
  return h->T == 16 && tr.c == 2 && (h->tc_variant & 8) == 0;
}

int arx_tc_prep_query(arx_handle *h, ArxTransformer &tr, const float *G, int64_t n_win, __half *kq_img, bool slot_order, cudaStream_t st) {
  const float alpha = ARX_SOFTMAX_LOG2E / sqrtf((float)h->D);
  const int32_t *table = tr.tuples;
  int rows = tr.N;
  if (slot_order) {
    if (!tr.q_slots) {
      int32_t host[256];
      arx_tc2_slot_table(host);
      ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tr.q_slots), sizeof(host)));
      ARX_CUDA(h, cudaMemcpy(tr.q_slots, host, sizeof(host), cudaMemcpyHostToDevice));
    }
    table = tr.q_slots;
    rows = 128;
  }
  k_prep_k_img<<<(unsigned)n_win, 256, 0, st>>>(G, table, tr.ln_g, tr.ln_b, kq_img, h->T, tr.c, rows, 2 * tr.c * h->D, alpha);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}

int arx_tc_attention(arx_handle *h, const ArxTransformer &tr, const __half *kq_img, const float *G, const float *Vq, int64_t n_win,
                     int way, float *partial, float *logits, int32_t *chosen, int variant, int g_ld, int g_voff, bool g_chunked, bool episodes,
                     cudaStream_t st) {
  AttnParams p{};
  p.kq_img = kq_img; p.kc_img = tr.ks_img; p.vct_img = tr.vs_img; p.G = G; p.Vq = Vq; p.partial = partial;
  p.n_win = (int)n_win; p.way = way; p.N = tr.N; p.T = h->T; p.ldg = g_ld; p.voff = g_voff;
  p.trace = h->trace_buf;
  const bool mode0 = (h->T == 16 && tr.c == 2 && G != nullptr);
  if (!mode0 && !Vq) return arx_fail(h, ARX_ERR_INVALID, "tc_attention: generic epilogue needs Vq");
  if (episodes && !(mode0 && arx_tc_slot_order(h, tr) && (variant & 128) == 0))
    return arx_fail(h, ARX_ERR_INVALID, "tc_attention: episode mode needs the third-generation kernel");
  if (mode0 && arx_tc_slot_order(h, tr)) {
    // third-generation kernel by default; variant bit 7 (128) selects the second generation
    int rc = (variant & 128) ? arx_tc2_attention_launch(h, tr, kq_img, G, n_win, way, partial, g_ld, g_voff, st)
                             : arx_tc3_attention_launch(h, tr, kq_img, G, n_win, way, partial, g_ld, g_voff, g_chunked, episodes, st);
    if (rc) return rc;
    ARX_CUDA(h, arx_launch_pdl(k_finish_tc, dim3((unsigned)((n_win + 127) / 128)), dim3(128), 0, st, h->pdl, (const float *)partial, logits, chosen,
                               (int64_t)n_win, way, tr.N));
    h->launches++;
    return ARX_OK;
  }
  const int groups = (int)((n_win + GROUP - 1) / GROUP);
  const int grid = groups < h->sm_count ? groups : h->sm_count;
  const bool p_mn = (variant & 1) == 0;     // variant bit 0: use the K-major P layout (2-byte stores) instead of MN-major
  void (*kern)(const AttnParams) = mode0 ? (p_mn ? k_attn_tc<0, true> : k_attn_tc<0, false>) : (p_mn ? k_attn_tc<1, true> : k_attn_tc<1, false>);
  { const int rc_ = arx_func_smem(h, kern, (int)SMEM_BYTES); if (rc_) return rc_; }
  kern<<<grid, NTHREADS, SMEM_BYTES, st>>>(p);
  ARX_LAUNCH_CHECK(h);
  k_finish_tc<<<(unsigned)((n_win + 127) / 128), 128, 0, st>>>(partial, logits, chosen, n_win, way, tr.N);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}

bool arx_tc_head_supported(const arx_handle *h, const ArxTransformer &tr) {
  return arx_tc_supported(h, tr) && (h->T == 16 || h->T == 32) && h->cfg.has_discriminator;
}

int arx_tc_prep_head_weights(arx_handle *h, cudaStream_t st) {
  if (!h->cfg.has_discriminator || (h->T != 16 && h->T != 32)) return ARX_OK;
  if (!h->wdr_img) ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->wdr_img), (size_t)h->T * DD * 2));
  k_prep_wdr_img<<<1, 256, 0, st>>>(h->dr_w, h->wdr_img, h->T);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}

// y (n_win, N*T) fp32 = dimensionality_reduction(diff of the winning class), computed on tensor cores
int arx_tc_head_features(arx_handle *h, const ArxTransformer &tr, const __half *kq_img, const float *G, const float *Vq, int64_t n_win,
                         int way, const int32_t *chosen, float *y, __half *y_img, int y_nk, int g_ld, int g_voff, cudaStream_t st) {
  HeadParams p{};
  p.y_img = y_img; p.y_nk = y_nk;
  p.kq_img = kq_img; p.kc_img = tr.ks_img; p.vct_img = tr.vs_img; p.wdr_img = h->wdr_img; p.G = G; p.Vq = Vq; p.dr_b = h->dr_b;
  p.chosen = chosen; p.y = y; p.n_win = (int)n_win; p.way = way; p.N = tr.N; p.T = h->T; p.ldg = g_ld; p.voff = g_voff;
  p.L = h->T;
  const bool mode0 = (h->T == 16 && tr.c == 2 && G != nullptr);
  if (!mode0 && !Vq) return arx_fail(h, ARX_ERR_INVALID, "tc_head: generic epilogue needs Vq");
  void (*kern)(const HeadParams) = nullptr;
  const bool slot = mode0 && arx_tc_slot_order(h, tr);
  if (slot && !(h->dev_init & ARX_INIT_SLOT_RANK)) {
    short host[128];
    for (int q = 0; q < 128; ++q) {
      const int i = arx_slot_i(q), j = arx_slot_j(q);
      host[q] = (j == i) ? (short)-1 : (short)(i * (2 * 16 - i - 1) / 2 + (j - i - 1));
    }
    ARX_CUDA(h, cudaMemcpyToSymbol(c_slot_rank, host, sizeof(host)));
    h->dev_init |= ARX_INIT_SLOT_RANK;
  }
  if (h->T == 16) kern = mode0 ? (slot ? k_head_tc<0, 16, true> : k_head_tc<0, 16, false>) : k_head_tc<1, 16, false>;
  else if (h->T == 32) kern = k_head_tc<1, 32, false>;
  else return arx_fail(h, ARX_ERR_INVALID, "tc_head: unsupported seq_len");
  const int grid = n_win < h->sm_count ? (int)n_win : h->sm_count;
  { const int rc_ = arx_func_smem(h, kern, (int)H_SMEM_BYTES); if (rc_) return rc_; }
  kern<<<grid, NTHREADS, H_SMEM_BYTES, st>>>(p);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}

// One-launch support build from the per-frame projections G (way*T, 2cD); with_images also fills the tcgen05 operands.
int arx_tc_support_build(arx_handle *h, ArxTransformer &tr, const float *G, int way, bool with_images, cudaStream_t st) {
  if (with_images && !tr.ks_img) {
    ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tr.ks_img), (size_t)h->way_cap * IMG_BYTES));
    ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tr.vs_img), (size_t)h->way_cap * IMG_BYTES));
  }
  if (with_images && !tr.vs_img_bf) ARX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tr.vs_img_bf), (size_t)h->way_cap * IMG_BYTES));
  k_support_build<<<dim3(way, 8), 256, 0, st>>>(G, tr.tuples, tr.ln_g, tr.ln_b, tr.ks, tr.vs, with_images ? tr.ks_img : nullptr,
                                       with_images ? tr.vs_img : nullptr, with_images ? tr.vs_img_bf : nullptr, h->T, tr.c, tr.N,
                                       2 * tr.c * h->D);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}
