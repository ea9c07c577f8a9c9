// tcgen05 cross-attention + distance kernel, second generation, specialised for T=16 pair tuples (N=120):
// the BASELINE.json metric configuration.  Same math and operand formats as arx_tc.cu (see there for the
// reference citations and the MMA mapping); what changes is the schedule, driven by measurements on B200
// (tools/ubench, tools/trace_attn.py):
//   * MUFU.EX2 runs at 16 lanes/clk/SM, i.e. 1024 clk per 128x128 tile -- the same as the two MMAs.  One softmax
//     warp per SM sub-partition cannot hide its other ~400 instructions behind that, so TWO softmax warpgroups
//     alternate tiles: one group's loads / sums / scaling / stores run under the other group's MUFU stream.
//   * packed fp32x2 instructions (FADD2/FMUL2/FFMA2) halve the issue slots of the sums, the scaling and the
//     epilogue; registers are re-balanced between the roles with setmaxnreg.
//   * the 120 query tuples are laid out in a padded-triangular order of exactly 128 slots (row i of the (i,j)
//     triangle starts on an even j), so every fp32x2 operand pair is register-aligned: Vq[(i,j),(i,j+1)] =
//     {a_i,a_i} + {b_j,b_j+1}.  The order is internal to the kernel (the distance is a sum over tuples).
#include "arx_internal.cuh"
#include "arx_ptx.cuh"
#include <utility>

namespace {
using namespace ptx;

constexpr int TILE = 128;
constexpr int DD = 128;
constexpr uint32_t IMG_BYTES = TILE * DD * 2;
constexpr uint32_t SUB_BYTES = TILE * 64 * 2;
constexpr int GROUP = 2;
constexpr int NTHREADS2 = 512;

constexpr uint32_t OFF_KQ = 0;
constexpr uint32_t OFF_KC = 2 * IMG_BYTES;
constexpr uint32_t OFF_VCT = 4 * IMG_BYTES;
constexpr uint32_t OFF_P = 6 * IMG_BYTES;
constexpr uint32_t OFF_BAR = 7 * IMG_BYTES;
enum { B_FULL_KQ = 0, B_EMPTY_KQ = 2, B_FULL_KC = 4, B_EMPTY_KC = 6, B_S_FULL = 8, B_S_EMPTY = 10, B_P_FULL = 12, B_P_EMPTY = 13,
       B_O_FULL = 14, B_O_EMPTY = 16, B_FULL_VC = 18, B_EMPTY_VC = 20, B_P_EMPTY1 = 22, B_COUNT = 23 };
constexpr uint32_t SMEM_BYTES = OFF_BAR + B_COUNT * 8 + 16 + 1024;

struct Attn2Params {
  const __half *kq_img, *kc_img, *vct_img;
  const float *G;
  float *partial;
  int n_win, way, ldg, voff;
  long long *trace;
};
#define ARX_TRACE_TILES 64
#define TRACE2(role, tile, k) do { if (p.trace && blockIdx.x == 0 && (tile) < ARX_TRACE_TILES) p.trace[(((role) * ARX_TRACE_TILES) + (tile)) * 8 + (k)] = clock64(); } while (0)

struct TileIter {
  int n_win, way, n_groups, gstride, group, gi, c, w, nw;
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
  __device__ void next_group() { c = way - 1; w = nw - 1; next(); }
  __device__ int window() const { return group * GROUP + w; }
  __device__ int cls_counter() const { return gi * way + c; }
};

}  // namespace

namespace {

__device__ __forceinline__ uint32_t ex2_bits(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}

// one fp32x2 pair of the epilogue: columns Q, Q+1 of chunk registers r (Q even)
template <int Q> __device__ __forceinline__ void epi_pair(const float (&a)[16], const uint64_t (&bb)[8], const uint32_t (&r)[32], uint64_t &acc) {
  constexpr int I = arx_slot_i(Q), J = arx_slot_j(Q);
  static_assert(J % 2 == 0, "pairs start on an even j");
  uint64_t bj;
  if constexpr (J == I) bj = pack2(-a[I], __uint_as_float((uint32_t)(bb[J / 2] >> 32)));   // pad lane: a_i + (-a_i) == 0 == proto
  else bj = bb[J / 2];
  const uint64_t v = add2(pack2(a[I], a[I]), bj);
  const uint64_t d = sub2(v, pack2u(r[Q & 31], r[(Q & 31) + 1]));
  acc = fma2(d, d, acc);
}
template <int CH, int... Ks>
__device__ __forceinline__ void epi_chunk(const float (&a)[16], const uint64_t (&bb)[8], const uint32_t (&r)[32], uint64_t &acc,
                                          std::integer_sequence<int, Ks...>) {
  (epi_pair<CH * 32 + 2 * Ks>(a, bb, r, acc), ...);
}

template <int... Is> __device__ __forceinline__ void zero_pads(uint32_t (&r)[128], std::integer_sequence<int, Is...>) {
  ((r[arx_slot_row_start(2 * Is)] = 0u), ...);     // the pad slot of every even row
}

__global__ void __launch_bounds__(NTHREADS2, 1) k_attn_tc2(const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + B_COUNT * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_FULL_KQ + i], 1); mbar_init(&bars[B_EMPTY_KQ + i], 1);
      mbar_init(&bars[B_FULL_KC + i], 1); mbar_init(&bars[B_EMPTY_KC + i], 1);
      mbar_init(&bars[B_FULL_VC + i], 1); mbar_init(&bars[B_EMPTY_VC + i], 1);
      mbar_init(&bars[B_S_FULL + i], 1); mbar_init(&bars[B_S_EMPTY + i], 128);
      mbar_init(&bars[B_O_FULL + i], 1); mbar_init(&bars[B_O_EMPTY + i], 128);
    }
    mbar_init(&bars[B_P_FULL], 128); mbar_init(&bars[B_P_EMPTY], 1); mbar_init(&bars[B_P_EMPTY1], 1);
    mbar_init_fence();
  }
  if (warp == 3) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_S = tmem, TM_O = tmem + 256;

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      if (elect_one()) {            // producer: class operands, one stage per class, reused by the group's windows
        TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
        int cl = 0;
        while (it.valid) {
          for (int c = 0; c < p.way; ++c, ++cl) {
            const int st = cl & 1;
            const uint8_t *kc = reinterpret_cast<const uint8_t *>(p.kc_img) + (size_t)c * IMG_BYTES;
            const uint8_t *vc = reinterpret_cast<const uint8_t *>(p.vct_img) + (size_t)c * IMG_BYTES;
            mbar_wait(&bars[B_EMPTY_KC + st], ((cl >> 1) & 1) ^ 1);        // free once the class's last S tile is computed
            mbar_arrive_expect_tx(&bars[B_FULL_KC + st], IMG_BYTES);
            bulk_g2s(smem + OFF_KC + st * IMG_BYTES, kc, SUB_BYTES, &bars[B_FULL_KC + st]);
            bulk_g2s(smem + OFF_KC + st * IMG_BYTES + SUB_BYTES, kc + SUB_BYTES, SUB_BYTES, &bars[B_FULL_KC + st]);
            mbar_wait(&bars[B_EMPTY_VC + st], ((cl >> 1) & 1) ^ 1);        // free once the class's last prototype tile is computed
            mbar_arrive_expect_tx(&bars[B_FULL_VC + st], IMG_BYTES);
            bulk_g2s(smem + OFF_VCT + st * IMG_BYTES, vc, SUB_BYTES, &bars[B_FULL_VC + st]);
            bulk_g2s(smem + OFF_VCT + st * IMG_BYTES + SUB_BYTES, vc + SUB_BYTES, SUB_BYTES, &bars[B_FULL_VC + st]);
          }
          it.next_group();
        }
      }
    } else if (warp == 2) {
      if (elect_one()) {            // producer: Kq images, one slot per window of the group
        TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
        while (it.valid) {
          for (int w = 0; w < it.nw; ++w) {
            mbar_wait(&bars[B_EMPTY_KQ + w], (it.gi & 1) ^ 1);
            mbar_arrive_expect_tx(&bars[B_FULL_KQ + w], IMG_BYTES);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(p.kq_img) + (size_t)(it.group * GROUP + w) * IMG_BYTES;
            bulk_g2s(smem + OFF_KQ + w * IMG_BYTES, src, SUB_BYTES, &bars[B_FULL_KQ + w]);
            bulk_g2s(smem + OFF_KQ + w * IMG_BYTES + SUB_BYTES, src + SUB_BYTES, SUB_BYTES, &bars[B_FULL_KQ + w]);
          }
          it.next_group();
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {            // MMA1 issuer: S^T tiles, as far ahead as the two S buffers allow
        constexpr uint64_t DESC_K = smem_desc_sw128(16, 1024);
        constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);
        const uint32_t sbase = smem_u32(smem);
        TileIter it1; it1.init(p.n_win, p.way, blockIdx.x, gridDim.x);
        for (int f1 = 0; it1.valid; ++f1, it1.next()) {
          const int cc = it1.cls_counter(), st = cc & 1, buf = f1 & 1;
          if (it1.c == 0) mbar_wait(&bars[B_FULL_KQ + it1.w], it1.gi & 1);
          if (it1.w == 0) mbar_wait(&bars[B_FULL_KC + st], (cc >> 1) & 1);
          mbar_wait(&bars[B_S_EMPTY + buf], ((f1 >> 1) & 1) ^ 1);
          tc_fence_after();
          TRACE2(0, f1, 0);
          const uint32_t a0 = sbase + OFF_KC + st * IMG_BYTES, b0 = sbase + OFF_KQ + it1.w * IMG_BYTES;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
            mma_f16_ss(TM_S + buf * 128, smem_desc_at(DESC_K, a0 + off), smem_desc_at(DESC_K, b0 + off), IDESC1, kk > 0);
          }
          mma_commit(&bars[B_S_FULL + buf]);
          if (it1.c == p.way - 1) mma_commit(&bars[B_EMPTY_KQ + it1.w]);
          if (it1.w == it1.nw - 1) mma_commit(&bars[B_EMPTY_KC + st]);
        }
      }
    } else {
      if (elect_one()) {            // MMA2 issuer (its own thread, so a prototype MMA never queues behind a look-ahead S^T MMA)
        constexpr uint64_t DESC_K = smem_desc_sw128(16, 1024);
        constexpr uint64_t DESC_MN = smem_desc_sw128(16384, 1024);
        constexpr uint32_t IDESC2 = idesc_bf16(128, 128, 0, 1);      // Vc^T and P are bf16 (see the softmax warps)
        const uint32_t sbase = smem_u32(smem);
        TileIter it2; it2.init(p.n_win, p.way, blockIdx.x, gridDim.x);
        for (int f2 = 0; it2.valid; ++f2, it2.next()) {
          const int cc = it2.cls_counter(), st = cc & 1, buf = f2 & 1;
          TRACE2(0, f2, 1);
          if (it2.w == 0) mbar_wait(&bars[B_FULL_VC + st], (cc >> 1) & 1);
          mbar_wait(&bars[B_P_FULL], f2 & 1);
          TRACE2(0, f2, 2);
          mbar_wait(&bars[B_O_EMPTY + buf], ((f2 >> 1) & 1) ^ 1);
          tc_fence_after();
          TRACE2(0, f2, 3);
          const uint32_t a0 = sbase + OFF_VCT + st * IMG_BYTES, b0 = sbase + OFF_P;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
            mma_f16_ss(TM_O + buf * 128, smem_desc_at(DESC_K, a0 + off), smem_desc_at(DESC_MN, b0 + kk * 2048), IDESC2, kk > 0);
          }
          mma_commit(&bars[B_O_FULL + buf]);
          // the P buffer goes to the group that owns tile f2+1; one barrier per group so that every waiter sees
          // consecutive phases (a parity wait cannot tell phase k from phase k+2)
          mma_commit(&bars[((f2 + 1) & 1) ? B_P_EMPTY1 : B_P_EMPTY]);
          if (it2.w == it2.nw - 1) mma_commit(&bars[B_EMPTY_VC + st]);
        }
      }
    }
  } else if (warp < 12) {
    // ---------------- two softmax warpgroups; group g owns the tiles with (f & 1) == g and the S buffer g
    setmaxnreg_inc<160>();
    const int g = (warp - 4) >> 2, quad = warp & 3;
    const int s = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint8_t *prow = smem + OFF_P + (s >> 3) * 1024 + (s & 7) * 128;
    const bool tr = (threadIdx.x == 128 + g * 128);
    TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
    for (int f = 0; it.valid; ++f, it.next()) {
      if ((f & 1) != g) continue;
      if (tr) TRACE2(1, f, 0);
      mbar_wait(&bars[B_S_FULL + g], (f >> 1) & 1);
      tc_fence_after();
      if (tr) TRACE2(1, f, 1);
      uint32_t r[128];
      tmem_ld32(TM_S + lane_base + g * 128 + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tmem_ld32(TM_S + lane_base + g * 128 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_ld32(TM_S + lane_base + g * 128 + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
      tmem_ld32(TM_S + lane_base + g * 128 + 96, *reinterpret_cast<uint32_t(*)[32]>(&r[96]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars[B_S_EMPTY + g]);
      if (tr) TRACE2(1, f, 2);
#pragma unroll
      for (int j = 0; j < 128; ++j) r[j] = ex2_bits(r[j]);
      zero_pads(r, std::make_integer_sequence<int, 8>{});          // pad slots: E = 0 -> P = 0 -> proto = 0
      uint64_t z0 = 0ull, z1 = 0ull;
#pragma unroll
      for (int k = 0; k < 64; k += 2) {
        z0 = add2(z0, pack2u(r[2 * k], r[2 * k + 1]));
        z1 = add2(z1, pack2u(r[2 * k + 2], r[2 * k + 3]));
      }
      float zl, zh;
      unpack2(add2(z0, z1), zl, zh);
      // P = E / Z as bf16 WITHOUT conversion instructions (F2FP runs at quarter rate on the same XU pipe as MUFU):
      // scale by zinv*(1 + c) and truncate with a byte permute.  Truncation loses u*ulp, u ~ U[0,1); relative to a
      // log-uniform mantissa its mean is ulp * (1/(2 ln 2)) = 0.7213 * 2^-8, so c = 0.7213 * 2^-8 centres the error:
      // same variance as round-to-nearest, no bias.
      const float zinv = __frcp_rn(zl + zh) * 1.0028177f;
      const uint64_t zz = pack2(zinv, zinv);
      if (tr) TRACE2(1, f, 3);
      if (g == 0) mbar_wait(&bars[B_P_EMPTY], ((f >> 1) & 1) ^ 1);   // tile 0 finds the buffer free
      else mbar_wait(&bars[B_P_EMPTY1], (f >> 1) & 1);
      if (tr) TRACE2(1, f, 4);
#pragma unroll
      for (int c16 = 0; c16 < 16; ++c16) {
        uint32_t h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t m = mul2(pack2u(r[c16 * 8 + 2 * k], r[c16 * 8 + 2 * k + 1]), zz);
          h[k] = __byte_perm((uint32_t)m, (uint32_t)(m >> 32), 0x7632);     // {bf16(lo), bf16(hi)}
        }
        *reinterpret_cast<uint4 *>(prow + (c16 >> 3) * 16384 + (((c16 & 7) ^ (s & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
      }
      if (tr) TRACE2(1, f, 5);
      fence_proxy_async_smem();
      mbar_arrive(&bars[B_P_FULL]);
      if (tr) TRACE2(1, f, 6);
    }
  } else {
    // ---------------- epilogue warps: thread == output dimension d == TMEM lane
    setmaxnreg_inc<152>();
    const int quad = warp & 3;
    const int d = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const bool tr = (threadIdx.x == 384);
    TileIter it; it.init(p.n_win, p.way, blockIdx.x, gridDim.x);
    float a0[16], a1[16];
    uint64_t bb0[8], bb1[8];
    for (int f = 0; it.valid; ++f, it.next()) {
      if (it.c == 0 && it.w == 0) {
        // per-frame V projections of the group's windows (v bias already folded into part 0)
        const float *g0 = p.G + (size_t)(it.group * GROUP) * 16 * p.ldg + p.voff + d;
#pragma unroll
        for (int i = 0; i < 16; ++i) a0[i] = __ldg(g0 + (size_t)i * p.ldg);
#pragma unroll
        for (int m = 0; m < 8; ++m) bb0[m] = pack2(__ldg(g0 + (size_t)(2 * m) * p.ldg + DD), __ldg(g0 + (size_t)(2 * m + 1) * p.ldg + DD));
        if (it.nw > 1) {
          const float *g1 = g0 + (size_t)16 * p.ldg;
#pragma unroll
          for (int i = 0; i < 16; ++i) a1[i] = __ldg(g1 + (size_t)i * p.ldg);
#pragma unroll
          for (int m = 0; m < 8; ++m) bb1[m] = pack2(__ldg(g1 + (size_t)(2 * m) * p.ldg + DD), __ldg(g1 + (size_t)(2 * m + 1) * p.ldg + DD));
        }
      }
      const int buf = f & 1;
      if (tr) TRACE2(2, f, 0);
      mbar_wait(&bars[B_O_FULL + buf], (f >> 1) & 1);
      tc_fence_after();
      if (tr) TRACE2(2, f, 1);
      uint64_t acc2 = 0ull;
      auto run = [&](const float (&a)[16], const uint64_t (&bb)[8]) {
        uint32_t r[32], r2[32];
        tmem_ld32(TM_O + lane_base + buf * 128 + 0, r);
        tmem_ld32(TM_O + lane_base + buf * 128 + 32, r2);
        tmem_ld_wait();
        epi_chunk<0>(a, bb, r, acc2, std::make_integer_sequence<int, 16>{});
        tmem_ld32(TM_O + lane_base + buf * 128 + 64, r);
        epi_chunk<1>(a, bb, r2, acc2, std::make_integer_sequence<int, 16>{});
        tmem_ld_wait();
        tmem_ld32(TM_O + lane_base + buf * 128 + 96, r2);
        epi_chunk<2>(a, bb, r, acc2, std::make_integer_sequence<int, 16>{});
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars[B_O_EMPTY + buf]);
        epi_chunk<3>(a, bb, r2, acc2, std::make_integer_sequence<int, 16>{});
      };
      if (it.w == 0) run(a0, bb0); else run(a1, bb1);
      float al, ah;
      unpack2(acc2, al, ah);
      float acc = al + ah;
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) p.partial[((size_t)it.window() * p.way + it.c) * 4 + quad] = acc;
      if (tr) TRACE2(2, f, 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

// Host-side slot table (128 x 2): frame indices of every slot, -1 for pads; index math only.
void arx_tc2_slot_table(int32_t *out /* 256 */) {
  for (int q = 0; q < 128; ++q) {
    const int i = arx_slot_i(q), j = arx_slot_j(q);
    out[2 * q] = (j == i) ? -1 : i;
    out[2 * q + 1] = (j == i) ? -1 : j;
  }
}

int arx_tc2_attention_launch(arx_handle *h, const ArxTransformer &tr, const __half *kq_img, const float *G, int64_t n_win, int way,
                             float *partial, int g_ld, int g_voff, cudaStream_t st) {
  Attn2Params p{};
  p.kq_img = kq_img; p.kc_img = tr.ks_img; p.vct_img = tr.vs_img_bf; p.G = G; p.partial = partial;
  p.n_win = (int)n_win; p.way = way; p.ldg = g_ld; p.voff = g_voff; p.trace = h->trace_buf;
  const int groups = (int)((n_win + GROUP - 1) / GROUP);
  const int grid = groups < h->sm_count ? groups : h->sm_count;
  ARX_CUDA(h, cudaFuncSetAttribute(k_attn_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  k_attn_tc2<<<grid, NTHREADS2, SMEM_BYTES, st>>>(p);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}
