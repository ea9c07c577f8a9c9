// fp32 CUDA-core kernels: the general-shape path (any T, cardinality, way) and the on-device
// cross-check for the tcgen05 kernels.  Plain IEEE fp32 FMA arithmetic, deterministic (no atomics).
//
// Reference semantics restated (paths relative to the reference root):
//   linear/bias/ReLU/PE     modules/ar/utils/model.py:164-180, 26-28
//   tuple gather + LN       model.py:69-72, 75-84
//   cross attention         model.py:95-126  (softmax over the QUERY-tuple axis, dim=-2)
//   distances / logits      model.py:130-146
//   argmax + head selection model.py:323-324
#include "arx_internal.cuh"
#include <math.h>

namespace {

constexpr int TM = 64, TN = 64, TK = 16, TPAD = 4;

// acc[i][j] += sum_k A[m0+ty*4+i][k] * B[n0+tx*4+j][k]; A (M rows, lda), B (Nr rows, ldb), both K-contiguous.
__device__ __forceinline__ void tile_mainloop(const float *__restrict__ A, int lda, int64_t m0, int64_t M,
                                              const float *__restrict__ B, int ldb, int64_t n0, int64_t Nr,
                                              int K, float (&acc)[4][4], float (*As)[TM + TPAD],
                                              float (*Bs)[TN + TPAD]) {
  const int tid = threadIdx.x;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t arow = m0 + lrow, brow = n0 + lrow;
  const float *ap = A + arow * (int64_t)lda;
  const float *bp = B + brow * (int64_t)ldb;
  const bool aok = arow < M, bok = brow < Nr;
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int k = k0 + lk + e;
      As[lk + e][lrow] = (aok && k < K) ? __ldg(ap + k) : 0.f;
      Bs[lk + e][lrow] = (bok && k < K) ? __ldg(bp + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_linear(const float *__restrict__ A, int lda, const float *__restrict__ W,
                                                int ldw, const float *__restrict__ bias, float *__restrict__ C,
                                                int ldc, int64_t M, int N, int K, int act,
                                                const float *__restrict__ pe, int peT) {
  __shared__ __align__(16) float As[TK][TM + TPAD];
  __shared__ __align__(16) float Bs[TK][TN + TPAD];
  const int64_t m0 = (int64_t)blockIdx.y * TM;
  const int64_t n0 = (int64_t)blockIdx.x * TN;
  float acc[4][4] = {};
  tile_mainloop(A, lda, m0, M, W, ldw, n0, N, K, acc, As, Bs);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = (int)n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (act == ARX_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == ARX_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      if (pe) v += pe[(int)(m % peT) * (int64_t)N + n];
      C[m * (int64_t)ldc + n] = v;
    }
  }
}

// One warp per (sequence, tuple).  K = LayerNorm(sum_p Gk_p[frame_p] ), V = sum_p Gv_p[frame_p]
// (biases already folded into part 0 by the projection).  D == 128: 4 floats per lane.
__global__ void __launch_bounds__(256) k_build_tuples(const float *__restrict__ G, const int32_t *__restrict__ tuples,
                                                      const float *__restrict__ ln_g, const float *__restrict__ ln_b,
                                                      float *__restrict__ Ko, float *__restrict__ Vo, int64_t n_seq,
                                                      int T, int c, int N, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= n_seq * N) return;
  const int64_t seq = wid / N;
  const int tup = (int)(wid % N);
  const int ld = 2 * c * D;
  for (int d0 = lane * 4; d0 < D; d0 += 128) {  // D == 128 -> one iteration
    float4 k = make_float4(0, 0, 0, 0), v = make_float4(0, 0, 0, 0);
    for (int p = 0; p < c; ++p) {
      int f = tuples[tup * c + p];
      const float *row = G + (seq * T + f) * (int64_t)ld;
      float4 a = *reinterpret_cast<const float4 *>(row + p * D + d0);
      float4 b = *reinterpret_cast<const float4 *>(row + (c + p) * D + d0);
      k.x += a.x; k.y += a.y; k.z += a.z; k.w += a.w;
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    // LayerNorm over D (eps 1e-5, biased variance), two-pass in registers
    float s = k.x + k.y + k.z + k.w;
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / D;
    float4 dlt = make_float4(k.x - mean, k.y - mean, k.z - mean, k.w - mean);
    float q = dlt.x * dlt.x + dlt.y * dlt.y + dlt.z * dlt.z + dlt.w * dlt.w;
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / D + 1e-5f);
    float4 g = *reinterpret_cast<const float4 *>(ln_g + d0);
    float4 be = *reinterpret_cast<const float4 *>(ln_b + d0);
    float4 o4 = make_float4(dlt.x * rstd * g.x + be.x, dlt.y * rstd * g.y + be.y, dlt.z * rstd * g.z + be.z,
                            dlt.w * rstd * g.w + be.w);
    *reinterpret_cast<float4 *>(Ko + wid * D + d0) = o4;
    *reinterpret_cast<float4 *>(Vo + wid * D + d0) = v;
  }
}

// Pass A: per support tuple s of class c, online (max, sum) over ALL query tuples q of window b:
//   M[s] = max_q S[q,s],  Z[s] = sum_q exp(S[q,s] - M[s]),  S = Kq.Ks^T / sqrt(D)     (model.py:101-109)
__global__ void __launch_bounds__(256) k_colstats(const float *__restrict__ Kq, const float *__restrict__ Ks,
                                                  float *__restrict__ Zo, int N, int D, int way, float scale,
                                                  const int32_t *__restrict__ chosen) {
  __shared__ __align__(16) float As[TK][TM + TPAD];
  __shared__ __align__(16) float Bs[TK][TN + TPAD];
  const int sb = blockIdx.x;
  const int64_t b = blockIdx.z;
  const int c = chosen ? chosen[b] : blockIdx.y;
  const float *ks = Ks + (int64_t)c * N * D;
  const float *kq = Kq + b * (int64_t)N * D;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float mx[4], zs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { mx[i] = -INFINITY; zs[i] = 0.f; }
  for (int q0 = 0; q0 < N; q0 += TN) {
    float acc[4][4] = {};
    tile_mainloop(ks, D, (int64_t)sb * TM, N, kq, D, q0, N, D, acc, As, Bs);   // acc[i][j] = S^T[s_i][q_j]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (q0 + tx * 4 + j < N) tmax = fmaxf(tmax, acc[i][j] * scale);
#pragma unroll
      for (int o = 8; o; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      float nm = fmaxf(mx[i], tmax);
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (q0 + tx * 4 + j < N) part += expf(acc[i][j] * scale - nm);
#pragma unroll
      for (int o = 8; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      zs[i] = zs[i] * expf(mx[i] - nm) + part;
      mx[i] = nm;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int s = sb * TM + ty * 4 + i;
      if (s < N) {
        float *z = Zo + ((b * way + c) * (int64_t)N + s) * 2;
        z[0] = mx[i];
        z[1] = zs[i];
      }
    }
  }
}

// Pass B: for a block of 64 query tuples: proto = P.Vs, diff = Vq - proto, partial squared distance
// (model.py:125-132).  chosen != nullptr -> only class chosen[b] and emit y = diff.Wdr^T + bdr (model.py:196).
constexpr int PS_LD = TN + 1;
constexpr int VS_LD = 128;
constexpr int DF_LD = 129;
__global__ void __launch_bounds__(256) k_attend(const float *__restrict__ Kq, const float *__restrict__ Vq,
                                                const float *__restrict__ Ks, const float *__restrict__ Vs,
                                                const float *__restrict__ Zi, float *__restrict__ partial,
                                                const int32_t *__restrict__ chosen, float *__restrict__ y,
                                                const float *__restrict__ dr_w, const float *__restrict__ dr_b, int L,
                                                float *__restrict__ probs, float *__restrict__ protos, int N, int D,
                                                int way, float scale) {
  extern __shared__ __align__(16) float smem[];
  float(*As)[TM + TPAD] = reinterpret_cast<float(*)[TM + TPAD]>(smem);
  float(*Bs)[TN + TPAD] = reinterpret_cast<float(*)[TN + TPAD]>(smem + TK * (TM + TPAD));
  float *Ps = smem + 2 * TK * (TM + TPAD);
  float *Vt = Ps + TM * PS_LD;                      // 64 x 128 (also reused as diff, 64 x 129)
  __shared__ float red[8];
  const int qb = blockIdx.x;
  const int64_t b = blockIdx.z;
  const int c = chosen ? chosen[b] : blockIdx.y;
  const int nqb = gridDim.x;
  const float *ks = Ks + (int64_t)c * N * D;
  const float *vs = Vs + (int64_t)c * N * D;
  const float *kq = Kq + b * (int64_t)N * D;
  const float *zi = Zi + (b * way + c) * (int64_t)N * 2;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc2[4][8] = {};
  for (int s0 = 0; s0 < N; s0 += TN) {
    float acc[4][4] = {};
    tile_mainloop(kq, D, (int64_t)qb * TM, N, ks, D, s0, N, D, acc, As, Bs);   // acc[i][j] = S[q_i][s_j]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int s = s0 + tx * 4 + j;
      float m = 0.f, rz = 0.f;
      if (s < N) { m = zi[2 * s]; rz = 1.0f / zi[2 * s + 1]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float p = (s < N) ? expf(acc[i][j] * scale - m) * rz : 0.f;
        Ps[(ty * 4 + i) * PS_LD + tx * 4 + j] = p;
        int q = qb * TM + ty * 4 + i;
        if (probs && s < N && q < N) probs[((b * way + c) * (int64_t)N + q) * N + s] = p;
      }
    }
    for (int e = tid; e < TN * (VS_LD / 4); e += 256) {
      int s = e / (VS_LD / 4), d4 = e % (VS_LD / 4);
      float4 v = make_float4(0, 0, 0, 0);
      if (s0 + s < N) v = *reinterpret_cast<const float4 *>(vs + (int64_t)(s0 + s) * D + d4 * 4);
      *reinterpret_cast<float4 *>(Vt + s * VS_LD + d4 * 4) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int s = 0; s < TN; ++s) {
      float4 v0 = *reinterpret_cast<const float4 *>(Vt + s * VS_LD + tx * 4);
      float4 v1 = *reinterpret_cast<const float4 *>(Vt + s * VS_LD + 64 + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float p = Ps[(ty * 4 + i) * PS_LD + s];
        acc2[i][0] = fmaf(p, v0.x, acc2[i][0]); acc2[i][1] = fmaf(p, v0.y, acc2[i][1]);
        acc2[i][2] = fmaf(p, v0.z, acc2[i][2]); acc2[i][3] = fmaf(p, v0.w, acc2[i][3]);
        acc2[i][4] = fmaf(p, v1.x, acc2[i][4]); acc2[i][5] = fmaf(p, v1.y, acc2[i][5]);
        acc2[i][6] = fmaf(p, v1.z, acc2[i][6]); acc2[i][7] = fmaf(p, v1.w, acc2[i][7]);
      }
    }
    __syncthreads();
  }
  // epilogue: diff and squared distance
  float dist = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int q = qb * TM + ty * 4 + i;
    float df[8] = {};
    if (q < N) {
      const float *vq = Vq + (b * (int64_t)N + q) * D;
      float4 a0 = *reinterpret_cast<const float4 *>(vq + tx * 4);
      float4 a1 = *reinterpret_cast<const float4 *>(vq + 64 + tx * 4);
      df[0] = a0.x - acc2[i][0]; df[1] = a0.y - acc2[i][1]; df[2] = a0.z - acc2[i][2]; df[3] = a0.w - acc2[i][3];
      df[4] = a1.x - acc2[i][4]; df[5] = a1.y - acc2[i][5]; df[6] = a1.z - acc2[i][6]; df[7] = a1.w - acc2[i][7];
#pragma unroll
      for (int e = 0; e < 8; ++e) dist = fmaf(df[e], df[e], dist);
      if (protos) {
        float *pr = protos + ((b * way + c) * (int64_t)N + q) * D;
        *reinterpret_cast<float4 *>(pr + tx * 4) = make_float4(acc2[i][0], acc2[i][1], acc2[i][2], acc2[i][3]);
        *reinterpret_cast<float4 *>(pr + 64 + tx * 4) = make_float4(acc2[i][4], acc2[i][5], acc2[i][6], acc2[i][7]);
      }
    }
    if (y) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        Vt[(ty * 4 + i) * DF_LD + tx * 4 + e] = df[e];
        Vt[(ty * 4 + i) * DF_LD + 64 + tx * 4 + e] = df[4 + e];
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, o);
  if ((tid & 31) == 0) red[tid >> 5] = dist;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    if (!chosen) partial[(b * way + c) * (int64_t)nqb + qb] = t;
  }
  if (y) {
    // y[b][q*L + l] = diff[q,:] . dr_w[l,:] + dr_b[l]
    const int ql = tid >> 2;
    const int q = qb * TM + ql;
    if (q < N) {
      for (int l = (tid & 3); l < L; l += 4) {
        const float *w = dr_w + (int64_t)l * D;
        float a = 0.f;
        for (int d = 0; d < D; ++d) a = fmaf(Vt[ql * DF_LD + d], __ldg(w + d), a);
        y[b * (int64_t)N * L + (int64_t)q * L + l] = a + dr_b[l];
      }
    }
  }
}

// logits[b][c] = -(sum of partials)/N (model.py:131-135); chosen = first argmax (model.py:323)
__global__ void k_finish(const float *__restrict__ partial, float *__restrict__ logits, int32_t *__restrict__ chosen,
                         int64_t n_win, int way, int nqb, int N) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_win) return;
  float best = -INFINITY;
  int bi = 0;
  for (int c = 0; c < way; ++c) {
    float t = 0.f;
    for (int k = 0; k < nqb; ++k) t += partial[(b * way + c) * (int64_t)nqb + k];
    float lg = -(t / (float)N);
    logits[b * way + c] = lg;
    if (lg > best) { best = lg; bi = c; }
  }
  if (chosen) chosen[b] = bi;
}

}  // namespace

int arx_fp32_linear(arx_handle *h, const float *A, int lda, const float *W, int ldw, const float *bias, float *C,
                    int ldc, int64_t M, int N, int K, int act, const float *pe, int peT, cudaStream_t st) {
  if (M <= 0) return ARX_OK;
  const int64_t max_y = 65535;
  for (int64_t m0 = 0; m0 < M; m0 += max_y * TM) {
    int64_t mm = M - m0 < max_y * TM ? M - m0 : max_y * TM;
    dim3 grid((N + TN - 1) / TN, (unsigned)((mm + TM - 1) / TM));
    // pe row index uses the GLOBAL row (m0 is a multiple of TM; peT divides into rows independently)
    const float *pe_adj = pe;
    if (pe && (m0 % peT) != 0) return arx_fail(h, ARX_ERR_INVALID, "linear: chunk not aligned to seq_len");
    k_linear<<<grid, 256, 0, st>>>(A + m0 * lda, lda, W, ldw, bias, C + m0 * ldc, ldc, mm, N, K, act, pe_adj, peT);
    ARX_LAUNCH_CHECK(h);
  }
  return ARX_OK;
}

int arx_fp32_build_tuples(arx_handle *h, const ArxTransformer &tr, const float *G, int64_t n_seq, float *K, float *V,
                          cudaStream_t st) {
  if (h->D != 128) return arx_fail(h, ARX_ERR_INVALID, "build_tuples: out_dim must be 128");
  int64_t warps = n_seq * tr.N;
  if (warps == 0) return ARX_OK;
  int64_t blocks = (warps + 7) / 8;
  k_build_tuples<<<(unsigned)blocks, 256, 0, st>>>(G, tr.tuples, tr.ln_g, tr.ln_b, K, V, n_seq, h->T, tr.c, tr.N,
                                                   h->D);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}

int arx_fp32_attention(arx_handle *h, const ArxTransformer &tr, const float *Kq, const float *Vq, int64_t n_win,
                       int way, float *Z, float *partial, float *logits, int32_t *chosen, float *y, float *probs,
                       float *protos, cudaStream_t st) {
  const int N = tr.N, D = h->D;
  const int nb = (N + TM - 1) / TM;
  const float scale = 1.0f / sqrtf((float)D);
  const size_t smem = (size_t)(2 * TK * (TM + TPAD) + TM * PS_LD + TM * DF_LD) * sizeof(float);
  { const int rc_ = arx_func_smem(h, k_attend, (int)smem); if (rc_) return rc_; }      // cached per handle (per device)
  for (int64_t b0 = 0; b0 < n_win; b0 += 65535) {
    int64_t nb_win = n_win - b0 < 65535 ? n_win - b0 : 65535;
    for (int c0 = 0; c0 < way; c0 += 65535) {  // way never exceeds this; kept for form
      dim3 gA(nb, way, (unsigned)nb_win);
      k_colstats<<<gA, 256, 0, st>>>(Kq + b0 * N * D, tr.ks, Z + b0 * way * N * 2, N, D, way, scale, nullptr);
      ARX_LAUNCH_CHECK(h);
      k_attend<<<gA, 256, smem, st>>>(Kq + b0 * N * D, Vq + b0 * N * D, tr.ks, tr.vs, Z + b0 * way * N * 2,
                                      partial + b0 * way * nb, nullptr, nullptr, nullptr, nullptr, 0,
                                      probs ? probs + b0 * way * N * N : nullptr,
                                      protos ? protos + b0 * way * N * D : nullptr, N, D, way, scale);
      ARX_LAUNCH_CHECK(h);
    }
  }
  k_finish<<<(unsigned)((n_win + 127) / 128), 128, 0, st>>>(partial, logits, chosen, n_win, way, nb, N);
  ARX_LAUNCH_CHECK(h);
  if (y) {
    if (!chosen) return arx_fail(h, ARX_ERR_INVALID, "attention: y requires chosen");
    for (int64_t b0 = 0; b0 < n_win; b0 += 65535) {
      int64_t nb_win = n_win - b0 < 65535 ? n_win - b0 : 65535;
      dim3 gB(nb, 1, (unsigned)nb_win);
      k_attend<<<gB, 256, smem, st>>>(Kq + b0 * N * D, Vq + b0 * N * D, tr.ks, tr.vs, Z + b0 * way * N * 2, nullptr,
                                      chosen + b0, y + b0 * N * h->T, h->dr_w, h->dr_b, h->T, nullptr, nullptr, N, D,
                                      way, scale);
      ARX_LAUNCH_CHECK(h);
    }
  }
  return ARX_OK;
}

// Open-set head input for an already-known winning class (model.py:323-324,196): softmax statistics and the
// attention pass for class chosen[b] only, emitting y = diff.Wdr^T + bdr.
