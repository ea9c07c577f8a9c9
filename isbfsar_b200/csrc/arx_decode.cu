// MetrABS-style heatmap decoder (kernel 5): soft-argmax over the 8x8(x8) logits, FOV test,
// absolute reconstruction, homography undo, 32->n_out joint remap, root-centring.
// HBM-bound: 73,728 B in / 360 B out per frame; one WARP per frame, lane == joint.
//
// Reference restated (paths relative to the reference root):
//   modules/hpe/hpe.py:108-146   softmax + soft-argmax (x<-w, y<-h, z<-d; 2-D head x255)
//   modules/hpe/hpe.py:149-153   FOV test, frame dropped when < 1/4 of joints are inside
//   modules/hpe/utils/misc.py:141-208  reconstruct_absolute / reconstruct_ref_fullpersp / back_project
//   modules/hpe/hpe.py:159-169   @ homo_inv, @ expand_joints (column-selected), joint subset
//   main.py:103-105              pose -= pose[0]; flatten
// The reference runs the tail in float64 (numpy promotion); here the soft-argmax accumulates in
// fp32 and the 3x3 weighted least squares is solved in fp64 from its normal equations.
#include "arx_internal.cuh"
#include <math.h>

namespace {

constexpr int NJ = 32;     // head joints
constexpr int ND = 8;      // depth bins
constexpr int CH = NJ * (1 + ND);   // 288 channels per (h,w) cell

struct DecodeParams {
  float invK[9];   // inverse intrinsics (fp32, as np.linalg.inv of a float32 matrix)
  double R[9];     // homo_inv
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ONE WARP PER FRAME, lane == joint.  The first version used one 256-thread CTA per frame: two block-wide barriers, then
// seven warps exited while one ran the long dependent fp64 tail -- 16 % of HBM bandwidth.  Here a warp streams its frame
// once (every load is a 128-byte coalesced row of 32 joints; four cells = 36 independent loads in flight per lane) with an
// ONLINE softmax (running maximum, accumulators rescaled when it moves), so nothing is kept in registers across the frame
// and no shared memory or barrier exists; the fp64 tail of one warp hides under the loads of the other resident warps.
constexpr int WARPS_PER_CTA = 4;

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_decode(const float *__restrict__ logits, const float *__restrict__ expand,
                                                              int n_out, DecodeParams prm, float *__restrict__ poses,
                                                              uint8_t *__restrict__ valid, int64_t n_frames,
                                                              const float *__restrict__ Ks, const float *__restrict__ Rs) {
  const int lane = threadIdx.x & 31;
  const int64_t f = (int64_t)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  if (Ks) {
    // per-frame camera (test-time augmentation, hpe.py:88-93: every augmented crop has its own intrinsics and its own
    // rotation/flip to undo): inverse intrinsics by the adjugate in double, rounded to float32 like np.linalg.inv(float32)
    double a[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = (double)__ldg(Ks + f * 9 + i);
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    const double inv[9] = {(a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det, (a[1] * a[5] - a[2] * a[4]) / det,
                           (a[5] * a[6] - a[3] * a[8]) / det, (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                           (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det, (a[0] * a[4] - a[1] * a[3]) / det};
#pragma unroll
    for (int i = 0; i < 9; ++i) { prm.invK[i] = (float)inv[i]; prm.R[i] = (double)__ldg(Rs + f * 9 + i); }
  }
  const float *src = logits + f * (int64_t)(64 * CH) + lane;
  const float inv7 = 1.0f / 7.0f;
  // running maxima and sums: 2-D head (S, X, Y), 3-D head (S, X, Y, Z); coordinates are linspace(0,1,8)
  float m2 = -1e30f, S2 = 0.f, X2 = 0.f, Y2 = 0.f;
  float m3 = -1e30f, S3 = 0.f, X3 = 0.f, Y3 = 0.f, Z3 = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 4) {
    float v[4][9];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 9; ++k) v[u][k] = __ldcs(src + (c0 + u) * CH + k * NJ);      // streamed once: evict-first
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int cell = c0 + u;
      const float yh = (float)(cell >> 3) * inv7, xw = (float)(cell & 7) * inv7;       // cell = h*8 + w; x <- w, y <- h
      {
        const float mn = fmaxf(m2, v[u][0]);
        const float sc = __expf(m2 - mn), e = __expf(v[u][0] - mn);
        S2 = fmaf(S2, sc, e); X2 = fmaf(X2, sc, e * xw); Y2 = fmaf(Y2, sc, e * yh);
        m2 = mn;
      }
      float mc = v[u][1];
#pragma unroll
      for (int d = 1; d < ND; ++d) mc = fmaxf(mc, v[u][1 + d]);
      const float mn = fmaxf(m3, mc);
      const float sc = __expf(m3 - mn);
      float es = 0.f, ez = 0.f;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const float e = __expf(v[u][1 + d] - mn);
        es += e;
        ez = fmaf(e, (float)d * inv7, ez);
      }
      S3 = fmaf(S3, sc, es); X3 = fmaf(X3, sc, es * xw); Y3 = fmaf(Y3, sc, es * yh); Z3 = fmaf(Z3, sc, ez);
      m3 = mn;
    }
  }
  // lane == joint j
  const double p2x = (double)(X2 / S2) * 255.0, p2y = (double)(Y2 / S2) * 255.0;
  const double r3x = (double)(X3 / S3), r3y = (double)(Y3 / S3), r3z = (double)(Z3 / S3);
  const bool fov = p2x >= 18.0 && p2x <= 238.0 && p2y >= 18.0 && p2y <= 238.0;   // misc.py:218-220
  const unsigned fmask = __ballot_sync(0xffffffffu, fov);
  const bool ok = __popc(fmask) * 4 >= NJ;                       // hpe.py:152
  if (lane == 0) valid[f] = ok ? 1 : 0;
  if (!ok) {
    for (int e = lane; e < n_out * 3; e += 32) poses[f * (int64_t)(n_out * 3) + e] = 0.f;
    return;
  }
  // normalised image coordinates: [x y 1] @ invK^T  (misc.py:185-186)
  const double nx = p2x * (double)prm.invK[0] + p2y * (double)prm.invK[1] + (double)prm.invK[2];
  const double ny = p2x * (double)prm.invK[3] + p2y * (double)prm.invK[4] + (double)prm.invK[5];
  // reconstruct_ref_fullpersp (misc.py:141-176): weighted LSQ  [I2 | -n/s2] ref' = b/sb
  const double s2 = sqrt(warp_sum(nx * nx + ny * ny) / (2.0 * NJ));
  const double bx = nx * r3z - r3x, by = ny * r3z - r3y;
  const double sb = sqrt(warp_sum(bx * bx + by * by) / (2.0 * NJ));
  const double wgt = (double)((fov ? 1.0f : 0.0f) + 1e-4f);
  const double w2 = wgt * wgt;
  const double ax = -nx / s2, ay = -ny / s2, rx = bx / sb, ry = by / sb;
  // normal equations M x = g, M = [[Sw,0,Sax],[0,Sw,Say],[Sax,Say,Saa]]
  const double Sw = warp_sum(w2), Sax = warp_sum(w2 * ax), Say = warp_sum(w2 * ay), Saa = warp_sum(w2 * (ax * ax + ay * ay));
  const double g0 = warp_sum(w2 * rx), g1 = warp_sum(w2 * ry), g2 = warp_sum(w2 * (ax * rx + ay * ry));
  // eliminate x0, x1:  x2 (Saa - (Sax^2+Say^2)/Sw) = g2 - (Sax g0 + Say g1)/Sw
  const double x2 = (g2 - (Sax * g0 + Say * g1) / Sw) / (Saa - (Sax * Sax + Say * Say) / Sw);
  const double x0 = (g0 - Sax * x2) / Sw, x1 = (g1 - Say * x2) / Sw;
  const double refx = x0 * sb, refy = x1 * sb, refz = x2 / s2 * sb;
  // misc.py:196-204
  double ax3, ay3, az3;
  if (fov) {
    const double zz = r3z + refz;
    ax3 = nx * zz; ay3 = ny * zz; az3 = zz;
  } else {
    ax3 = r3x + refx; ay3 = r3y + refy; az3 = r3z + refz;
  }
  // @ homo_inv (hpe.py:159): row vector times R
  const double qx = ax3 * prm.R[0] + ay3 * prm.R[3] + az3 * prm.R[6];
  const double qy = ax3 * prm.R[1] + ay3 * prm.R[4] + az3 * prm.R[7];
  const double qz = ax3 * prm.R[2] + ay3 * prm.R[5] + az3 * prm.R[8];
  // joint remap (hpe.py:162-164): out[k] = sum_j q[j] * E[j][k]; root-centre on output joint 0 (main.py:103): the root is
  // lane 0's own sum, broadcast -- so the root row comes out exactly 0.
  double r0x = 0, r0y = 0, r0z = 0;
  for (int k0 = 0; k0 < n_out; k0 += 32) {
    const int k = k0 + lane;
    double ox = 0, oy = 0, oz = 0;
#pragma unroll 8
    for (int j = 0; j < NJ; ++j) {
      const double ej = (k < n_out) ? (double)__ldg(expand + j * n_out + k) : 0.0;
      ox += __shfl_sync(0xffffffffu, qx, j) * ej;
      oy += __shfl_sync(0xffffffffu, qy, j) * ej;
      oz += __shfl_sync(0xffffffffu, qz, j) * ej;
    }
    if (k0 == 0) {
      r0x = __shfl_sync(0xffffffffu, ox, 0); r0y = __shfl_sync(0xffffffffu, oy, 0); r0z = __shfl_sync(0xffffffffu, oz, 0);
    }
    if (k < n_out) {
      float *o = poses + f * (int64_t)(n_out * 3) + k * 3;
      o[0] = (float)(ox - r0x); o[1] = (float)(oy - r0y); o[2] = (float)(oz - r0z);
    }
  }
}

}  // namespace

int arx_decode_launch(arx_handle *h, const float *logits, int64_t n_frames, const float *expand, int n_out, const float *K9,
                      const float *R9, float *poses, uint8_t *valid, cudaStream_t st, const float *Ks_dev, const float *Rs_dev) {
  if (n_frames == 0) return ARX_OK;
  DecodeParams p{};
  if (Ks_dev) {
    for (int64_t f0 = 0; f0 < n_frames; f0 += 1 << 30) {
      int64_t n = n_frames - f0 < (1 << 30) ? n_frames - f0 : (1 << 30);
      k_decode<<<(unsigned)((n + WARPS_PER_CTA - 1) / WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(logits + f0 * 64 * CH, expand, n_out, p,
                                                                                                 poses + f0 * n_out * 3, valid + f0, n, Ks_dev + f0 * 9,
                                                                                                 Rs_dev + f0 * 9);
      ARX_LAUNCH_CHECK(h);
    }
    return ARX_OK;
  }
  // inverse of the (float32) intrinsics via the adjugate in double, rounded to float32 like np.linalg.inv(float32)
  double a[9];
  for (int i = 0; i < 9; ++i) a[i] = (double)K9[i];
  const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  if (det == 0.0) return arx_fail(h, ARX_ERR_INVALID, "decode: singular intrinsics");
  const double inv[9] = {(a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det, (a[1] * a[5] - a[2] * a[4]) / det,
                         (a[5] * a[6] - a[3] * a[8]) / det, (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                         (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det, (a[0] * a[4] - a[1] * a[3]) / det};
  for (int i = 0; i < 9; ++i) { p.invK[i] = (float)inv[i]; p.R[i] = (double)R9[i]; }
  for (int64_t f0 = 0; f0 < n_frames; f0 += 1 << 30) {
    int64_t n = n_frames - f0 < (1 << 30) ? n_frames - f0 : (1 << 30);
    k_decode<<<(unsigned)((n + WARPS_PER_CTA - 1) / WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(logits + f0 * 64 * CH, expand, n_out, p,
                                                                                               poses + f0 * n_out * 3, valid + f0, n, nullptr, nullptr);
    ARX_LAUNCH_CHECK(h);
  }
  return ARX_OK;
}
