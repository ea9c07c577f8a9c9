// Micro-benchmarks of the sm_100a primitives the attention kernel leans on (dev tool).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../isbfsar_b200/csrc -o ubench ubench.cu
#include <cstdio>
#include <cstdlib>
#include "arx_ptx.cuh"
using namespace ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// A: TMEM read throughput.  nwarps warps, each loops `iters` times over 4 x (ld 32 cols) + wait.
__global__ void k_ldtm(int iters, int mode, long long *out, float *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[4][32];
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tm + ((i & 1) * 128 + c * 32), r[c]);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) { tmem_ld32(tm + ((i & 1) * 128 + c * 32), r[c]); tmem_ld_wait(); }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc += __uint_as_float(r[c][0] ^ r[c][31]);
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

// B: MUFU.EX2 throughput: each thread does `iters` x 32 independent ex2
__global__ void k_ex2(int iters, long long *out, float *sink) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = 0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
  }
  long long t1 = clock64();
  __syncthreads();
  float a = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) a += x[j];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (a == 123.456f) sink[0] = a;
}

// B2: FFMA throughput (to calibrate polynomial exp emulation)
__global__ void k_ffma(int iters, long long *out, float *sink) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = 0.001f * (threadIdx.x + j);
  float c = 1.0001f + threadIdx.x * 1e-9f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fmaf(x[j], c, 0.5f);
  }
  long long t1 = clock64();
  __syncthreads();
  float a = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) a += x[j];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (a == 123.456f) sink[0] = a;
}

// C: UTCHMMA 128x128x16 (SS, K-major SW128) issue rate, optionally with 4 extra warps hammering TMEM reads / smem stores
__global__ void k_mma(int iters, int side, int nmma_n, long long *out, float *sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = clock64(), t1 = t0;
  if (warp == 0) {
    if (elect_one()) {
      const uint64_t D = smem_desc_sw128(16, 1024);
      const uint32_t idesc = idesc_f16(128, nmma_n, 0, 0);
      const uint32_t sb = smem_u32(smem);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          mma_f16_ss(tm + (i & 1) * 256, smem_desc_at(D, sb + off), smem_desc_at(D, sb + 32768 + off), idesc, kk > 0);
        }
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0);
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  } else if (side == 1 && warp >= 4 && warp < 8) {   // concurrent TMEM reads
    const uint32_t tml = tm + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0;
    for (int i = 0; i < iters * 2; ++i) {
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tml + 128 + c * 32, r[c]);
      tmem_ld_wait();
      acc += __uint_as_float(r[0][0] ^ r[3][31]);
    }
    if (acc == 123.456f) sink[0] = acc;
  } else if (side == 2 && warp >= 4 && warp < 8) {   // concurrent 16-byte smem stores (P-like), 32 KB per iter
    uint8_t *pb = smem + 65536 + (threadIdx.x - 128) * 16;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 16; ++c) *reinterpret_cast<uint4 *>(pb + c * 2048) = make_uint4(i, c, i, c);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

// B3: F2FP (fp32x2 -> fp16x2) and FMUL2 throughput
__global__ void k_cvt(int iters, int mode, long long *out, float *sink) {
  float x[32]; uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = 0.001f * (threadIdx.x + j);
  uint64_t zz = pack2(1.0001f, 0.9999f);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) { uint32_t h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[j + 1]), "f"(x[j])); acc ^= h; }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2) { uint64_t v = mul2(pack2(x[j], x[j + 1]), zz); unpack2(v, x[j], x[j + 1]); }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float a = acc;
#pragma unroll
  for (int j = 0; j < 32; ++j) a += x[j];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (a == 123.456f) sink[0] = a;
}

// B4: clean FFMA2 / FFMA / PRMT / IADD throughput, 16 independent chains per thread
__global__ void k_pk(int iters, int mode, long long *out, float *sink) {
  uint64_t v[16]; float x[32]; uint32_t u[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = pack2(0.001f * (threadIdx.x + j), 0.002f * (threadIdx.x + j));
#pragma unroll
  for (int j = 0; j < 32; ++j) { x[j] = 0.001f * (threadIdx.x + j); u[j] = threadIdx.x * 77 + j; }
  const uint64_t c2 = pack2(1.0001f, 0.9999f), d2 = pack2(0.5f, 0.25f);
  const float c = 1.0001f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fma2(v[j], c2, d2);
    } else if (mode == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaf(x[j], c, 0.5f);
    } else if (mode == 2) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) u[j] = __byte_perm(u[j], u[j + 1], 0x7632) + i;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = add2(v[j], c2);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float a = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) { float lo, hi; unpack2(v[j], lo, hi); a += lo + hi; }
#pragma unroll
  for (int j = 0; j < 32; ++j) a += x[j] + u[j];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (a == 123.456f) sink[0] = a;
}

// D: MUFU batch warps (0-3) next to "epilogue-like" warps (4-7) on the same SMSPs.
// side: 0 none, 1 FFMA stream, 2 LDTM+FFMA, 3 FADD/FSEL/ISETP stream, 4 shuffle stream
__global__ void k_mix(int iters, int side, long long *out, float *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float a = 0;
  if (warp < 4) {
    uint32_t x[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) x[j] = __float_as_uint(0.001f * (threadIdx.x + j));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 128; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[j]));
    }
    long long t1 = clock64();
#pragma unroll
    for (int j = 0; j < 128; ++j) a += __uint_as_float(x[j]);
    if (threadIdx.x == 0) out[0] = t1 - t0;
  } else if (side == 1) {
    float y[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = 0.001f * (threadIdx.x + j);
    float c = 1.0001f + threadIdx.x * 1e-9f;
    for (int i = 0; i < iters * 4; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = fmaf(y[j], c, 0.5f);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) a += y[j];
  } else if (side == 2) {
    float acc = 0;
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(tm + c * 32, r); tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { float d = (acc + 1.0f) - __uint_as_float(r[j]); acc = fmaf(d, d, acc * 0.5f); }
      }
    }
    a = acc;
  } else if (side == 3) {
    float y[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters * 4; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = (y[j] > 0.5f ? y[j] : 0.25f) + 0.125f;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) a += y[j];
  } else if (side == 4) {
    float y = threadIdx.x;
    for (int i = 0; i < iters * 64; ++i) y += __shfl_xor_sync(0xffffffffu, y, 1 + (i & 15));
    a = y;
  }
  if (a == 123.456f) sink[0] = a;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long *out; float *sink;
  CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMalloc(&sink, 16));
  long long h[148];
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      k_ldtm<<<1, nw * 32>>>(iters, mode, out, sink); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      double cyc = (double)h[0] / iters;
      printf("LDTM mode=%d (%s) warps=%2d: %.1f clk per 4x(32x32b.x32)/warp -> %.1f B/clk/SM\n", mode, mode ? "wait each" : "4 then wait", nw, cyc,
             nw * 4.0 * 32 * 32 * 4 / cyc);
    }
  for (int nw : {1, 4, 8, 16, 32}) {
    k_ex2<<<1, nw * 32>>>(iters, out, sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("EX2 warps=%2d: %.2f clk per warp-instr/warp -> %.2f lanes/clk/SM\n", nw, (double)h[0] / iters / 32, nw * 32.0 * 32 * iters / h[0]);
    k_ffma<<<1, nw * 32>>>(iters, out, sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("FFMA warps=%2d: %.2f clk per warp-instr/warp -> %.2f lanes/clk/SM\n", nw, (double)h[0] / iters / 32, nw * 32.0 * 32 * iters / h[0]);
  }
  for (int mode = 0; mode < 2; ++mode)
    for (int nw : {4, 8}) {
      k_cvt<<<1, nw * 32>>>(iters, mode, out, sink); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      printf("%s warps=%d: %.2f clk per warp-instr/warp\n", mode ? "FMUL2" : "F2FP.PACK", nw, (double)h[0] / iters / 16);
    }
  for (int mode = 0; mode < 4; ++mode)
    for (int nw : {4, 8}) {
      k_pk<<<1, nw * 32>>>(iters, mode, out, sink); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      const char *nm[] = {"FFMA2 (16/iter)", "FFMA (32/iter)", "PRMT+IADD (16/iter)", "FADD2 (16/iter)"};
      printf("%s warps=%d: %.2f clk per iteration per warp\n", nm[mode], nw, (double)h[0] / iters);
    }
  for (int side = 0; side < 5; ++side) {
    k_mix<<<1, 256>>>(200, side, out, sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("MIX side=%d: MUFU batch of 128 takes %.0f clk (%.2f clk/instr)\n", side, (double)h[0] / 200, (double)h[0] / 200 / 128);
  }
  CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int n : {128, 256})
    for (int side = 0; side < 3; ++side) {
      k_mma<<<1, 256, 200 * 1024>>>(1000, side, n, out, sink); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      printf("UTCHMMA 128x%dx16 SS side=%d (%s): %.1f clk per MMA (1 CTA)\n", n, side, side == 0 ? "alone" : side == 1 ? "+TMEM reads" : "+smem stores",
             (double)h[0] / 1000 / 8);
    }
  // full chip: 148 CTAs of MMA alone (power/clock effects)
  k_mma<<<148, 256, 200 * 1024>>>(4000, 0, 128, out, sink); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost));
  double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("UTCHMMA 128x128x16 on 148 CTAs: %.1f clk per MMA (max CTA)\n", mx / 4000 / 8);
  return 0;
}
