// Tuple index table built on device, bit-exact with itertools.combinations(range(T), c)
// (reference: modules/ar/utils/model.py:51-55 -- lexicographic order).
// One thread per rank: combinatorial-number-system unranking, integer arithmetic only.
#include "arx_internal.cuh"

namespace {

__device__ __forceinline__ unsigned long long binom(int n, int k) {
  if (k < 0 || k > n) return 0ull;
  if (k > n - k) k = n - k;
  unsigned long long r = 1ull;
  for (int i = 1; i <= k; ++i) r = r * (unsigned long long)(n - k + i) / (unsigned long long)i;  // exact at every step
  return r;
}

__global__ void k_tuple_table(int32_t *__restrict__ out, int T, int c, int N) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  unsigned long long rem = (unsigned long long)r;
  int prev = -1;
  for (int p = 0; p < c; ++p) {
    int i = prev + 1;
    for (;; ++i) {
      // number of combinations whose p-th element is i (given the prefix): C(T-1-i, c-1-p)
      unsigned long long cnt = binom(T - 1 - i, c - 1 - p);
      if (rem < cnt) break;
      rem -= cnt;
    }
    out[r * c + p] = i;
    prev = i;
  }
}

}  // namespace

int arx_build_tuple_table(arx_handle *h, int T, int c, int N, int32_t *out_dev, cudaStream_t st) {
  k_tuple_table<<<(N + 127) / 128, 128, 0, st>>>(out_dev, T, c, N);
  ARX_LAUNCH_CHECK(h);
  return ARX_OK;
}
