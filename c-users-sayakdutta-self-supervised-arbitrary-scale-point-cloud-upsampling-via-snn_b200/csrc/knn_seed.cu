// K1: exact brute-force seed -> input-cloud kNN (replaces sklearn KDTree.query, generation.py:127,153).
//
// One warp owns one seed.  The cloud is streamed through shared memory in fp32 tiles; each lane tests
// one cloud point per iteration with a CONSERVATIVE fp32 filter and only survivors pay for the exact
// fp64 squared distance ((dx*dx + dy*dy) + dz*dz, no contraction -- the KDTree's reduced distance).
// The warp keeps the K best as a sorted list of (fp64 distance, index) in shared memory; survivors are
// queued and merged by rank (keys are unique, ties resolved by the lowest index), so the result is the
// exact fp64 ordering independent of how the cloud is tiled.
//
// Filter soundness: with p~,q~ the fp32 roundings of p,q and Rmax >= every |coordinate|,
// |sqrt(d32) - sqrt(d)| <= 2^-22 Rmax + 2^-21 sqrt(d); a point with d <= tau therefore always has
// d32 <= ((1+2^-20) sqrt(tau) + 2^-21 Rmax)^2 (1+2^-20) =: tau32, recomputed whenever tau changes.
//
// Roofline: FP32 issue (about 8 lane-ops per pair), cloud tiles come from L2; HBM traffic is only
// S*(24 + 4K) bytes.
#include "common.cuh"
#include "kernels.h"
#include <float.h>
#include <algorithm>

namespace sapcu {

constexpr int KNN_KMAX = 128;
constexpr int KNN_WARPS = 8;
constexpr int KNN_TILE = 1024;
constexpr int KNN_QCAP = 64;

__global__ void cloud_to_f32_kernel(const double* __restrict__ cloud, int64_t n3, float* __restrict__ out,
                                    float* __restrict__ rmax) {
  float m = 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = cloud[i];
    if (out) out[i] = (float)v;
    m = fmaxf(m, (float)fabs(v) * 1.0000002f);
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(rmax), __float_as_int(m));   // m >= 0
}

struct KnnKey { double d; int i; };
__device__ __forceinline__ bool key_less(double da, int ia, double db, int ib) {
  return da < db || (da == db && ia < ib);
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_seed_kernel(const double* __restrict__ cloud, const float* __restrict__ cloud32, int64_t N,
                const double* __restrict__ seeds, int64_t S, int K, const float* __restrict__ rmax_p,
                int32_t* __restrict__ out_idx) {
  __shared__ float tx[KNN_TILE], ty[KNN_TILE], tz[KNN_TILE];
  __shared__ double ld[KNN_WARPS][2][KNN_KMAX];
  __shared__ int li[KNN_WARPS][2][KNN_KMAX];
  __shared__ double qd[KNN_WARPS][KNN_QCAP];
  __shared__ int qi[KNN_WARPS][KNN_QCAP];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * KNN_WARPS + warp;
  const bool active = s < S;
  const float rmax = *rmax_p;

  double sx = 0, sy = 0, sz = 0;
  if (active) { sx = seeds[3 * s]; sy = seeds[3 * s + 1]; sz = seeds[3 * s + 2]; }
  const float fx = (float)sx, fy = (float)sy, fz = (float)sz;

  int cur = 0;
  for (int j = lane; j < KNN_KMAX; j += 32) { ld[warp][0][j] = DBL_MAX; li[warp][0][j] = INT_MAX; }
  __syncwarp();
  double tau = DBL_MAX; int tau_i = INT_MAX;
  float tau32 = FLT_MAX;
  int qn = 0;

  auto flush = [&]() {
    // merge queue (qn unsorted, unique keys) into the sorted list by rank
    double* L = ld[warp][cur]; int* LI = li[warp][cur];
    double* Lo = ld[warp][cur ^ 1]; int* LIo = li[warp][cur ^ 1];
    for (int e = lane; e < qn; e += 32) {
      const double d = qd[warp][e]; const int id = qi[warp][e];
      int rank = 0;
      for (int j = 0; j < qn; ++j) rank += key_less(qd[warp][j], qi[warp][j], d, id) ? 1 : 0;
      int lo = 0, hi = K;   // first list position whose key is not less than (d,id)
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (key_less(L[mid], LI[mid], d, id)) lo = mid + 1; else hi = mid; }
      const int pos = rank + lo;
      if (pos < K) { Lo[pos] = d; LIo[pos] = id; }
    }
    for (int e = lane; e < K; e += 32) {
      const double d = L[e]; const int id = LI[e];
      int rank = 0;
      for (int j = 0; j < qn; ++j) rank += key_less(qd[warp][j], qi[warp][j], d, id) ? 1 : 0;
      const int pos = rank + e;
      if (pos < K) { Lo[pos] = d; LIo[pos] = id; }
    }
    __syncwarp();
    cur ^= 1; qn = 0;
    tau = ld[warp][cur][K - 1]; tau_i = li[warp][cur][K - 1];
    if (tau < DBL_MAX) {
      const float rt = sqrtf((float)tau) * 1.000002f;   // >= (1+2^-20) sqrt(tau) incl. rounding of the cast/sqrt
      const float b = rt + 4.76837158e-7f * rmax;       // 2^-21 Rmax
      tau32 = b * b * 1.000002f;
    }
    __syncwarp();
  };

  for (int64_t base = 0; base < N; base += KNN_TILE) {
    const int tn = (int)min((int64_t)KNN_TILE, N - base);
    __syncthreads();
    for (int j = threadIdx.x; j < tn; j += blockDim.x) {
      tx[j] = cloud32[3 * (base + j)]; ty[j] = cloud32[3 * (base + j) + 1]; tz[j] = cloud32[3 * (base + j) + 2];
    }
    __syncthreads();
    if (!active) continue;
    for (int j0 = 0; j0 < tn; j0 += 32) {
      const int j = j0 + lane;
      bool pass = false;
      if (j < tn) {
        const float dx = tx[j] - fx, dy = ty[j] - fy, dz = tz[j] - fz;
        const float d32 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        pass = d32 <= tau32;
      }
      if (__any_sync(0xffffffffu, pass)) {
        double d = 0; const int gi = (int)(base + j);
        if (pass) {
          const double ex = cloud[3 * (int64_t)gi] - sx, ey = cloud[3 * (int64_t)gi + 1] - sy, ez = cloud[3 * (int64_t)gi + 2] - sz;
          d = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
          pass = key_less(d, gi, tau, tau_i);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
          if (pass) {
            const int pos = qn + __popc(bal & ((1u << lane) - 1u));
            qd[warp][pos] = d; qi[warp][pos] = gi;
          }
          qn += __popc(bal);
          __syncwarp();
          if (qn > KNN_QCAP - 32) flush();
        }
      }
    }
  }
  if (active) {
    if (qn) flush();
    for (int j = lane; j < K; j += 32) out_idx[s * K + j] = li[warp][cur][j];
  }
}

int launch_knn_seed(const double* cloud, int64_t N, const double* seeds, int64_t S, int K, int32_t* idx,
                    float* cloud32_scratch, float* rmax_scratch, cudaStream_t st) {
  SAPCU_REQUIRE(K >= 1 && K <= KNN_KMAX, "sapcu_knn: K=%d outside [1,%d]", K, KNN_KMAX);
  SAPCU_REQUIRE(N >= K, "sapcu_knn: K=%d > N=%lld", K, (long long)N);
  SAPCU_REQUIRE(N < (int64_t)INT32_MAX / 3, "sapcu_knn: N too large for int32 indices");
  if (S == 0) return 0;
  SAPCU_CUDA_CHECK(cudaMemsetAsync(rmax_scratch, 0, sizeof(float), st));
  const int blocks = (int)std::min<int64_t>(ceil_div(3 * N, 256), 148 * 8);
  cloud_to_f32_kernel<<<blocks, 256, 0, st>>>(cloud, 3 * N, cloud32_scratch, rmax_scratch);
  SAPCU_LAUNCH_CHECK();
  const int sblocks = (int)std::min<int64_t>(ceil_div(3 * S, 256), 148 * 8);
  cloud_to_f32_kernel<<<sblocks, 256, 0, st>>>(seeds, 3 * S, nullptr, rmax_scratch);
  SAPCU_LAUNCH_CHECK();
  knn_seed_kernel<<<(unsigned)ceil_div(S, KNN_WARPS), KNN_WARPS * 32, 0, st>>>(cloud, cloud32_scratch, N, seeds, S, K,
                                                                               rmax_scratch, idx);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

}  // namespace sapcu
