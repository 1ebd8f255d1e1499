// K3: intra-patch kNN graphs -- `knn()` of fn/snn_coder.py:31-39 and fd/snn_coder.py:25-32.
//
// One CTA per patch.  The M x M score matrix is built in the reference's expanded form
//   score[i][j] = ((-xx[j]) - inner[i][j]) - xx[i],  inner = -2 * <f_i, f_j>,  xx = sum_c f_c^2
// (fp32, squares and the xx sum individually rounded as torch does; the dot product is an fmaf chain
// over the channels in ascending order), then each row's top-k is extracted by a warp with the
// deterministic tie-break "larger score first, then lower index".  Works for xyz (C=3) and for the
// feature-space graphs of fd blocks 1..3 (C=64/128/256, rows `ld` floats apart).
// Roofline: shared-memory / FP32 bound; HBM traffic M*C*4 B in, M*k*4 B out per patch.
#include "common.cuh"
#include "kernels.h"

namespace sapcu {

constexpr int IK_THREADS = 256;
constexpr int IK_MMAX = 128;
constexpr int IK_CC = 64;     // channel chunk staged in shared memory

__global__ void __launch_bounds__(IK_THREADS)
intra_knn_kernel(const float* __restrict__ feat, int64_t ld, int M, int C, int k, int32_t* __restrict__ out) {
  extern __shared__ float sm[];
  const int nb = (M + 3) >> 2;          // 4x4 blocks per side
  const int Mp = nb * 4;
  const int cc_max = C < IK_CC ? C : IK_CC;
  const int fs = cc_max + 1;            // padded row stride of the staged chunk
  float* fsm = sm;                      // [Mp][fs]
  float* xx = fsm + Mp * fs;            // [Mp]
  float* sc = xx + Mp;                  // [M][M+1]
  const int tid = threadIdx.x;
  const float* base = feat + (int64_t)blockIdx.x * M * ld;

  float acc[3][4][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][i][j] = 0.0f;
  float myxx = 0.0f;   // thread i < M accumulates xx[i]

  for (int c0 = 0; c0 < C; c0 += cc_max) {
    const int cc = min(cc_max, C - c0);
    __syncthreads();
    for (int i = tid / 32; i < Mp; i += IK_THREADS / 32) {        // one warp per row: coalesced, no per-element division
      for (int c = tid & 31; c < cc; c += 32) fsm[i * fs + c] = (i < M) ? base[(int64_t)i * ld + c0 + c] : 0.0f;
    }
    __syncthreads();
    if (tid < M) {
      for (int c = 0; c < cc; ++c) { const float v = fsm[tid * fs + c]; myxx = __fadd_rn(myxx, __fmul_rn(v, v)); }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int b = tid + r * IK_THREADS;
      if (b < nb * nb) {
        // block (bi, bj) owns rows {bi + nb*i} x {bj + nb*j}: consecutive lanes read consecutive rows (stride fs = odd
        // number of words) -> conflict-free shared-memory loads
        const int bi = b / nb, bj = b - bi * nb;
        const float* fa = fsm + bi * fs;
        const float* fb = fsm + bj * fs;
        const int rs = nb * fs;
        for (int c = 0; c < cc; ++c) {
          float a[4], bb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { a[i] = fa[i * rs + c]; bb[i] = fb[i * rs + c]; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][i][j] = fmaf(a[i], bb[j], acc[r][i][j]);
        }
      }
    }
  }
  if (tid < M) xx[tid] = myxx;
  __syncthreads();
  const int ss = M + 1;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int b = tid + r * IK_THREADS;
    if (b < nb * nb) {
      const int bi = b / nb, bj = b - bi * nb;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int gi = bi + nb * i, gj = bj + nb * j;
          if (gi < M && gj < M) {
            const float inner = __fmul_rn(-2.0f, acc[r][i][j]);
            sc[gi * ss + gj] = __fsub_rn(__fsub_rn(-xx[gj], inner), xx[gi]);
          }
        }
    }
  }
  __syncthreads();
  // top-k per row: one warp per row
  const int warp = tid >> 5, lane = tid & 31;
  int32_t* o = out + (int64_t)blockIdx.x * M * k;
  for (int i = warp; i < M; i += IK_THREADS / 32) {
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int j = lane + 32 * q; v[q] = (j < M) ? sc[i * ss + j] : -INFINITY; }
    for (int t = 0; t < k; ++t) {
      float bv = v[0]; int bj = lane;
#pragma unroll
      for (int q = 1; q < 4; ++q) if (v[q] > bv) { bv = v[q]; bj = lane + 32 * q; }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
        if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
      }
      if ((bj & 31) == lane) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if ((bj >> 5) == q) v[q] = -INFINITY;
      }
      if (lane == 0) o[i * k + t] = bj;
    }
  }
}

static size_t intra_knn_smem(int M, int C) {
  const int nb = (M + 3) >> 2, Mp = nb * 4;
  const int cc = C < IK_CC ? C : IK_CC;
  return sizeof(float) * ((size_t)Mp * (cc + 1) + Mp + (size_t)M * (M + 1));
}

int launch_intra_knn(const float* feat, int64_t ld, int64_t S, int M, int C, int k, int32_t* idx, cudaStream_t st) {
  SAPCU_REQUIRE(M >= 1 && M <= IK_MMAX, "intra_knn: M=%d outside [1,%d]", M, IK_MMAX);
  SAPCU_REQUIRE(k >= 1 && k <= M, "intra_knn: k=%d outside [1,M=%d]", k, M);
  SAPCU_REQUIRE(((M + 3) / 4) * ((M + 3) / 4) <= 3 * IK_THREADS, "intra_knn: M too large");
  if (S == 0) return 0;
  const size_t smem = intra_knn_smem(M, C);
  static bool attr_done = false;
  if (!attr_done) {
    SAPCU_CUDA_CHECK(cudaFuncSetAttribute(intra_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_done = true;
  }
  intra_knn_kernel<<<(unsigned)S, IK_THREADS, smem, st>>>(feat, ld, M, C, k, idx);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

}  // namespace sapcu
