// K3: intra-patch kNN graphs -- `knn()` of fn/snn_coder.py:31-39 and fd/snn_coder.py:25-32.
//
// One CTA per patch.  The M x M score matrix is built in the reference's expanded form
//   score[i][j] = ((-xx[j]) - inner[i][j]) - xx[i],  inner = -2 * <f_i, f_j>,  xx = sum_c f_c^2
// (fp32, squares and the xx sum individually rounded as torch does; the dot product is an fmaf chain
// over the channels in ascending order), then each row's top-k is extracted by a warp with the
// deterministic tie-break "larger score first, then lower index".  Works for xyz (C=3) and for the
// feature-space graphs of fd blocks 1..3 (C=64/128/256, rows `ld` floats apart).
//   * <f_i, f_j> is bitwise symmetric (same products, same order), so only the 4x4 register blocks on or above the
//     diagonal are accumulated (325 of 625 for M = 100) and each feeds both score[i][j] and score[j][i];
//   * top-k: scores become order-preserving 32-bit keys, each lane keeps its 4 candidates sorted; per extraction one
//     REDUX.MAX over the lanes' heads finds the best key and one REDUX.MIN the lowest index holding it, the owner shifts
//     its list -- ~12 instructions, no divergence.
// Roofline: shared-memory / issue bound; HBM traffic M*C*4 B in, M*k*4 B out per patch.
#include "common.cuh"
#include "kernels.h"

namespace sapcu {

constexpr int IK_THREADS = 352;
constexpr int IK_ROUNDS = 2;  // upper-triangular 4x4 blocks per thread: nb(nb+1)/2 <= 704 covers M <= 128
constexpr int IK_MMAX = 128;
constexpr int IK_CC = 64;     // channel chunk staged in shared memory

__device__ __forceinline__ uint32_t ik_key(float x) {     // monotone float -> uint32 (larger score = larger key)
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(IK_THREADS)
intra_knn_kernel(const float* __restrict__ feat, int64_t ld, int M, int C, int k, int32_t* __restrict__ out) {
  extern __shared__ float sm[];
  const int nb = (M + 3) >> 2;          // 4x4 blocks per side
  const int Mp = nb * 4;
  const int nblk = nb * (nb + 1) / 2;   // blocks with bi <= bj
  const int cc_max = C < IK_CC ? C : IK_CC;
  const int fs = cc_max + 1;            // padded row stride of the staged chunk
  float* fsm = sm;                      // [Mp][fs]
  float* xx = fsm + Mp * fs;            // [Mp]
  float* sc = xx + Mp;                  // [M][M+1]
  const int tid = threadIdx.x;
  const float* base = feat + (int64_t)blockIdx.x * M * ld;

  // block id -> (bi, bj), row-major over the upper triangle: row bi starts at bi*nb - bi(bi-1)/2
  int tbi[IK_ROUNDS], tbj[IK_ROUNDS];
#pragma unroll
  for (int r = 0; r < IK_ROUNDS; ++r) {
    const int b = tid + r * IK_THREADS;
    int bi = 0, start = 0;
    while (bi + 1 < nb && start + (nb - bi) <= b) { start += nb - bi; ++bi; }
    tbi[r] = bi; tbj[r] = bi + (b - start);
  }

  float acc[IK_ROUNDS][4][4];
#pragma unroll
  for (int r = 0; r < IK_ROUNDS; ++r)
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
    for (int r = 0; r < IK_ROUNDS; ++r) {
      if (tid + r * IK_THREADS < nblk) {
        // block (bi, bj) owns rows {bi + nb*i} x {bj + nb*j}: lanes of a warp share bi (broadcast) and walk consecutive bj
        // (row stride fs = odd number of words) -> conflict-free shared-memory loads
        const float* fa = fsm + tbi[r] * fs;
        const float* fb = fsm + tbj[r] * fs;
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
  for (int r = 0; r < IK_ROUNDS; ++r) {
    if (tid + r * IK_THREADS < nblk) {
      const int bi = tbi[r], bj = tbj[r];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int gi = bi + nb * i, gj = bj + nb * j;
          if (gi < M && gj < M) {
            const float inner = __fmul_rn(-2.0f, acc[r][i][j]);
            sc[gi * ss + gj] = __fsub_rn(__fsub_rn(-xx[gj], inner), xx[gi]);
            if (bi != bj) sc[gj * ss + gi] = __fsub_rn(__fsub_rn(-xx[gi], inner), xx[gj]);
          }
        }
    }
  }
  __syncthreads();
  // top-k per row: one warp per row, 4 candidates per lane (j = lane + 32 q) kept SORTED (key descending, index ascending
  // among equal keys) so that the lane's best candidate is always key[0]: retiring a winner is a predicated shift instead of
  // a divergent rescan, and the winners go to the row's own (consumed) score storage, then out in coalesced stores
  const int warp = tid >> 5, lane = tid & 31;
  int32_t* o = out + (int64_t)blockIdx.x * M * k;
  for (int i = warp; i < M; i += IK_THREADS / 32) {
    uint32_t key[4], id[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = lane + 32 * q;
      key[q] = (j < M) ? ik_key(__fadd_rn(sc[i * ss + j], 0.0f)) : 0u;   // (-0 -> +0: equal scores, equal keys)
      id[q] = (uint32_t)j;
    }
    auto cex = [&](int x, int y) {                           // order (x, y): larger key first, then lower index
      const bool sw = key[x] < key[y] || (key[x] == key[y] && id[x] > id[y]);
      const uint32_t kx = sw ? key[y] : key[x], ky = sw ? key[x] : key[y], ix = sw ? id[y] : id[x], iy = sw ? id[x] : id[y];
      key[x] = kx; key[y] = ky; id[x] = ix; id[y] = iy;
    };
    cex(0, 1); cex(2, 3); cex(0, 2); cex(1, 3); cex(1, 2);
    uint32_t ids = id[0] | (id[1] << 8) | (id[2] << 16) | (id[3] << 24);   // indices < 128: one byte each, head in the low byte
    __syncwarp();                                            // every lane has read its scores: the row is free for the winners
    int32_t* res = reinterpret_cast<int32_t*>(sc + i * ss);
    for (int t = 0; t < k; ++t) {
      const uint32_t best = __reduce_max_sync(0xffffffffu, key[0]);
      const uint32_t head = ids & 0xffu;
      const uint32_t w = __reduce_min_sync(0xffffffffu, key[0] == best ? head : 0xffffffffu);
      if (key[0] == best && w == head) {                     // the owner records and retires the winner
        res[t] = (int32_t)w;
        key[0] = key[1]; key[1] = key[2]; key[2] = key[3]; key[3] = 0u;
        ids >>= 8;
      }
    }
    __syncwarp();
    for (int t = lane; t < k; t += 32) o[i * k + t] = res[t];
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
  SAPCU_REQUIRE(((M + 3) / 4) * ((M + 3) / 4 + 1) / 2 <= IK_ROUNDS * IK_THREADS, "intra_knn: M too large");
  if (S == 0) return 0;
  const size_t smem = intra_knn_smem(M, C);
  static PerDeviceOnce once;
  {
    const int rc = once.run([]() -> int {
      SAPCU_CUDA_CHECK(cudaFuncSetAttribute(intra_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      return 0;
    });
    if (rc) return rc;
  }
  intra_knn_kernel<<<(unsigned)S, IK_THREADS, smem, st>>>(feat, ld, M, C, k, idx);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

}  // namespace sapcu
