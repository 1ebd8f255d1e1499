// K2 / K9b: the memory-bound glue of generation.py around the two networks.
//   gather_center_rotate : data[idx] - seed (fp64), optional per-seed Rodrigues rotation n̂ -> x̂ (fp64),
//                          cast to fp32                       (generation.py:128-129,137,154-160,168)
//   renormalize          : F.normalize(n, dim=-1), eps 1e-12  (generation.py:139)
//   displace             : seed + (double)(n * d)             (generation.py:171-172)
// Roofline: HBM (reads 4K B idx + L2-resident cloud gathers, writes 12K B per seed).
#include "common.cuh"
#include "kernels.h"

namespace sapcu {

// rotation_matrix_from_vectors(n, [1,0,0]) of generation.py:30-47, in the same mixed precision:
// a = n / ||n|| in fp32, everything after that in fp64; identity when cross(a, x̂) == 0.
__device__ __forceinline__ void rodrigues_to_x(const float* __restrict__ n, double (&R)[3][3]) {
  const float n0 = n[0], n1 = n[1], n2 = n[2];
  // np.linalg.norm(float32[3]) = sqrt(x.dot(x)): products rounded to fp32, accumulated in double, sum rounded to fp32
  const float nn = sqrtf((float)(((double)__fmul_rn(n0, n0) + (double)__fmul_rn(n1, n1)) + (double)__fmul_rn(n2, n2)));
  const double a0 = (double)__fdiv_rn(n0, nn), a1 = (double)__fdiv_rn(n1, nn), a2 = (double)__fdiv_rn(n2, nn);
  // v = a x (1,0,0) = (0, a2, -a1)
  const double v0 = 0.0, v1 = a2, v2 = -a1;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
  if (v1 == 0.0 && v2 == 0.0) return;   // `any(v)` false: identical OR opposite direction -> identity
  const double c = a0;
  const double s = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1)), __dmul_rn(v2, v2)));
  const double Km[3][3] = {{0.0, -v2, v1}, {v2, 0.0, -v0}, {-v1, v0, 0.0}};
  const double f = (1.0 - c) / __dmul_rn(s, s);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double k2 = __dmul_rn(Km[i][0], Km[0][j]);
      k2 = __dadd_rn(k2, __dmul_rn(Km[i][1], Km[1][j]));
      k2 = __dadd_rn(k2, __dmul_rn(Km[i][2], Km[2][j]));
      R[i][j] = __dadd_rn(__dadd_rn(R[i][j], Km[i][j]), __dmul_rn(k2, f));
    }
}

__global__ void gather_center_rotate_kernel(const double* __restrict__ cloud, const double* __restrict__ seeds,
                                            const int32_t* __restrict__ idx, int64_t S, int K,
                                            const float* __restrict__ normals, float* __restrict__ patches) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= S * K) return;
  const int64_t s = e / K;
  const int64_t p = idx[e];
  const double dx = cloud[3 * p] - seeds[3 * s];
  const double dy = cloud[3 * p + 1] - seeds[3 * s + 1];
  const double dz = cloud[3 * p + 2] - seeds[3 * s + 2];
  double ox = dx, oy = dy, oz = dz;
  if (normals) {
    double R[3][3];
    rodrigues_to_x(normals + 3 * s, R);
    ox = __dadd_rn(__dadd_rn(__dmul_rn(R[0][0], dx), __dmul_rn(R[0][1], dy)), __dmul_rn(R[0][2], dz));
    oy = __dadd_rn(__dadd_rn(__dmul_rn(R[1][0], dx), __dmul_rn(R[1][1], dy)), __dmul_rn(R[1][2], dz));
    oz = __dadd_rn(__dadd_rn(__dmul_rn(R[2][0], dx), __dmul_rn(R[2][1], dy)), __dmul_rn(R[2][2], dz));
  }
  patches[3 * e] = (float)ox; patches[3 * e + 1] = (float)oy; patches[3 * e + 2] = (float)oz;
}

__global__ void renormalize_kernel(float* __restrict__ n, int64_t S) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float x = n[3 * s], y = n[3 * s + 1], z = n[3 * s + 2];
  const float nn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  const float den = fmaxf(nn, 1e-12f);
  n[3 * s] = __fdiv_rn(x, den); n[3 * s + 1] = __fdiv_rn(y, den); n[3 * s + 2] = __fdiv_rn(z, den);
}

__global__ void displace_kernel(const double* __restrict__ seeds, const float* __restrict__ n,
                                const float* __restrict__ d, int64_t S, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * S) return;
  out[i] = seeds[i] + (double)__fmul_rn(n[i], d[i / 3]);
}

int launch_gather_center_rotate(const double* cloud, const double* seeds, const int32_t* idx, int64_t S, int K,
                                const float* normals, float* patches, cudaStream_t st) {
  if (S == 0) return 0;
  gather_center_rotate_kernel<<<(unsigned)ceil_div(S * K, 256), 256, 0, st>>>(cloud, seeds, idx, S, K, normals, patches);
  SAPCU_LAUNCH_CHECK();
  return 0;
}
int launch_renormalize(float* n, int64_t S, cudaStream_t st) {
  if (S == 0) return 0;
  renormalize_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(n, S);
  SAPCU_LAUNCH_CHECK();
  return 0;
}
int launch_displace(const double* seeds, const float* n, const float* d, int64_t S, double* out, cudaStream_t st) {
  if (S == 0) return 0;
  displace_kernel<<<(unsigned)ceil_div(3 * S, 256), 256, 0, st>>>(seeds, n, d, S, out);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

}  // namespace sapcu
