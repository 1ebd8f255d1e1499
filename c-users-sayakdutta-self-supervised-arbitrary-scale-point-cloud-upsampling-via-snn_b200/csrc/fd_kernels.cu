// Non-GEMM kernels of the fd (distance estimation) forward, fd/snn_coder.py:392-492,711-798.
//   fd_block0        : the four multi-scale EdgeConvs on xyz (6 -> 64, BN, LeakyReLU, max over k_s), :413-417
//   neuron_unroll    : T steps of an EIF/LIF layer from the zero state, every step's spikes stored (:432-474);
//                      also the stand-alone LIF^T operator of the C ABI
//   temporal_lif     : softmax-weighted sum over time of the pooled features + the final LIF step (:482-490)
//   head_attention   : StandardSelfAttention core on [S,64] (:782-795)
//   layernorm_rows   : LayerNorm over the last dim (:798)
//   fd_tail          : Linear 32 -> 1 + Softplus(beta=5) (:722-725)
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"
#include "neuron.cuh"

namespace sapcu {

constexpr int B0_MAXK = 64;      // largest k_scale supported
constexpr int B0_MAXSCALES = 8;

struct Block0Args {
  const float* xyz; const int32_t* idx; int ldi; int Mpts; int64_t P;
  int nscales; int ks[B0_MAXSCALES];
  const float* W[B0_MAXSCALES]; const float* scale[B0_MAXSCALES]; const float* shift[B0_MAXSCALES];
  float* out;   // [P, nscales*64]
};

// one CTA per point, nscales*64 threads: thread (s, c) owns channel c of scale s
__global__ void fd_block0_kernel(const Block0Args a) {
  __shared__ float e[B0_MAXK][6];
  const int64_t pt = blockIdx.x;
  const int64_t patch0 = (pt / a.Mpts) * a.Mpts;
  const int kmax = a.ks[a.nscales - 1];
  const float cx = a.xyz[3 * pt], cy = a.xyz[3 * pt + 1], cz = a.xyz[3 * pt + 2];
  for (int j = threadIdx.x; j < kmax; j += blockDim.x) {
    const int64_t nb = patch0 + a.idx[pt * a.ldi + j];
    const float nx = a.xyz[3 * nb], ny = a.xyz[3 * nb + 1], nz = a.xyz[3 * nb + 2];
    e[j][0] = __fsub_rn(nx, cx); e[j][1] = __fsub_rn(ny, cy); e[j][2] = __fsub_rn(nz, cz);
    e[j][3] = nx; e[j][4] = ny; e[j][5] = nz;
  }
  __syncthreads();
  const int s = threadIdx.x >> 6, c = threadIdx.x & 63;
  const float* w = a.W[s] + 6 * c;
  const float w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4], w5 = w[5];
  const float sc = a.scale[s][c], sh = a.shift[s][c];
  float m = -INFINITY;
  const int ks = a.ks[s];
  for (int j = 0; j < ks; ++j) {
    float y = __fmul_rn(w0, e[j][0]);
    y = fmaf(w1, e[j][1], y); y = fmaf(w2, e[j][2], y); y = fmaf(w3, e[j][3], y);
    y = fmaf(w4, e[j][4], y); y = fmaf(w5, e[j][5], y);
    y = __fadd_rn(__fmul_rn(y, sc), sh);
    m = fmaxf(m, act_leaky(y));
  }
  a.out[pt * (a.nscales * 64) + threadIdx.x] = m;
}

// Factorised EdgeConv (tensor-core mode):  W cat(x_j - x_i, x_j) = (Wa+Wb) x_j - Wa x_i =: P_j - Q_i, PQ rows = (P | Q).
// Fused tail of such a block: for one patch and a slab of 128 channels,
//   u[i,c] = max_j LeakyReLU(scale_c (P[nb_ij,c] - Q[i,c]) + shift_c)   (P rows staged in shared memory)
// followed by the block's T-step EIF/LIF recurrence, every step's spike written into the [point, t, 960] spike
// tensor (no HBM round trip between the max-pool and the recurrence).
constexpr int EGU_PTS = 4;     // points processed together per thread (ILP for the recurrence)
constexpr int EGU_THREADS = 256;  // 64 channel pairs x 4 quarters of the patch's points
template <bool EIF>
__global__ void __launch_bounds__(EGU_THREADS)
edge_gather_unroll_kernel(const float* __restrict__ PQ, int C, const int32_t* __restrict__ idx, int kk, int Mpts,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ np, const float* __restrict__ ep, int T,
                          float* __restrict__ U, float* __restrict__ spk, int64_t ldspk_row, int ldo,
                          int h2, float* __restrict__ spk0, int64_t plane, int choff) {
  // h2: spikes go out as fp16 (hi, lo) planes of s * 2^13, row (point*T + t) of `ldo` halfs, for the fp16x3 conv5; the
  // step-0 spikes -- the only ones the next block's graph and EdgeConv read -- also as fp32 into spk0 [point, ldo].
  // A thread owns two adjacent channels: one 8-byte shared-memory load per neighbour serves both (the gather is
  // instruction-bound), and every global store is a 4- or 8-byte pair.
  extern __shared__ float egs[];
  float* Ps = egs;                                                 // [Mpts][128]
  uint8_t* nbs = reinterpret_cast<uint8_t*>(egs + Mpts * 128);     // [Mpts][kk] local neighbour indices (Mpts <= 256)
  const int64_t patch0 = (int64_t)blockIdx.x * Mpts;
  {
    const int tx = threadIdx.x & 127, hf = threadIdx.x >> 7;
    const int cs = blockIdx.y * 128 + tx;                          // C is a multiple of 128
    for (int m = hf; m < Mpts; m += 2) Ps[m * 128 + tx] = PQ[(patch0 + m) * 2 * C + cs];
  }
  for (int e = threadIdx.x; e < Mpts * kk; e += EGU_THREADS) nbs[e] = (uint8_t)idx[patch0 * kk + e];
  __syncthreads();
  const int tp = threadIdx.x & 63, quarter = threadIdx.x >> 6;
  const int c = blockIdx.y * 128 + 2 * tp;                         // channels c, c + 1
  const float sc0 = scale[c], sh0 = shift[c], sc1 = scale[c + 1], sh1 = shift[c + 1];
  FastNeuronK k0, k1;
  {
    const NeuronParams p0{np[c], np[C + c], np[2 * C + c], np[3 * C + c]};
    const NeuronParams p1{np[c + 1], np[C + c + 1], np[2 * C + c + 1], np[3 * C + c + 1]};
    EifParams q0{1.0f, 1.0f}, q1{1.0f, 1.0f};
    if (EIF) { q0.dT = ep[c]; q0.thrh = ep[C + c]; q1.dT = ep[c + 1]; q1.thrh = ep[C + c + 1]; }
    k0 = fast_neuron_k(p0, q0); k1 = fast_neuron_k(p1, q1);
  }
  const int mq = (Mpts + 3) >> 2;
  const int i_begin = quarter * mq, i_end = min(Mpts, i_begin + mq);
  // output bases of this patch for this thread's channel pair (offsets inside a patch fit 32 bits: Mpts * T * ldo halfs, checked on the host)
  __half* const hb = reinterpret_cast<__half*>(spk) + patch0 * T * (int64_t)ldo + choff + c;
  __half* const lb = hb + plane;
  float* const s0b = spk0 + patch0 * (int64_t)ldo + choff + c;
  float* const fb = spk + patch0 * ldspk_row + c;
  const float2* Pc = reinterpret_cast<const float2*>(Ps) + tp;      // row stride 64 float2
  constexpr int NP = EGU_PTS / 2;                                  // points per pass: NP x 2 channels = EGU_PTS recurrences
  for (int i0 = i_begin; i0 < i_end; i0 += NP) {
    float u[EGU_PTS], m[EGU_PTS], th[EGU_PTS], rho[EGU_PTS];        // [2a] channel c, [2a+1] channel c+1 of point i0+a
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      const int i = min(i0 + a, i_end - 1);
      const float2 qv = *reinterpret_cast<const float2*>(PQ + (patch0 + i) * 2 * C + C + c);
      // LeakyReLU(scale*(P-Q)+shift) is monotone in P for scale >= 0 and antitone otherwise: the max over the
      // neighbours is reached at max P or min P; both are tracked in one pass over the (byte-packed) graph row
      float mx0 = -INFINITY, mn0 = INFINITY, mx1 = -INFINITY, mn1 = INFINITY;
      const uint8_t* nr = nbs + i * kk;
      if ((kk & 3) == 0) {
        const uint32_t* nw = reinterpret_cast<const uint32_t*>(nr);
        for (int w = 0; w < (kk >> 2); ++w) {
          const uint32_t v = nw[w];
          const float2 p0 = Pc[(v & 255u) * 64], p1 = Pc[((v >> 8) & 255u) * 64], p2 = Pc[((v >> 16) & 255u) * 64], p3 = Pc[(v >> 24) * 64];
          mx0 = fmaxf(fmaxf(mx0, p0.x), fmaxf(p1.x, fmaxf(p2.x, p3.x)));
          mn0 = fminf(fminf(mn0, p0.x), fminf(p1.x, fminf(p2.x, p3.x)));
          mx1 = fmaxf(fmaxf(mx1, p0.y), fmaxf(p1.y, fmaxf(p2.y, p3.y)));
          mn1 = fminf(fminf(mn1, p0.y), fminf(p1.y, fminf(p2.y, p3.y)));
        }
      } else {
        for (int j = 0; j < kk; ++j) {
          const float2 pv = Pc[nr[j] * 64];
          mx0 = fmaxf(mx0, pv.x); mn0 = fminf(mn0, pv.x); mx1 = fmaxf(mx1, pv.y); mn1 = fminf(mn1, pv.y);
        }
      }
      u[2 * a] = act_leaky(fmaf((sc0 < 0.0f ? mn0 : mx0) - qv.x, sc0, sh0));
      u[2 * a + 1] = act_leaky(fmaf((sc1 < 0.0f ? mn1 : mx1) - qv.y, sc1, sh1));
      if (i0 + a < i_end) *reinterpret_cast<float2*>(U + (patch0 + i) * C + c) = make_float2(u[2 * a], u[2 * a + 1]);
    }
    float s[EGU_PTS];
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      s[2 * a] = neuron_step_fast<EIF, true>(u[2 * a], m[2 * a], th[2 * a], rho[2 * a], k0);
      s[2 * a + 1] = neuron_step_fast<EIF, true>(u[2 * a + 1], m[2 * a + 1], th[2 * a + 1], rho[2 * a + 1], k1);
    }
    // stores: per-thread 64-bit bases of the PATCH (hb / lb / s0b / fb, set up once per CTA) + 32-bit offsets inside it that
    // advance by one row per step -- one multiply-add per address instead of a 64-bit (row * T + t) * ldo product per store
    uint32_t offs[NP];
#pragma unroll
    for (int a = 0; a < NP; ++a) offs[a] = h2 ? (uint32_t)(i0 + a) * (uint32_t)T * (uint32_t)ldo : (uint32_t)(i0 + a) * (uint32_t)ldspk_row;
    auto put = [&](int a, bool first, float v0, float v1) {
      if (h2) {
        const float y0 = v0 * 8192.0f, y1 = v1 * 8192.0f;
        const __half2 hv = __floats2half2_rn(y0, y1);
        const float2 hf = __half22float2(hv);
        *reinterpret_cast<__half2*>(hb + offs[a]) = hv;
        if (plane) *reinterpret_cast<__half2*>(lb + offs[a]) = __floats2half2_rn(y0 - hf.x, y1 - hf.y);   // plane == 0: hi plane only (fast mode)
        if (first) *reinterpret_cast<float2*>(s0b + (uint32_t)(i0 + a) * (uint32_t)ldo) = make_float2(v0, v1);
      } else {
        *reinterpret_cast<float2*>(fb + offs[a]) = make_float2(v0, v1);
      }
      offs[a] += (uint32_t)ldo;
    };
#pragma unroll
    for (int a = 0; a < NP; ++a)
      if (i0 + a < i_end) put(a, true, s[2 * a], s[2 * a + 1]);
    for (int t = 1; t < T; ++t) {
#pragma unroll
      for (int a = 0; a < NP; ++a) {
        s[2 * a] = neuron_step_fast<EIF, false>(0.0f, m[2 * a], th[2 * a], rho[2 * a], k0);
        s[2 * a + 1] = neuron_step_fast<EIF, false>(0.0f, m[2 * a + 1], th[2 * a + 1], rho[2 * a + 1], k1);
      }
#pragma unroll
      for (int a = 0; a < NP; ++a)
        if (i0 + a < i_end) put(a, false, s[2 * a], s[2 * a + 1]);
    }
  }
}

int launch_edge_gather_unroll(bool eif, const float* PQ, int C, const int32_t* idx, int kk, int Mpts, int64_t S,
                              const float* scale, const float* shift, const float* np, const float* ep, int T, float* U,
                              float* spk, int64_t ldspk_row, int ldo, cudaStream_t st, bool h2, float* spk0, int64_t plane, int choff) {
  SAPCU_REQUIRE(C % 128 == 0, "edge_gather_unroll: C=%d must be a multiple of 128", C);
  SAPCU_REQUIRE(Mpts >= 1 && Mpts <= 256, "edge_gather_unroll: M=%d outside [1,256]", Mpts);
  SAPCU_REQUIRE((int64_t)Mpts * T * ldo < ((int64_t)1 << 31) && (int64_t)Mpts * ldspk_row < ((int64_t)1 << 31), "edge_gather_unroll: patch rows exceed 32-bit offsets");
  if (S == 0) return 0;
  const size_t smem = sizeof(float) * (size_t)Mpts * 128 + (((size_t)Mpts * kk + 15) & ~(size_t)15);
  static PerDeviceOnce once;
  {
    const int rc = once.run([]() -> int {
      SAPCU_CUDA_CHECK(cudaFuncSetAttribute(edge_gather_unroll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      SAPCU_CUDA_CHECK(cudaFuncSetAttribute(edge_gather_unroll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      return 0;
    });
    if (rc) return rc;
  }
  SAPCU_REQUIRE(smem <= 160 * 1024, "edge_gather_unroll: patch too large for shared memory");
  dim3 grid((unsigned)S, (unsigned)(C / 128));
  if (eif) edge_gather_unroll_kernel<true><<<grid, EGU_THREADS, smem, st>>>(PQ, C, idx, kk, Mpts, scale, shift, np, ep, T, U, spk, ldspk_row, ldo, h2 ? 1 : 0, spk0, plane, choff);
  else     edge_gather_unroll_kernel<false><<<grid, EGU_THREADS, smem, st>>>(PQ, C, idx, kk, Mpts, scale, shift, np, ep, T, U, spk, ldspk_row, ldo, h2 ? 1 : 0, spk0, plane, choff);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

// out[row*(T*ldo) + t*ldo + c] = spike at step t of channel c for input U[row*ldu + c]   (all_steps)
// out[row*ldo + c]             = spike at step T-1                                        (!all_steps)
template <bool EIF, bool PRECISE>
__global__ void neuron_unroll_kernel(const float* __restrict__ U, int64_t ldu, int64_t rows, int C, int T,
                                     const float* __restrict__ np, const float* __restrict__ ep, int all_steps,
                                     float* __restrict__ out, int64_t ldo, int h2, float* __restrict__ out0, int64_t plane, int choff) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const int64_t row = e / C;
  const int c = (int)(e - row * C);
  const NeuronParams p{np[c], np[C + c], np[2 * C + c], np[3 * C + c]};
  EifParams q{1.0f, 1.0f};
  if (EIF) { q.dT = ep[c]; q.thrh = ep[C + c]; }
  NeuronState st = neuron_init(p);
  float s = U[row * ldu + c];
  float* o = all_steps ? out + row * (int64_t)T * ldo + c : out + row * ldo + c;
  for (int t = 0; t < T; ++t) {
    s = neuron_step<EIF, PRECISE>(s, st, p, q);
    if (h2) {                                            // fp16 planes for conv5 + fp32 step-0 copy (see edge_gather_unroll)
      __half* hp = reinterpret_cast<__half*>(out) + (row * T + t) * ldo + choff + c;
      const float ys = s * 8192.0f;
      const __half hv = __float2half_rn(ys);
      hp[0] = hv;
      if (plane) hp[plane] = __float2half_rn(ys - __half2float(hv));   // plane == 0: hi plane only (fast mode)
      if (t == 0) out0[row * ldo + choff + c] = s;
    } else if (all_steps) o[(int64_t)t * ldo] = s;
  }
  if (!all_steps) *o = s;
}

template <bool PRECISE>
__global__ void temporal_lif_kernel(const float* __restrict__ pool, int64_t S, int Tt, int C,
                                    const float* __restrict__ wsm, const float* __restrict__ np,
                                    float* __restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= S * C) return;
  const int64_t s = e / C;
  const int c = (int)(e - s * C);
  float z = 0.0f;
  for (int t = 0; t < Tt; ++t) z = fmaf(wsm[t], pool[(s * Tt + t) * C + c], z);
  const NeuronParams p{np[c], np[C + c], np[2 * C + c], np[3 * C + c]};
  NeuronState st = neuron_init(p);
  EifParams q{1.0f, 1.0f};
  out[e] = neuron_step<false, PRECISE>(z, st, p, q);
}

// qkv: [S, 3*dim]; out: [S, dim]; softmax over heads of (q_h . k_h) * scale, then a_h * v_h
__global__ void head_attention_kernel(const float* __restrict__ qkv, int64_t S, int H, int hd, float scale,
                                      float* __restrict__ out) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int dim = H * hd;
  const float* q = qkv + s * 3 * dim; const float* k = q + dim; const float* v = k + dim;
  float a[16];
  float mx = -INFINITY;
  for (int h = 0; h < H; ++h) {
    float d = 0.0f;
    for (int i = 0; i < hd; ++i) d = fmaf(q[h * hd + i], k[h * hd + i], d);
    a[h] = d * scale; mx = fmaxf(mx, a[h]);
  }
  float sum = 0.0f;
  for (int h = 0; h < H; ++h) { a[h] = expf(a[h] - mx); sum += a[h]; }
  for (int h = 0; h < H; ++h) {
    const float w = a[h] / sum;
    for (int i = 0; i < hd; ++i) out[s * dim + h * hd + i] = w * v[h * hd + i];
  }
}

// one warp per row, in place allowed
__global__ void layernorm_rows_kernel(const float* __restrict__ X, int64_t S, int C, const float* __restrict__ w,
                                      const float* __restrict__ b, float* __restrict__ out) {
  const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= S) return;
  float sum = 0.0f;
  for (int c = lane; c < C; c += 32) sum += X[s * C + c];
  for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float mean = sum / (float)C;
  float var = 0.0f;
  for (int c = lane; c < C; c += 32) { const float d = X[s * C + c] - mean; var = fmaf(d, d, var); }
  for (int off = 16; off; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
  const float rstd = 1.0f / sqrtf(var / (float)C + 1e-5f);
  for (int c = lane; c < C; c += 32) out[s * C + c] = (X[s * C + c] - mean) * rstd * w[c] + b[c];
}

__global__ void fd_tail_kernel(const float* __restrict__ H, int64_t S, int K, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ out) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= S) return;
  float y = 0.0f;
  for (int c = 0; c < K; ++c) y = fmaf(w[c], H[s * K + c], y);
  y += b[0];
  const float beta = 5.0f;
  const float z = y * beta;
  out[s] = (z > 20.0f) ? y : log1pf(expf(z)) / beta;
}

int launch_fd_block0(const float* xyz, const int32_t* idx, int ldi, int Mpts, int64_t P, int nscales, const int* ks,
                     const float* const* W, const float* const* scale, const float* const* shift, float* out,
                     cudaStream_t st) {
  SAPCU_REQUIRE(nscales >= 1 && nscales <= B0_MAXSCALES, "fd_block0: %d scales unsupported", nscales);
  Block0Args a;
  a.xyz = xyz; a.idx = idx; a.ldi = ldi; a.Mpts = Mpts; a.P = P; a.nscales = nscales; a.out = out;
  for (int s = 0; s < nscales; ++s) {
    SAPCU_REQUIRE(ks[s] <= B0_MAXK && (s == 0 || ks[s] >= ks[s - 1]), "fd_block0: k_scales must be ascending and <= %d", B0_MAXK);
    a.ks[s] = ks[s]; a.W[s] = W[s]; a.scale[s] = scale[s]; a.shift[s] = shift[s];
  }
  if (P == 0) return 0;
  fd_block0_kernel<<<(unsigned)P, nscales * 64, 0, st>>>(a);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

int launch_neuron_unroll(bool eif, bool precise, const float* U, int64_t ldu, int64_t rows, int C, int T,
                         const float* np, const float* ep, int all_steps, float* out, int64_t ldo, cudaStream_t st,
                         bool h2, float* out0, int64_t plane, int choff) {
  if (rows == 0) return 0;
  const unsigned grid = (unsigned)ceil_div(rows * C, 256);
#define SAPCU_NU(E, P) neuron_unroll_kernel<E, P><<<grid, 256, 0, st>>>(U, ldu, rows, C, T, np, ep, all_steps, out, ldo, h2 ? 1 : 0, out0, plane, choff)
  if (eif) { if (precise) SAPCU_NU(true, true); else SAPCU_NU(true, false); }
  else     { if (precise) SAPCU_NU(false, true); else SAPCU_NU(false, false); }
#undef SAPCU_NU
  SAPCU_LAUNCH_CHECK();
  return 0;
}

int launch_temporal_lif(bool precise, const float* pool, int64_t S, int Tt, int C, const float* wsm, const float* np,
                        float* out, cudaStream_t st) {
  if (S == 0) return 0;
  const unsigned grid = (unsigned)ceil_div(S * C, 256);
  if (precise) temporal_lif_kernel<true><<<grid, 256, 0, st>>>(pool, S, Tt, C, wsm, np, out);
  else         temporal_lif_kernel<false><<<grid, 256, 0, st>>>(pool, S, Tt, C, wsm, np, out);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

int launch_head_attention(const float* qkv, int64_t S, int H, int hd, float scale, float* out, cudaStream_t st) {
  SAPCU_REQUIRE(H <= 16, "head_attention: %d heads > 16", H);
  if (S == 0) return 0;
  head_attention_kernel<<<(unsigned)ceil_div(S, 128), 128, 0, st>>>(qkv, S, H, hd, scale, out);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

int launch_layernorm_rows(const float* X, int64_t S, int C, const float* w, const float* b, float* out, cudaStream_t st) {
  if (S == 0) return 0;
  layernorm_rows_kernel<<<(unsigned)ceil_div(S * 32, 256), 256, 0, st>>>(X, S, C, w, b, out);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

int launch_fd_tail(const float* H, int64_t S, int K, const float* w, const float* b, float* out, cudaStream_t st) {
  if (S == 0) return 0;
  fd_tail_kernel<<<(unsigned)ceil_div(S, 128), 128, 0, st>>>(H, S, K, w, b, out);
  SAPCU_LAUNCH_CHECK();
  return 0;
}

}  // namespace sapcu
