// Tabulated LIF^T chains (SAPCU_MODE_FAST).
//
// In eval mode the soft spike is strictly positive, so the refractory gate of the reference neuron
// (fn/snn_coder.py:109-133) closes after step 0 and `for t in range(T): s, *st = lif(s, *st)` is a smooth scalar function
// F_c(u) of the chain's first input per channel c: |F'| < 0.2, |F''| < 0.3 over the whole clamp range of
// the neuron parameters.  The fast mode evaluates F_c from a per-channel piecewise-cubic table held in shared memory
// (one LDS.U16 + one LDS.128 + ~20 ALU instructions) instead of T x (12 FP + 3 MUFU) instructions.
//
// Grid: x = u - theta0_c, y = 1 + |x|.  A CELL is one binade of y on one side of theta0 (2 * LT_NB cells per channel); each
// cell is split into 2^k equal segments, k in [0, 7] chosen per cell on the host so that the cubic (Chebyshev-node
// interpolant of the exact fp64 chain) stays within LT_TOL of the exact chain.  The float bits of y give cell, segment and
// the local coordinate without any transcendental:  cell = exponent(y), segment = top k mantissa bits, tau = y - floor_k(y).
// |x| >= 2^LT_NB - 1 = 255 (not seen behind a BatchNorm with sane statistics) takes the exact MUFU chain.  Channels of a
// block with identical neuron parameters (every channel of a default-initialised layer) share their segments.
//
// Memory image of one 128-channel block (copied verbatim to shared memory by the kernels):
//   uint2  desc[LT_NCELL][128]   .x = 23 - k (the shift that turns the float bits of y into a segment number), .y = byte offset
//                                of the cell's (virtual) segment 0 relative to the coefficient array, biased so that
//                                segment address = coef + .y + 16 * (bits(y) >> .x) needs no masking of the exponent; the 32
//                                channels of a warp read 256 contiguous bytes when they sit in the same cell
//   float4 coef[nseg]            s = c.x + tau*(c.y + tau*(c.z + tau*c.w)); channel c's segments start at an index that is
//                                congruent to c mod 8, so the 8 lanes of an LDS.128 phase that sit in the same relative
//                                segment hit 8 different bank groups
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

namespace sapcu {

constexpr int LT_NB = 8;                     // binades per side: |x| < 255
constexpr int LT_NCELL = 2 * LT_NB;
constexpr int LT_KMAX = 7;                   // up to 128 segments per cell
constexpr int LT_CH = 128;                   // channels per block (= UMMA M per CTA)
constexpr uint32_t LT_DESC_BYTES = LT_CH * LT_NCELL * 8;
constexpr double LT_TOL = 4e-5;              // acceptance bound of the fit, absolute, on soft spikes in (0, 0.7)
constexpr uint32_t LT_SMEM_BUDGET = 120 * 1024;   // a block above this keeps its layer on the MUFU path
constexpr uint32_t LT_SMEM_BUDGET_TC = 96 * 1024; // same next to the two 64 KiB stages of the parity-grade (fp16x3) kernels

struct LifTableBlock { size_t off_bytes = 0; uint32_t bytes = 0; uint32_t nseg = 0; };
struct LifTableHost {
  int C = 0, T = 0;
  std::vector<LifTableBlock> blocks;         // ceil(C / 128)
  std::vector<uint8_t> image;                // concatenated block images (each 256-byte aligned)
  double max_err = 0.0;                      // largest |cubic - exact chain| seen at the acceptance points
  uint32_t max_block_bytes = 0;
  bool usable = false;                       // every block fits LT_SMEM_BUDGET
};
// np4 = [4][C] rows {d, a, r, theta0}, already clamped; exact chain evaluated in fp64 on the host
void lif_table_build(const float* np4, int C, int T, LifTableHost* out);
// exact T-step chain from the zero state in fp64 (the reference's formula, clamps included)
double lif_chain_exact_host(double u, double d, double a, double r, double th0, int T);

#ifdef __CUDACC__
// NV chains of ONE channel, phase-major so the NV descriptor loads, then the NV coefficient loads, are in flight together
// (no branch between them).  x[i] = u_i - theta0 on entry, the soft spike on exit.  desc_c: this channel's column of the
// block's descriptor array (cell stride LT_CH); coef: the block's coefficient array; both in shared memory.  Inputs outside
// the tabulated range (|x| >= 255, NaN) are clamped for the lookup and flagged in the returned bit mask: the caller
// re-evaluates those with the exact chain.  The index arithmetic is laid out for the SM's two integer-capable pipes: 7
// ALU-pipe operations (min, compare, 4 shifts, and-mask), the rest multiply-adds.
template <int NV>
__device__ __forceinline__ uint32_t lif_table_eval_vec(float (&x)[NV], const uint2* __restrict__ desc_c,
                                                       const float4* __restrict__ coef) {
  float y[NV]; uint2 d[NV]; uint32_t oob = 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float ya = fabsf(x[i]) + 1.0f;
    if (!(ya < 256.0f)) oob |= 1u << i;
    y[i] = fminf(ya, 255.99998f);                                       // NaN -> 255.99998 as well
    const uint32_t cell = (__float_as_uint(y[i]) >> 23) * (uint32_t)LT_CH + (__float_as_uint(x[i]) >> 31) * (uint32_t)(LT_NB * LT_CH);
    d[i] = desc_c[(int)cell - 127 * LT_CH];
  }
  float4 c[NV]; float tau[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t yb = __float_as_uint(y[i]);
    const uint32_t t1 = yb >> d[i].x;
    c[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(coef) + (int)(t1 * 16u + d[i].y));
    tau[i] = y[i] - __uint_as_float(yb & (0xFFFFFFFFu << d[i].x));
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) x[i] = fmaf(fmaf(fmaf(c[i].w, tau[i], c[i].z), tau[i], c[i].y), tau[i], c[i].x);
  return oob;
}
#endif

}  // namespace sapcu
