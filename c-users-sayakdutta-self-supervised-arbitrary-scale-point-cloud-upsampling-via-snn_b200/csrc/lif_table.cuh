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
//   uint16 desc[128][LT_NCELL]   (k << 13) | first segment of the cell (relative to the block's coefficient array)
//   float4 coef[nseg]            s = c.x + tau*(c.y + tau*(c.z + tau*c.w))
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

namespace sapcu {

constexpr int LT_NB = 8;                     // binades per side: |x| < 255
constexpr int LT_NCELL = 2 * LT_NB;
constexpr int LT_KMAX = 7;                   // up to 128 segments per cell
constexpr int LT_CH = 128;                   // channels per block (= UMMA M per CTA)
constexpr uint32_t LT_DESC_BYTES = LT_CH * LT_NCELL * 2;
constexpr double LT_TOL = 4e-5;              // acceptance bound of the fit, absolute, on soft spikes in (0, 0.7)
constexpr uint32_t LT_SMEM_BUDGET = 120 * 1024;   // a block above this keeps its layer on the MUFU path

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
// desc_c: this channel's LT_NCELL descriptors; coef: the block's coefficient array (both in shared memory).
// Returns false when u is outside the tabulated range (caller evaluates the exact chain).
__device__ __forceinline__ bool lif_table_eval(float u, float th0, const uint16_t* __restrict__ desc_c,
                                               const float4* __restrict__ coef, float& s) {
  const float x = u - th0;
  const float y = fabsf(x) + 1.0f;
  if (!(y < 256.0f)) return false;                                       // also NaN
  const uint32_t yb = __float_as_uint(y);
  const uint32_t cell = (yb >> 23) - 127u + (__float_as_uint(x) >> 31) * (uint32_t)LT_NB;
  const uint32_t d = desc_c[cell];
  const uint32_t sh = 23u - (d >> 13);
  const uint32_t seg = (d & 0x1FFFu) + ((yb & 0x7FFFFFu) >> sh);
  const float tau = y - __uint_as_float(yb & (0xFFFFFFFFu << sh));
  const float4 c = coef[seg];
  s = fmaf(fmaf(fmaf(c.w, tau, c.z), tau, c.y), tau, c.x);
  return true;
}
#endif

}  // namespace sapcu
