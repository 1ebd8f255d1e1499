// Tabulated LIF^T chains (SAPCU_MODE_FAST).
//
// In eval mode the soft spike is strictly positive, so the refractory gate of the reference neuron
// (fn/snn_coder.py:109-133) closes after step 0 and `for t in range(T): s, *st = lif(s, *st)` is a smooth scalar function
// F_c(u) of the chain's first input per channel c: |F'| < 0.2, |F''| < 0.3 over the whole clamp range of
// the neuron parameters.  The fast mode evaluates F_c from a per-channel piecewise-cubic table held in shared memory
// (one LDS.U16 + one LDS.128 + ~20 ALU instructions) instead of T x (12 FP + 3 MUFU) instructions.
//
// Grid: x = u - theta0_c, y = 2|x| + 2 in [2, 512).  A CELL is one binade of y on one side of theta0 (2 * LT_NB cells per
// channel; the low three exponent bits of y are the cell number); each cell is split into 2^k equal segments, k in [0, 7]
// chosen per cell on the host so that the cubic (Chebyshev-node interpolant of the exact fp64 chain, in the segment's own
// coordinate t in [0, 1]) stays within LT_TOL of the exact chain.  Segment and t come out of two multiply-adds: with the
// cell's scale S = 2^(k-e-1), q = y*S lies in [2^k, 2^(k+1)) and  qm = fma(y, S, 2^23 - 0.5)  is floor(q) + 2^23 (an exact
// tie rounds to the neighbouring segment's end point, where the piecewise function is continuous), t = fma(y, S, -(qm - 2^23)).
// Everything but the clamp, the cell address (shift + and-or) and the sign test runs on the FMA pipe -- the first version
// of this lookup (shift / mask arithmetic on bits(y)) saturated the ALU pipe.
// |x| >= 255 (not seen behind a BatchNorm with sane statistics) and NaN take the exact MUFU chain.  Channels of a block with
// identical neuron parameters (every channel of a default-initialised layer) share their segments.
//
// Memory image of one 128-channel block (copied verbatim to shared memory by the kernels):
//   uint2  desc[LT_NCELL][128]   .x = bits of S, .y = byte offset (mod 2^32) such that the segment's coefficients sit at
//                                coef + .y + 16 * bits(qm); the 32 channels of a warp read 256 contiguous bytes when they
//                                sit in the same cell; cells 0..7: x >= 0, cells 8..15: x < 0
//   float4 coef[nseg]            s = c.x + t*(c.y + t*(c.z + t*c.w)); every side starts with one guard segment (the constant
//                                F(theta0)) that catches the downward tie at q = 2^k of its first cell; channel c's segments
//                                start at an index congruent to c mod 8, so the 8 lanes of an LDS.128 phase that sit in the
//                                same relative segment hit 8 different bank groups
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
// host restatement of the device lookup for channel c at x = u - theta0 (|x| < 255); NaN if the lookup leaves the table
float lif_table_eval_host(const LifTableHost& t, int c, float x);
// exact T-step chain from the zero state in fp64 (the reference's formula, clamps included)
double lif_chain_exact_host(double u, double d, double a, double r, double th0, int T);

#ifdef __CUDACC__
// NV chains of ONE channel, phase-major so the NV descriptor loads, then the NV coefficient loads, are in flight together
// (no branch between them).  x[i] = u_i - theta0 on entry, the soft spike on exit.  desc_lane: shared-memory address of this
// channel's entry of cell 0 (cell stride LT_CH * 8 bytes); coef: shared-memory address of the block's coefficient array.
// Returns true when some input lies outside the tabulated range (|x| >= 255) or is NaN: the caller then re-evaluates
// every element with |x| >= 255 / NaN by the exact chain (the lookup clamps, so its loads stay inside the table).
// Copy one block image into shared memory with `nthreads` threads (thread `tid`), relocating the descriptors' byte offsets
// to absolute shared-memory addresses so that the lookup needs no base add.
__device__ __forceinline__ void lif_table_load(const uint8_t* __restrict__ src, uint8_t* dst, uint32_t bytes, uint32_t tid, uint32_t nthreads) {
  const uint32_t coef = (uint32_t)__cvta_generic_to_shared(dst) + LT_DESC_BYTES;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (uint32_t i = tid; i < bytes / 16; i += nthreads) {
    uint4 v = s4[i];
    if (i < LT_DESC_BYTES / 16) { v.y += coef; v.w += coef; }           // two descriptors per 16 bytes
    d4[i] = v;
  }
}

template <int NV>
__device__ __forceinline__ bool lif_table_eval_vec(float (&x)[NV], uint32_t desc_lane) {
  float y[NV]; uint32_t dS[NV], dO[NV];
  float ymax = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float y2 = fmaf(fabsf(x[i]), 2.0f, 2.0f);
    y[i] = fminf(y2, 511.99997f);                                       // NaN -> 511.99997 as well
    const uint32_t a = desc_lane + ((__float_as_uint(y[i]) >> 13) & 0x1C00u);     // cell = low 3 exponent bits, 1 KiB per cell
    // predicated pair of loads instead of a select on the address: the sign costs one compare on the ALU pipe
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %2, 0f00000000;\n\t"
        "@p ld.shared.v2.u32 {%0, %1}, [%3+8192];\n\t"
        "@!p ld.shared.v2.u32 {%0, %1}, [%3];\n\t}"
        : "=r"(dS[i]), "=r"(dO[i]) : "f"(x[i]), "r"(a));
  }
  // out of range or NaN <=> the clamp bit: y == 511.99997 (3-input maxima: half an instruction per element)
#pragma unroll
  for (int i = 0; i + 1 < NV; i += 2) ymax = fmaxf(fmaxf(ymax, y[i]), y[i + 1]);
  if (NV & 1) ymax = fmaxf(ymax, y[NV - 1]);
  float4 c[NV]; float t[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float S = __uint_as_float(dS[i]);
    const float qm = fmaf(y[i], S, 8388607.5f);                         // floor(y * S) + 2^23
    t[i] = fmaf(y[i], S, -(qm - 8388608.0f));
    const uint32_t addr = __float_as_uint(qm) * 16u + dO[i];            // the kernels add the coefficient array's address to .y while copying the table
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c[i].x), "=f"(c[i].y), "=f"(c[i].z), "=f"(c[i].w) : "r"(addr));
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) x[i] = fmaf(fmaf(fmaf(c[i].w, t[i], c[i].z), t[i], c[i].y), t[i], c[i].x);
  return ymax >= 511.99997f;
}
__device__ __forceinline__ bool lif_table_oob(float x) { return !(fabsf(x) < 255.0f); }
#endif

}  // namespace sapcu
