// Host-side builder of the tabulated LIF^T chains (see lif_table.cuh).  Runs once per model at sapcu_model_finalize.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <array>
#include <map>
#include <thread>
#include "lif_table.cuh"

namespace sapcu {

// One soft spike (fn/snn_coder.py:135-153) in fp64, clamps included.
static inline double spike_exact(double v) {
  const double vc = v < -10.0 ? -10.0 : (v > 10.0 ? 10.0 : v);
  return 0.5 * exp(-(vc * vc) / 2.0) / 2.5066282746310002 + 0.5 / (1.0 + exp(-10.0 * vc));
}

double lif_chain_exact_host(double u, double d, double a, double r, double th0, int T) {
  double m = 0.0, th = th0, rho = 0.0, x = u, s = 0.0;
  for (int t = 0; t < T; ++t) {
    x = rho <= 0.0 ? x : 0.0;                       // x * float(rho <= 0)
    m = m * d * (1.0 - rho) + x;
    s = spike_exact(m - th);
    m = m * (1.0 - s);
    rho = rho * r + s;
    th = th + a * s;
    th = th0 + (th - th0) * 0.95;
    x = s;
  }
  return s;
}

namespace {

struct ChanTab {
  uint16_t k[LT_NCELL];
  std::vector<std::array<float, 4>> coef;          // cells in order, segments in order inside a cell
  double err = 0.0;
};

// Chebyshev nodes on (0, 1) and the inverse Vandermonde matrix mapping node values to monomial coefficients in t
struct Cheb {
  double t[4]; double inv[4][4];
  Cheb() {
    for (int i = 0; i < 4; ++i) t[i] = 0.5 * (1.0 + cos((2 * i + 1) * M_PI / 8.0));
    double V[4][8];
    for (int i = 0; i < 4; ++i) {
      double p = 1.0;
      for (int j = 0; j < 4; ++j) { V[i][j] = p; p *= t[i]; V[i][4 + j] = i == j ? 1.0 : 0.0; }
    }
    for (int c = 0; c < 4; ++c) {                  // Gauss-Jordan with partial pivoting
      int piv = c;
      for (int r = c + 1; r < 4; ++r) if (fabs(V[r][c]) > fabs(V[piv][c])) piv = r;
      for (int j = 0; j < 8; ++j) std::swap(V[c][j], V[piv][j]);
      const double d = V[c][c];
      for (int j = 0; j < 8; ++j) V[c][j] /= d;
      for (int r = 0; r < 4; ++r) if (r != c) { const double f = V[r][c]; for (int j = 0; j < 8; ++j) V[r][j] -= f * V[c][j]; }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[i][j] = V[i][4 + j];
  }
};
static const Cheb g_cheb;

// cubic in t in [0, 1] on the segment y in [y0, y0 + w) (y = 2|x| + 2) of side `sgn`; returns the largest deviation from the
// exact chain at 9 check points, evaluated with the device's fp32 Horner form
double fit_segment(double sgn, double y0, double w, double d, double a, double r, double th0, int T, float (&c)[4]) {
  auto u_of = [&](double y) { return th0 + sgn * (y * 0.5 - 1.0); };
  double f[4];
  for (int i = 0; i < 4; ++i) f[i] = lif_chain_exact_host(u_of(y0 + w * g_cheb.t[i]), d, a, r, th0, T);
  for (int j = 0; j < 4; ++j) {
    double s = 0.0;
    for (int i = 0; i < 4; ++i) s += g_cheb.inv[j][i] * f[i];
    c[j] = (float)s;
  }
  double worst = 0.0;
  for (int q = 0; q <= 8; ++q) {
    const float t = (float)(q / 8.0);
    const float p = fmaf(fmaf(fmaf(c[3], t, c[2]), t, c[1]), t, c[0]);
    const double e = fabs((double)p - lif_chain_exact_host(u_of(y0 + w * (double)t), d, a, r, th0, T));
    worst = e > worst ? e : worst;
  }
  return worst;
}

void build_channel(double d, double a, double r, double th0, int T, ChanTab* out) {
  out->coef.clear(); out->err = 0.0;
  std::vector<std::array<float, 4>> best;
  for (int side = 0; side < 2; ++side) {
    const double sgn = side ? -1.0 : 1.0;
    // guard segment: the constant F(theta0), reached only by the downward tie at the very start of the side's first cell
    out->coef.push_back({(float)lif_chain_exact_host(th0, d, a, r, th0, T), 0.0f, 0.0f, 0.0f});
    for (int e = 0; e < LT_NB; ++e) {
      const double y0 = ldexp(1.0, e + 1), cw = y0;   // cell: y = 2|x| + 2 in [2^(e+1), 2^(e+2))
      int k = 0; double kerr = 0.0;
      for (; k <= LT_KMAX; ++k) {
        const int n = 1 << k;
        const double w = cw / n;
        best.assign(n, std::array<float, 4>());
        kerr = 0.0;
        bool ok = true;
        for (int j = 0; j < n; ++j) {
          float c[4];
          const double e1 = fit_segment(sgn, y0 + j * w, w, d, a, r, th0, T, c);
          kerr = e1 > kerr ? e1 : kerr;
          if (e1 > LT_TOL && k < LT_KMAX) { ok = false; break; }
          best[j] = {c[0], c[1], c[2], c[3]};
        }
        if (ok) break;
      }
      if (k > LT_KMAX) k = LT_KMAX;
      out->k[side * LT_NB + e] = (uint16_t)k;
      out->err = kerr > out->err ? kerr : out->err;
      out->coef.insert(out->coef.end(), best.begin(), best.end());
    }
  }
}

}  // namespace

void lif_table_build(const float* np4, int C, int T, LifTableHost* out) {
  out->C = C; out->T = T; out->blocks.clear(); out->image.clear(); out->max_err = 0.0; out->max_block_bytes = 0; out->usable = true;
  // identical parameter tuples (every channel of a default-initialised layer) share one fit
  std::map<std::array<float, 4>, int> uniq;
  std::vector<int> which(C);
  std::vector<std::array<float, 4>> keys;
  for (int c = 0; c < C; ++c) {
    const std::array<float, 4> key = {np4[c], np4[C + c], np4[2 * C + c], np4[3 * C + c]};
    auto it = uniq.find(key);
    if (it == uniq.end()) { it = uniq.emplace(key, (int)keys.size()).first; keys.push_back(key); }
    which[c] = it->second;
  }
  std::vector<ChanTab> tabs(keys.size());
  const int nthr = (int)std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32u));
  auto work = [&](int t0) {
    for (size_t i = t0; i < keys.size(); i += nthr) build_channel(keys[i][0], keys[i][1], keys[i][2], keys[i][3], T, &tabs[i]);
  };
  if (keys.size() < 4 || nthr == 1) { for (int t = 0; t < nthr; ++t) work(t); }
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  const int nblk = (C + LT_CH - 1) / LT_CH;
  for (int b = 0; b < nblk; ++b) {
    LifTableBlock blk;
    blk.off_bytes = (out->image.size() + 255) / 256 * 256;
    uint32_t nseg = 0;
    std::map<int, uint32_t> first;                   // unique fit -> first segment inside this block
    for (int cl = 0; cl < LT_CH && b * LT_CH + cl < C; ++cl) {
      const int w = which[b * LT_CH + cl];
      if (first.find(w) == first.end()) {
        nseg += ((uint32_t)(cl & 7) + 8u - (nseg & 7u)) & 7u;          // stagger: first segment congruent to the channel mod 8
        first[w] = nseg; nseg += (uint32_t)tabs[w].coef.size();
      }
    }
    blk.nseg = nseg;
    blk.bytes = LT_DESC_BYTES + nseg * 16u;
    out->image.resize(blk.off_bytes + blk.bytes, 0);
    uint32_t* desc = reinterpret_cast<uint32_t*>(out->image.data() + blk.off_bytes);
    float* coef = reinterpret_cast<float*>(out->image.data() + blk.off_bytes + LT_DESC_BYTES);
    for (int cl = 0; cl < LT_CH && b * LT_CH + cl < C; ++cl) {
      const int w = which[b * LT_CH + cl];
      const ChanTab& t = tabs[w];
      out->max_err = t.err > out->max_err ? t.err : out->max_err;
      uint32_t base = first[w];
      for (int cell = 0; cell < LT_NCELL; ++cell) {
        // bits(qm) = 0x4B000000 + floor(y * S) with floor(y * S) = 2^k + segment-in-cell: fold both constants into the offset
        const int k = t.k[cell], e = cell % LT_NB;
        if (e == 0) base += 1;                                           // the side's guard segment
        const float S = (float)ldexp(1.0, k - e - 1);
        uint32_t Sb; memcpy(&Sb, &S, 4);
        desc[2 * (cell * LT_CH + cl)] = Sb;
        desc[2 * (cell * LT_CH + cl) + 1] = 16u * (base - (0x4B000000u + (1u << k)));
        base += 1u << k;
      }
      memcpy(coef + 4 * (size_t)first[w], t.coef.data(), t.coef.size() * 16);
    }
    if (blk.bytes > LT_SMEM_BUDGET) out->usable = false;
    out->max_block_bytes = blk.bytes > out->max_block_bytes ? blk.bytes : out->max_block_bytes;
    out->blocks.push_back(blk);
  }
}

// Host restatement of the device lookup (lif_table_eval_vec in lif_table.cuh) on a built image: the same fp32 operations in
// the same order, with the descriptor offsets taken relative to the block's coefficient array (the kernels add its shared-memory
// address while copying the block in).  Returns NaN when the lookup would leave the block's coefficient array.
float lif_table_eval_host(const LifTableHost& t, int c, float x) {
  const LifTableBlock& b = t.blocks[(size_t)(c / LT_CH)];
  const int cl = c % LT_CH;
  const uint8_t* img = t.image.data() + b.off_bytes;
  const float y = fminf(fmaf(fabsf(x), 2.0f, 2.0f), 511.99997f);
  uint32_t yb; memcpy(&yb, &y, 4);
  const uint32_t cell = ((yb >> 23) & 7u) + (x < 0.0f ? (uint32_t)LT_NB : 0u);
  const uint32_t* d = reinterpret_cast<const uint32_t*>(img) + 2 * ((size_t)cell * LT_CH + cl);
  float S; memcpy(&S, &d[0], 4);
  const float qm = fmaf(y, S, 8388607.5f);
  const float tt = fmaf(y, S, -(qm - 8388608.0f));
  uint32_t qb; memcpy(&qb, &qm, 4);
  const uint32_t off = qb * 16u + d[1];                         // byte offset inside the coefficient array (mod 2^32)
  if (off % 16u != 0 || off / 16u >= b.nseg) return NAN;
  const float* cf = reinterpret_cast<const float*>(img + LT_DESC_BYTES + off);
  return fmaf(fmaf(fmaf(cf[3], tt, cf[2]), tt, cf[1]), tt, cf[0]);
}

}  // namespace sapcu
