// Runtime dispatch of the GEMM engines over the (loader, activation, row-group) combinations that the
// fn and fd forwards use.
#include "gemm_simt.cuh"

namespace sapcu {

int launch_gemm_simt(const GemmArgs& g, int amode, bool precise, cudaStream_t st) {
  SAPCU_REQUIRE(g.K % GBK == 0, "gemm: K=%d must be a multiple of %d", g.K, GBK);
  SAPCU_REQUIRE(g.R >= 0 && g.N >= 1, "gemm: bad shape R=%lld N=%d", (long long)g.R, g.N);
  SAPCU_REQUIRE((g.lda % 4) == 0 || amode == A_EDGECAT, "gemm: lda=%lld must be a multiple of 4", (long long)g.lda);
  if (g.R == 0) return 0;
#define SAPCU_G(AM, ACT, GRP)                                                      \
  return precise ? launch_gemm_simt_t<AM, ACT, GRP, true>(g, st)                   \
                 : launch_gemm_simt_t<AM, ACT, GRP, false>(g, st)
  if (amode == A_PLAIN && g.group == 0) {
    switch (g.act) {
      case ACT_NONE:  return launch_gemm_simt_t<A_PLAIN, ACT_NONE, 0, true>(g, st);
      case ACT_LEAKY: return launch_gemm_simt_t<A_PLAIN, ACT_LEAKY, 0, true>(g, st);
      case ACT_GELU:  return launch_gemm_simt_t<A_PLAIN, ACT_GELU, 0, true>(g, st);
      case ACT_LIF:   SAPCU_G(A_PLAIN, ACT_LIF, 0);
    }
  }
  if (amode == A_EDGECAT && g.group == 32 && g.act == ACT_LEAKY) {
    SAPCU_REQUIRE(g.kk == 32 && (g.C % 8) == 0 && g.K == 2 * g.C, "gemm(edgecat): needs k=32, C%%8==0, K=2C");
    return launch_gemm_simt_t<A_EDGECAT, ACT_LEAKY, 32, true>(g, st);
  }
  if (amode == A_EDGECAT && g.group == 0 && g.act == ACT_LEAKY) {      // any k: per-edge activations out, max over k by the caller
    SAPCU_REQUIRE(g.kk >= 1 && (g.C % 8) == 0 && g.K == 2 * g.C, "gemm(edgecat): needs C%%8==0, K=2C");
    return launch_gemm_simt_t<A_EDGECAT, ACT_LEAKY, 0, true>(g, st);
  }
  if (amode == A_ATTNIN && g.group == 0 && g.act == ACT_LIF) SAPCU_G(A_ATTNIN, ACT_LIF, 0);
#undef SAPCU_G
  set_error("gemm: unsupported combination amode=%d act=%d group=%d", amode, g.act, g.group);
  return -1;
}

}  // namespace sapcu
