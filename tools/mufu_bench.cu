// Micro-benchmark: what the MUFU pipe and the fast LIF recurrence sustain on this GPU, in registers (no memory traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<pkg>/csrc tools/mufu_bench.cu -o gpurun_out/mufu_bench
//   ./mufu_bench            -> MUFU lane-ops/s (ex2 only; ex2+ex2+rcp) and LIF element-steps/s for 4..32 warps per SM
#include <cstdio>
#include <cuda_runtime.h>
#include "neuron.cuh"

using namespace sapcu;

template <int KIND>
__global__ void bench_kernel(float* out, int iters, float seed) {
  float u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = seed + 0.01f * (threadIdx.x + 8 * i);
  if (KIND == 0) {                       // 8 independent ex2 chains
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = exp2f_approx(u[i] * -0.5f);
    }
  } else if (KIND == 1) {                // the LIF mix of MUFU ops without its FP work: ex2, ex2, rcp
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = rcp_approx(2.0f + exp2f_approx(u[i]) + exp2f_approx(-u[i]));
    }
  } else {                               // the real recurrence, T = 4 per call
    NeuronParams p{0.9f, 0.01f, 0.5f, 1.0f};
    for (int it = 0; it < iters; ++it) {
      lif_chain_vec_fast<8>(u, p, 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = u[i] * 3.0f - 0.7f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += u[i];
  if (s == 12345.678f) out[0] = s;
}

template <int KIND>
void run(const char* name, double ops_per_iter_thread, int warps_per_sm) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out; cudaMalloc(&out, 4);
  const int threads = 128, blocks = sms * warps_per_sm / 4, iters = 4000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  bench_kernel<KIND><<<blocks, threads>>>(out, iters, 0.3f);
  cudaEventRecord(a);
  bench_kernel<KIND><<<blocks, threads>>>(out, iters, 0.3f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double ops = ops_per_iter_thread * iters * (double)threads * blocks;
  printf("%-28s warps/SM %2d : %8.3f T/s  (%.2f per clk per SM at 1.9 GHz)\n", name, warps_per_sm, ops / ms / 1e9,
         ops / (ms * 1e-3) / sms / 1.9e9);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<0>("ex2 lane-ops", 8, w);
  for (int w : {4, 8, 16, 32}) run<1>("ex2+ex2+rcp lane-ops", 24, w);
  for (int w : {4, 8, 16, 32, 48}) run<2>("LIF element-steps (T=4)", 32, w);
  return 0;
}
