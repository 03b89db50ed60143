// peak_fp.cu -- register-resident arithmetic micro-kernels that measure what this B200 actually sustains for
// the instruction mixes the rate kernels issue (SURVEY.md 8d asks for measured FP32 / FP64 denominators;
// MEASURED_PEAKS.json only holds HBM and bf16 tensor numbers).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peak_fp tools/peak_fp.cu && ./peak_fp > profiles/fp_peaks.json
//
// Every kernel keeps 8 independent dependency chains per thread, 8 warps x 4 CTAs per SM, and is timed with
// CUDA events over a launch long enough (~10 ms) to be at steady clocks. Prints one JSON object.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int kChains = 8, kUnroll = 16;

// mode 0: FFMA   1: FMUL+FADD (un-fused, alternating)   2: FMUL2+FADD2   3: FFMA2   4: DFMA   5: DMUL+DADD
// 6: FMUL+FADD with one IADD3 per FP instruction   7: FMUL2+FADD2 with one IADD3 per packed instruction
template <int MODE> __global__ void __launch_bounds__(256) k(float *out, int iters, float seed)
{
  float a[kChains], b = seed, c = 1.0f - seed * 1e-7f;
  double da[kChains], db = seed, dc = 1.0 - seed * 1e-9;
  u64 pa[kChains], pb, pc;
  int ia[kChains];
  {
    float2 t = make_float2(b, b); pb = *reinterpret_cast<u64 *>(&t);
    t = make_float2(c, c); pc = *reinterpret_cast<u64 *>(&t);
  }
  for (int j = 0; j < kChains; ++j) {
    a[j] = threadIdx.x * 1e-3f + j; da[j] = a[j]; ia[j] = threadIdx.x + j;
    float2 t = make_float2(a[j], -a[j]); pa[j] = *reinterpret_cast<u64 *>(&t);
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int j = 0; j < kChains; ++j) {
        if (MODE == 0) a[j] = __fmaf_rn(a[j], c, b);
        if (MODE == 1 || MODE == 6) a[j] = (u & 1) ? __fmul_rn(a[j], c) : __fadd_rn(a[j], b);
        if (MODE == 2 || MODE == 7) pa[j] = (u & 1) ? mul2(pa[j], pc) : add2(pa[j], pb);
        if (MODE == 3) pa[j] = fma2(pa[j], pc, pb);
        if (MODE == 4) da[j] = __fma_rn(da[j], dc, db);
        if (MODE == 5) da[j] = (u & 1) ? __dmul_rn(da[j], dc) : __dadd_rn(da[j], db);
        if (MODE == 6 || MODE == 7) asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[j]) : "r"(i));
      }
    }
  }
  float s = 0;
  for (int j = 0; j < kChains; ++j) {
    float2 t = *reinterpret_cast<float2 *>(&pa[j]);
    s += a[j] + (float)da[j] + t.x + t.y + (float)ia[j];
  }
  if (s == 123.456f) out[0] = s;   // keep the chains alive
}

template <int MODE> static double run(float *d, int sms, double *inst_rate)
{
  const int grid = sms * 4, iters = MODE == 4 || MODE == 5 ? 4000 : 8000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(d, iters / 8, 0.5f);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fp_inst = (double)grid * 256 * iters * kUnroll * kChains;     // thread-level FP instructions
  const double flop_per_inst = MODE == 0 || MODE == 4 ? 2 : MODE == 3 ? 4 : MODE == 2 || MODE == 7 ? 2 : 1;
  *inst_rate = fp_inst / 32 / (best * 1e-3);                                 // warp instructions per second
  return fp_inst * flop_per_inst / (best * 1e-3) / 1e12;
}

int main()
{
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no device\n"); return 1; }
  float *d; cudaMalloc(&d, 4096);
  const int sms = prop.multiProcessorCount;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double r[8], ir[8];
  r[0] = run<0>(d, sms, &ir[0]); r[1] = run<1>(d, sms, &ir[1]); r[2] = run<2>(d, sms, &ir[2]); r[3] = run<3>(d, sms, &ir[3]);
  r[4] = run<4>(d, sms, &ir[4]); r[5] = run<5>(d, sms, &ir[5]); r[6] = run<6>(d, sms, &ir[6]); r[7] = run<7>(d, sms, &ir[7]);
  const double per_sm_clk = 1.0 / sms / (clk * 1e3);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_attr\": %d,\n", prop.name, sms, clk / 1000);
  printf(" \"fp32_ffma_tflops\": %.2f, \"fp32_fmul_fadd_tflops\": %.2f, \"fp32_fmul2_fadd2_tflops\": %.2f, \"fp32_ffma2_tflops\": %.2f,\n",
         r[0], r[1], r[2], r[3]);
  printf(" \"fp64_dfma_tflops\": %.3f, \"fp64_dmul_dadd_tflops\": %.3f,\n", r[4], r[5]);
  printf(" \"fp32_fmul_fadd_with_iadd_tflops\": %.2f, \"fp32_fmul2_fadd2_with_iadd_tflops\": %.2f,\n", r[6], r[7]);
  printf(" \"warp_inst_per_sm_clk\": {\"ffma\": %.2f, \"fmul_fadd\": %.2f, \"fmul2_fadd2\": %.2f, \"ffma2\": %.2f, \"dfma\": %.3f, "
         "\"dmul_dadd\": %.3f, \"fmul_fadd+iadd (fp only)\": %.2f, \"fmul2_fadd2+iadd (fp only)\": %.2f},\n",
         ir[0] * per_sm_clk, ir[1] * per_sm_clk, ir[2] * per_sm_clk, ir[3] * per_sm_clk, ir[4] * per_sm_clk, ir[5] * per_sm_clk,
         ir[6] * per_sm_clk, ir[7] * per_sm_clk);
  printf(" \"how\": \"tools/peak_fp.cu: 8 independent chains per thread, 4 CTAs x 256 threads per SM, best of 5 launches, CUDA events\"}\n");
  return 0;
}
