// Pipe-rate micro-benchmarks behind the roofline denominators (FP32 FMA pipe, packed FFMA2, MUFU.EX2) and
// the mixed-issue experiments that guided the pairwise kernel design. Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench profiles/ubench/ubench.cu && gpurun_out/ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#define CHAINS 8
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 3-register FFMA: a = a * b + c with b, c in registers (loaded at run time)
__global__ void __launch_bounds__(256) k_ffma3(long iters, const float* p, float* sink) {
  float a[CHAINS]; float b = p[0], c = p[1];
  for (int i = 0; i < CHAINS; ++i) a[i] = p[2 + i] + threadIdx.x;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;
}
// packed FFMA2
__global__ void __launch_bounds__(256) k_ffma2(long iters, const float* p, float* sink) {
  float2 a[CHAINS]; float2 b = make_float2(p[0], p[0]), c = make_float2(p[1], p[1]);
  for (int i = 0; i < CHAINS; ++i) a[i] = make_float2(p[2 + i] + threadIdx.x, p[3 + i]);
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = __ffma2_rn(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y;
  if (s == 12345.678f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_ex2(long iters, const float* p, float* sink) {
  float a[CHAINS];
  for (int i = 0; i < CHAINS; ++i) a[i] = p[2 + i] * 0.01f + 0.5f;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = ex2f(-a[i]);
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;
}
// mixed: per iteration NF scalar FFMA + 1 MUFU per chain-group (ratio NF*CHAINS : CHAINS)
template <int NF>
__global__ void __launch_bounds__(256) k_mix(long iters, const float* p, float* sink) {
  float a[CHAINS], e[CHAINS]; float b = p[0], c = p[1];
  for (int i = 0; i < CHAINS; ++i) { a[i] = p[2 + i] + threadIdx.x; e[i] = 0.5f; }
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < NF; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = fmaf(a[i], b, c);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) e[i] = ex2f(-e[i]);
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i] + e[i];
  if (s == 12345.678f) sink[0] = s;
}
// mixed packed: NF FFMA2 + 2 MUFU per chain
template <int NF>
__global__ void __launch_bounds__(256) k_mix2(long iters, const float* p, float* sink) {
  float2 a[CHAINS]; float e[CHAINS], f[CHAINS]; float2 b = make_float2(p[0], p[0]), c = make_float2(p[1], p[1]);
  for (int i = 0; i < CHAINS; ++i) { a[i] = make_float2(p[2 + i] + threadIdx.x, p[3 + i]); e[i] = 0.5f; f[i] = 0.6f; }
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < NF; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = __ffma2_rn(a[i], b, c);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { e[i] = ex2f(-e[i]); f[i] = ex2f(-f[i]); }
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y + e[i] + f[i];
  if (s == 12345.678f) sink[0] = s;
}

// packed half / bfloat16 FMA (HFMA2) and the int8 dot product (IDP.4A): candidates for a low-precision pairwise screen
__global__ void __launch_bounds__(256) k_hfma2(long iters, const float* p, float* sink) {
  __half2 a[CHAINS]; __half2 b = __floats2half2_rn(p[0], p[0]), c = __floats2half2_rn(p[1], p[1]);
  for (int i = 0; i < CHAINS; ++i) a[i] = __floats2half2_rn(p[2 + i] + threadIdx.x * 1e-3f, p[3 + i]);
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = __hfma2(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += __low2float(a[i]) + __high2float(a[i]);
  if (s == 12345.678f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_bfma2(long iters, const float* p, float* sink) {
  __nv_bfloat162 a[CHAINS]; __nv_bfloat162 b = __floats2bfloat162_rn(p[0], p[0]), c = __floats2bfloat162_rn(p[1], p[1]);
  for (int i = 0; i < CHAINS; ++i) a[i] = __floats2bfloat162_rn(p[2 + i] + threadIdx.x * 1e-3f, p[3 + i]);
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = __hfma2(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += __low2float(a[i]) + __high2float(a[i]);
  if (s == 12345.678f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_dp4a(long iters, const float* p, float* sink) {
  int a[CHAINS]; int b = __float_as_int(p[0]), c = __float_as_int(p[1]) | 0x01010101;
  for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x + i;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = __dp4a(b ^ a[i], c, a[i]);
  }
  int s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
  if (s == 123456789) sink[0] = s;
}
// FMNMX (ALU pipe) alone and interleaved 1:4 with FFMA2: does the min/max of a screen ride along for free?
__global__ void __launch_bounds__(256) k_fmnmx(long iters, const float* p, float* sink) {
  float a[CHAINS]; float b = p[0];
  for (int i = 0; i < CHAINS; ++i) a[i] = p[2 + i] + threadIdx.x;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) { a[i] = fmaxf(a[i], b); b = -b; }
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;
}
template <int NF>
__global__ void __launch_bounds__(256) k_mix_mnmx(long iters, const float* p, float* sink) {
  float2 a[CHAINS]; float e[CHAINS]; float2 b = make_float2(p[0], p[0]), c = make_float2(p[1], p[1]);
  for (int i = 0; i < CHAINS; ++i) { a[i] = make_float2(p[2 + i] + threadIdx.x, p[3 + i]); e[i] = 0.5f + i; }
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < NF; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = __ffma2_rn(a[i], b, c);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { e[i] = fminf(e[i], a[i].x); e[i] = fminf(e[i], a[i].y); }
  }
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y + e[i];
  if (s == 12345.678f) sink[0] = s;
}

template <typename K>
double run(K kern, long iters, const float* p, float* sink, int blocks) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, 256>>>(iters / 8, p, sink); cudaDeviceSynchronize();
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); kern<<<blocks, 256>>>(iters, p, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best * 1e-3;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float h[16]; for (int i = 0; i < 16; ++i) h[i] = 0.001f * i; h[0] = 0.999999f; h[1] = 1e-7f;
  float *p, *sink; cudaMalloc(&p, sizeof(h)); cudaMalloc(&sink, 16); cudaMemcpy(p, h, sizeof(h), cudaMemcpyHostToDevice);
  const int blocks = sms * 8; const long iters = 4096; const double thr = (double)blocks * 256;
  double t;
  t = run(k_ffma3, iters, p, sink, blocks);
  printf("FFMA 3-reg      : %.2f TFLOP/s (%.1f G warp-inst/s/SMSP-equivalent lanes %.1f/clk/SM @1.965GHz)\n", thr * iters * 64 * 2 / t / 1e12, 0.0, thr * iters * 64 / t / sms / 1.965e9);
  t = run(k_ffma2, iters, p, sink, blocks);
  printf("FFMA2 packed    : %.2f TFLOP/s (lanes %.1f/clk/SM)\n", thr * iters * 64 * 4 / t / 1e12, thr * iters * 128 / t / sms / 1.965e9);
  t = run(k_ex2, iters, p, sink, blocks);
  printf("MUFU.EX2        : %.2f Tops/s (%.2f /clk/SM)\n", thr * iters * 64 / t / 1e12, thr * iters * 64 / t / sms / 1.965e9);
  t = run(k_mix<16>, iters, p, sink, blocks);
  printf("mix 16 FFMA:1 EX2      : FFMA %.2f TFLOP/s, EX2 %.2f Tops/s, issue %.1f%% of 4/clk/SM\n", thr * iters * 128 * 2 / t / 1e12, thr * iters * 8 / t / 1e12, 100 * thr / 32 * iters * 136 / t / sms / 4 / 1.965e9);
  t = run(k_mix<8>, iters, p, sink, blocks);
  printf("mix  8 FFMA:1 EX2      : FFMA %.2f TFLOP/s, EX2 %.2f Tops/s, issue %.1f%%\n", thr * iters * 64 * 2 / t / 1e12, thr * iters * 8 / t / 1e12, 100 * thr / 32 * iters * 72 / t / sms / 4 / 1.965e9);
  t = run(k_mix2<8>, iters, p, sink, blocks);
  printf("mix  8 FFMA2:2 EX2     : FMA %.2f TFLOP/s, EX2 %.2f Tops/s, issue %.1f%%\n", thr * iters * 64 * 4 / t / 1e12, thr * iters * 16 / t / 1e12, 100 * thr / 32 * iters * 80 / t / sms / 4 / 1.965e9);
  t = run(k_mix2<4>, iters, p, sink, blocks);
  printf("mix  4 FFMA2:2 EX2     : FMA %.2f TFLOP/s, EX2 %.2f Tops/s, issue %.1f%%\n", thr * iters * 32 * 4 / t / 1e12, thr * iters * 16 / t / 1e12, 100 * thr / 32 * iters * 48 / t / sms / 4 / 1.965e9);
  t = run(k_hfma2, iters, p, sink, blocks);
  printf("HFMA2 f16x2     : %.2f T half-FMA/s (%.1f half-FMA/clk/SM, %.1f warp-inst lanes/clk/SM)\n", thr * iters * 64 * 2 / t / 1e12, thr * iters * 128 / t / sms / 1.965e9, thr * iters * 64 / t / sms / 1.965e9);
  t = run(k_bfma2, iters, p, sink, blocks);
  printf("HFMA2 bf16x2    : %.2f T bf16-FMA/s (%.1f /clk/SM)\n", thr * iters * 64 * 2 / t / 1e12, thr * iters * 128 / t / sms / 1.965e9);
  t = run(k_dp4a, iters, p, sink, blocks);
  printf("IDP.4A          : %.2f T int8-MAC/s (%.1f MAC/clk/SM, %.1f lanes/clk/SM)\n", thr * iters * 64 * 4 / t / 1e12, thr * iters * 256 / t / sms / 1.965e9, thr * iters * 64 / t / sms / 1.965e9);
  t = run(k_fmnmx, iters, p, sink, blocks);
  printf("FMNMX           : %.2f Tops/s (%.1f lanes/clk/SM)\n", thr * iters * 64 / t / 1e12, thr * iters * 64 / t / sms / 1.965e9);
  t = run(k_mix_mnmx<8>, iters, p, sink, blocks);
  printf("mix 8 FFMA2:2 FMNMX    : FMA %.2f TFLOP/s, FMNMX %.2f Tops/s\n", thr * iters * 64 * 4 / t / 1e12, thr * iters * 16 / t / 1e12);
  t = run(k_mix_mnmx<4>, iters, p, sink, blocks);
  printf("mix 4 FFMA2:2 FMNMX    : FMA %.2f TFLOP/s, FMNMX %.2f Tops/s\n", thr * iters * 32 * 4 / t / 1e12, thr * iters * 16 / t / 1e12);
  return 0;
}
