// Instruction-throughput microbenchmarks for the NN search loop (run under gpurun):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
// Answers: (1) FP32 FFMA peak of this B200, (2) are FADD2/FMUL2/FFMA2 (packed f32x2) full- or
// half-rate per issue, (3) do FMNMX (ALU pipe) issues overlap with FMA-pipe issues.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int CH = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* sink, int iters) {
  float a[CH], m[CH];
  float2 p[CH / 2];
  const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-9f * (blockIdx.x + 1);
  const float2 x2 = make_float2(x, x * 1.0000001f), y2 = make_float2(y, y * 2.f);
#pragma unroll
  for (int i = 0; i < CH; ++i) { a[i] = (float)i + y; m[i] = 1e30f - i; }
#pragma unroll
  for (int i = 0; i < CH / 2; ++i) p[i] = make_float2((float)i + y, (float)i - y);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {            // 16 FFMA
#pragma unroll
      for (int i = 0; i < CH; ++i) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 1) {     // 8 FFMA2 (16 lanes-worth of FMAs)
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) p[i] = __ffma2_rn(p[i], x2, y2);
    } else if (MODE == 2) {     // 16 FFMA + 8 FMNMX
#pragma unroll
      for (int i = 0; i < CH; ++i) a[i] = fmaf(a[i], x, y);
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) m[i] = fminf(m[i], a[i + 8]);
    } else if (MODE == 3) {     // 8 FFMA2 + 8 FMNMX
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) m[i] = fminf(m[i], p[i].x);
    } else if (MODE == 4) {     // 8 FFMA2 + 16 FMNMX
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) { m[i] = fminf(m[i], p[i].x); m[i + 8] = fminf(m[i + 8], p[i].y); }
    } else if (MODE == 5) {     // 8 FADD2
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) p[i] = __fadd2_rn(p[i], y2);
    } else if (MODE == 6) {     // 16 FMNMX only
#pragma unroll
      for (int i = 0; i < CH; ++i) m[i] = fminf(m[i], a[i] + (float)it);
    } else if (MODE == 7) {     // 16 FFMA + 16 FMNMX
#pragma unroll
      for (int i = 0; i < CH; ++i) a[i] = fmaf(a[i], x, y);
#pragma unroll
      for (int i = 0; i < CH; ++i) m[i] = fminf(m[i], a[i]);
    } else if (MODE == 8) {     // 8 FFMA2 + 4 FMNMX3-able (min of 3)
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) p[i] = __ffma2_rn(p[i], x2, y2);
#pragma unroll
      for (int i = 0; i < CH / 4; ++i) m[i] = fminf(fminf(m[i], p[2 * i].x), p[2 * i + 1].y);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += a[i] + m[i];
#pragma unroll
  for (int i = 0; i < CH / 2; ++i) s += p[i].x + p[i].y;
  if (s == 123.456f) sink[0] = s;
}

template <int MODE>
void run(const char* name, double fma_lane_ops_per_iter, double issues_per_iter, float* sink, int sms, double mhz) {
  const int iters = 20000, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  k<MODE><<<blocks, threads>>>(sink, 1000);
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0));
    k<MODE><<<blocks, threads>>>(sink, iters);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double thr = (double)blocks * threads * iters;
  const double warps = thr / 32.0;
  const double issue_rate = warps * issues_per_iter / (best * 1e-3) / (sms * 4.0) / (mhz * 1e6);
  printf("%-28s %8.3f ms  %8.2f TFLOP/s(fma=2)  warp-instr/clk/SMSP @%.0fMHz = %.3f\n", name, best,
         thr * fma_lane_ops_per_iter * 2.0 / (best * 1e-3) / 1e12, mhz, issue_rate);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CHECK(cudaGetDevice(&dev));
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  float* sink; CHECK(cudaMalloc(&sink, 16));
  const double mhz = khz / 1000.0;
  printf("SMs=%d clockRate=%.0f MHz\n", sms, mhz);
  run<0>("16 FFMA", 16, 16, sink, sms, mhz);
  run<1>("8 FFMA2", 16, 8, sink, sms, mhz);
  run<5>("8 FADD2", 8, 8, sink, sms, mhz);
  run<6>("16 FADD+16 FMNMX", 8, 32, sink, sms, mhz);
  run<2>("16 FFMA + 8 FMNMX", 16, 24, sink, sms, mhz);
  run<7>("16 FFMA + 16 FMNMX", 16, 32, sink, sms, mhz);
  run<3>("8 FFMA2 + 8 FMNMX", 16, 16, sink, sms, mhz);
  run<4>("8 FFMA2 + 16 FMNMX", 16, 24, sink, sms, mhz);
  run<8>("8 FFMA2 + 4 FMNMX3", 16, 12, sink, sms, mhz);
  return 0;
}
