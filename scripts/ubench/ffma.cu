// FP32 issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (sm_100a) vs mma.sync tf32
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a0) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 a = make_float2(a0, a0 * 0.5f), b = make_float2(0.999f, 1.001f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, b.x, a.x); acc[i].y = fmaf(acc[i].y, b.y, a.y); }
      else acc[i] = __ffma2_rn(acc[i], b, a);
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) kmma(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f000000u, 0x3f000000u}, b0 = 0x3f800000u, b1 = 0x3e800000u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000; float ms;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 256>>>(out, iters, 0.25f); else k<1><<<148 * 4, 256>>>(out, iters, 0.25f);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    double fma = 148.0 * 4 * 256 * 32.0 * iters;
    printf("mode %d (%s): %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM @1.965GHz\n", mode, mode ? "FFMA2" : "FFMA", ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
  }
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); kmma<<<148 * 4, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  double mac = 148.0 * 4 * 8 * 8.0 * 1024 * iters;
  printf("mma.sync m16n8k8 tf32: %.3f ms  %.1f TFLOP/s  %.1f MAC/clk/SM  (%.1f clk per MMA per SMSP)\n", ms, 2 * mac / ms * 1e-9, mac / (ms * 1e-3) / 148 / 1.965e9,
         (ms * 1e-3 * 1.965e9) / (iters * 8.0 * 8 * 4 / 4));
  return 0;
}
