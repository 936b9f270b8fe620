// Micro-benchmark: skinny propup inner loop with packed FFMA2 (pairs along the batch index) vs scalar FFMA.
#include <cstdio>
#include <cuda_runtime.h>
template <int BT, int UNR>
__global__ void k2(float* out, int rows, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];
  float* vs = reinterpret_cast<float*>(sm);
  unsigned char* tile = sm + 32 * 20 * 4;
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int BTP = (BT + 3) / 4 * 4;
  for (int i = tid; i < 32 * 20; i += nt) vs[i] = 0.001f * i;
  for (int i = tid; i < 32 * nt * 4; i += nt) reinterpret_cast<float*>(tile)[i] = 0.002f * (i & 255);
  __syncthreads();
  float2 acc[BTP / 2][4];
#pragma unroll
  for (int b = 0; b < BTP / 2; ++b)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[b][c] = make_float2(0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll UNR
    for (int r = 0; r < rows; ++r) {
      const float4 w = *reinterpret_cast<const float4*>(tile + (size_t)(r & 31) * nt * 16 + tid * 16);
      const float2 wd[4] = {make_float2(w.x, w.x), make_float2(w.y, w.y), make_float2(w.z, w.z), make_float2(w.w, w.w)};
      const float4* vr = reinterpret_cast<const float4*>(vs + (r & 31) * BTP);
#pragma unroll
      for (int b4 = 0; b4 < BTP / 4; ++b4) {
        const float4 vv = vr[b4];
        const float2 p0 = make_float2(vv.x, vv.y), p1 = make_float2(vv.z, vv.w);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[b4 * 2][c] = __ffma2_rn(p0, wd[c], acc[b4 * 2][c]);
          if (b4 * 4 + 2 < BT) acc[b4 * 2 + 1][c] = __ffma2_rn(p1, wd[c], acc[b4 * 2 + 1][c]);
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int b = 0; b < BTP / 2; ++b)
#pragma unroll
    for (int c = 0; c < 4; ++c) s += acc[b][c].x + acc[b][c].y;
  out[blockIdx.x * nt + tid] = s;
  if (tid == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 148 * 1024)[0] = t1 - t0;
}
template <int BT, int UNR>
void run(float* out, int nt) {
  const int rows = 64, iters = 200;
  size_t smem = 32 * 20 * 4 + 32 * (size_t)nt * 16;
  cudaFuncSetAttribute(k2<BT, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k2<BT, UNR><<<148, nt, smem>>>(out, rows, iters);
  k2<BT, UNR><<<148, nt, smem>>>(out, rows, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk; cudaMemcpy(&clk, out + 148 * 1024, 8, cudaMemcpyDeviceToHost);
  double fma = (double)rows * iters * BT * 4 * nt;
  printf("FFMA2 BT=%d unroll=%d NT=%4d (%d warps/SMSP): %.1f FMA/clk/SM  (%.1f clk per row-iteration) %s\n", BT, UNR, nt, nt / 128,
         fma / clk, (double)clk / (rows * iters), cudaGetErrorString(e));
}
int main() {
  float* out; cudaMalloc(&out, (148 * 1024 + 16) * 4);
  for (int nt : {128, 256, 384}) { run<10, 1>(out, nt); run<10, 2>(out, nt); run<20, 1>(out, nt); run<20, 2>(out, nt); run<12, 1>(out, nt);}
  return 0;
}
