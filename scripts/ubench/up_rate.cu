// Micro-benchmark: issue rate of the skinny propup inner loop (LDS.128 W + broadcast LDS.128 v + BT*4 FFMA per
// row) as a function of warps per SM.  One CTA per SM, no barriers, data resident in shared memory.
#include <cstdio>
#include <cuda_runtime.h>
template <int BT, bool PIPE>
__global__ void k(float* out, int rows, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];
  float* vs = reinterpret_cast<float*>(sm);               // [32][BT padded to 12/20]
  unsigned char* tile = sm + 32 * 20 * 4;                 // [32 rows][NT*16 B]
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int BTP = (BT + 3) / 4 * 4;
  for (int i = tid; i < 32 * 20; i += nt) vs[i] = 0.001f * i;
  for (int i = tid; i < 32 * nt * 4; i += nt) reinterpret_cast<float*>(tile)[i] = 0.002f * (i & 255);
  __syncthreads();
  float4 acc[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) acc[b] = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (!PIPE) {
      for (int r = 0; r < rows; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(tile + (size_t)(r & 31) * nt * 16 + tid * 16);
        const float4* vr = reinterpret_cast<const float4*>(vs + (r & 31) * BTP);
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) {
          const float4 vv = vr[b4];
          const float xs[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int b = b4 * 4 + t;
            if (b < BT) {
              acc[b].x = fmaf(xs[t], w.x, acc[b].x); acc[b].y = fmaf(xs[t], w.y, acc[b].y);
              acc[b].z = fmaf(xs[t], w.z, acc[b].z); acc[b].w = fmaf(xs[t], w.w, acc[b].w);
            }
          }
        }
      }
    } else {
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(tile + (size_t)(r & 31) * nt * 16 + tid * 16);
        const float4* vr = reinterpret_cast<const float4*>(vs + (r & 31) * BTP);
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) {
          const float4 vv = vr[b4];
          const float xs[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int b = b4 * 4 + t;
            if (b < BT) {
              acc[b].x = fmaf(xs[t], w.x, acc[b].x); acc[b].y = fmaf(xs[t], w.y, acc[b].y);
              acc[b].z = fmaf(xs[t], w.z, acc[b].z); acc[b].w = fmaf(xs[t], w.w, acc[b].w);
            }
          }
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int b = 0; b < BT; ++b) s += acc[b].x + acc[b].y + acc[b].z + acc[b].w;
  out[blockIdx.x * nt + tid] = s;
  if (tid == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 148 * 1024)[0] = t1 - t0;
}
template <int BT, bool PIPE>
void run(float* out, int nt) {
  const int rows = 64, iters = 200;
  size_t smem = 32 * 20 * 4 + 32 * (size_t)nt * 16;
  cudaFuncSetAttribute(k<BT, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<BT, PIPE><<<148, nt, smem>>>(out, rows, iters);
  k<BT, PIPE><<<148, nt, smem>>>(out, rows, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk; cudaMemcpy(&clk, out + 148 * 1024, 8, cudaMemcpyDeviceToHost);
  double fma = (double)rows * iters * BT * 4 * nt;
  printf("BT=%d unroll=%d NT=%4d (%d warps/SMSP): %.1f FMA/clk/SM  (%.1f clk per warp-row) %s\n", BT, (int)PIPE, nt, nt / 128,
         fma / clk, (double)clk / (rows * iters), cudaGetErrorString(e));
}
int main() {
  float* out; cudaMalloc(&out, (148 * 1024 + 16) * 4);
  for (int nt : {128, 256, 384, 512}) { run<10, false>(out, nt); run<10, true>(out, nt); run<20, false>(out, nt); run<20, true>(out, nt); }
  return 0;
}
