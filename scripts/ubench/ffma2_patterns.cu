// FFMA2 issue-rate ceilings for the operand patterns of the skinny kernel (registers only, no memory), as a function
// of warps per scheduler.  One CTA per SM (dynamic shared memory forces it).  40 FFMA2 per iteration.
//   A: acc[c]   = ffma2(a[i] (pair, same for 4 instr), B[i][c] (pair), acc[c])       gradient, pairs over the batch
//   B: acc[cp]  = ffma2(B[b][cp] (pair), s[b] scalar,   acc[cp])                     gradient, pairs over columns
//   C: acc[i][c]= ffma2(a[i] (pair, same for 4 instr), w[c] scalar, acc[i][c])       propup / positive phase
//   D: as A with 8 accumulators alternating between two a's (worst case: no operand reuse)
//   E: g[c]     = fmaf(s[b] (scalar, same for 4 instr), F[b][c] (scalar), g[c])      scalar FFMA gradient: 80 per iteration
//   F: as E, but two rows in flight (16 accumulators)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef unsigned long long P2;
__device__ __forceinline__ P2 pack2(float lo, float hi) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ P2 ffma2(P2 a, P2 b, P2 c) { P2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float f2(P2 v) { return __uint_as_float((unsigned)v) + __uint_as_float((unsigned)(v >> 32)); }

template <int PAT>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float seed) {
  P2 B[10][4], a[10], acc[8], acc2[5][4];
  float s[20], w[4], Fm[20][4], g[16];
  for (int b = 0; b < 20; ++b) for (int c = 0; c < 4; ++c) Fm[b][c] = 1.0f + seed * (b + c + threadIdx.x);
  for (int c = 0; c < 16; ++c) g[c] = 0.f;
  for (int i = 0; i < 10; ++i) { a[i] = pack2(seed + i, seed * i); for (int c = 0; c < 4; ++c) B[i][c] = pack2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f - 1e-3f * c * seed); }
  for (int i = 0; i < 20; ++i) s[i] = seed * (i + 1);
  for (int c = 0; c < 4; ++c) w[c] = 1.0f + seed * c;
  for (int c = 0; c < 8; ++c) acc[c] = 0ULL;
  for (int i = 0; i < 5; ++i) for (int c = 0; c < 4; ++c) acc2[i][c] = 0ULL;
  for (int it = 0; it < iters; ++it) {
    if (PAT == 0) {          // A
#pragma unroll
      for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] = ffma2(a[i], B[i][c], acc[c]);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[4 + c] = ffma2(a[5 + i], B[5 + i][c], acc[4 + c]);
      }
    } else if (PAT == 1) {   // B: 20 batch rows x 2 column pairs
#pragma unroll
      for (int b = 0; b < 20; ++b) {
        acc[0] = ffma2(B[b >> 1][(b & 1) * 2], pack2(s[b], s[b]), acc[0]);
        acc[1] = ffma2(B[b >> 1][(b & 1) * 2 + 1], pack2(s[b], s[b]), acc[1]);
      }
    } else if (PAT == 2) {   // C
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc2[i][c] = ffma2(a[i], pack2(w[c], w[c]), acc2[i][c]);
    } else if (PAT == 4) {   // E: 10 + 10 batch rows x 4 columns, scalar FMAs (80 per iteration)
#pragma unroll
      for (int b = 0; b < 20; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) g[(b >= 10 ? 4 : 0) + c] = fmaf(s[b], Fm[b][c], g[(b >= 10 ? 4 : 0) + c]);
    } else if (PAT == 5) {   // F
#pragma unroll
      for (int b = 0; b < 20; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          g[(b >= 10 ? 4 : 0) + c] = fmaf(s[b], Fm[b][c], g[(b >= 10 ? 4 : 0) + c]);
          g[8 + (b >= 10 ? 4 : 0) + c] = fmaf(w[b & 3], Fm[b][c], g[8 + (b >= 10 ? 4 : 0) + c]);
        }
    } else {                 // D
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] = ffma2(a[i], B[i][c], acc[c]); acc[4 + c] = ffma2(a[5 + i], B[5 + i][(c + 1) & 3], acc[4 + c]); }
    }
  }
  float t = 0;
  for (int c = 0; c < 16; ++c) t += g[c];
  for (int c = 0; c < 8; ++c) t += f2(acc[c]);
  for (int i = 0; i < 5; ++i) for (int c = 0; c < 4; ++c) t += f2(acc2[i][c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int PAT>
void run(float* out, int nw) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(k<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 20000; float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k<PAT><<<148, nw * 32, 200 * 1024>>>(out, iters, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  const double cyc = ms * 1e-3 * 1.965e9;
  const double n_inst = PAT == 4 ? 80.0 : PAT == 5 ? 160.0 : 40.0, fma_per = PAT >= 4 ? 1.0 : 2.0;
  printf("pattern %c  warps/SM %2d: %.3f ms  -> %.2f clk per warp-instruction per scheduler, %.2f clk per FMA per lane\n", "ABCDEF"[PAT], nw, ms,
         cyc / (n_inst * iters * (nw / 4.0)), cyc / (n_inst * fma_per * iters * (nw / 4.0)));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  for (int nw : {4, 8, 12}) { run<0>(out, nw); run<2>(out, nw); run<4>(out, nw); run<5>(out, nw); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
