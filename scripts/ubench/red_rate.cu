// Micro-benchmark: cost of a grid-wide vector sum through L2 reductions (red.global.add), 148 CTAs x 256 threads,
// every CTA adds its own N-element partial into the same N-element accumulator.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(void* accv, int n, int stagger) {
  const int tid = threadIdx.x;
  const int rot = stagger ? (blockIdx.x * (n / gridDim.x)) : 0;
  if (MODE == 0) {            // u64
    unsigned long long* a = (unsigned long long*)accv;
    for (int i = tid; i < n; i += blockDim.x) { int j = i + rot; if (j >= n) j -= n;
      asm volatile("red.global.add.u64 [%0], %1;" ::"l"(a + j), "l"((unsigned long long)(i + 1)) : "memory"); }
  } else if (MODE == 1) {     // f32
    float* a = (float*)accv;
    for (int i = tid; i < n; i += blockDim.x) { int j = i + rot; if (j >= n) j -= n;
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a + j), "f"(1.0f) : "memory"); }
  } else {                    // v4.f32
    float* a = (float*)accv;
    for (int i = tid; i < n / 4; i += blockDim.x) { int j = i + rot / 4; if (j >= n / 4) j -= n / 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a + 4 * j), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory"); }
  }
}
int main() {
  void* acc; cudaMalloc(&acc, 1 << 20); cudaMemset(acc, 0, 1 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int n : {4000, 8000, 16000})
    for (int stag = 0; stag < 2; ++stag)
      for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
          cudaEventRecord(e0);
          if (mode == 0) k<0><<<148, 256>>>(acc, n, stag); else if (mode == 1) k<1><<<148, 256>>>(acc, n, stag); else k<2><<<148, 256>>>(acc, n, stag);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("n=%5d stagger=%d %s: %.2f us (whole kernel incl. ~2us launch)\n", n, stag, mode == 0 ? "u64   " : mode == 1 ? "f32   " : "v4.f32", best * 1e3);
      }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
