#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2b(float& d0, float& d1, float a0, float a1, float b) {
  unsigned long long d, a, bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
template <int MODE> __global__ void k(float* out, long long* cyc, int iters, float a, float b, float lo, float hi) {
  float x = threadIdx.x * 1e-3f, y = x + 1.0f, z = 0.5f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) { x = fmaf(x, a, b); }                                              // FFMA chain
      if (MODE == 1) { x = fmaxf(x, lo); x = fminf(x, hi); x = x * a; }                    // FMNMX FMNMX FMUL
      if (MODE == 2) { float d = fmaf(-x, a, b); d = fmaxf(d, lo); d = fminf(d, hi); x = fmaf(z, d, x); }          // the row chain, scalar
      if (MODE == 3) { float d = fmaf(-x, a, b); d = fmaxf(d, lo); d = fminf(d, hi); fma2b(x, y, z, a, d); }      // the row chain, FFMA2 update
      if (MODE == 4) { fma2b(x, y, z, a, x); }                                           // FFMA2 chain (d depends on x)
      if (MODE == 5) { float d = fmaf(-x, a, b); x = fmaf(z, d, x); }                      // FFMA FFMA
      if (MODE == 6) { x = fmaxf(x, lo); }                                               // FMNMX chain
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}
template <int MODE> void run(const char* name, int ops) {
  float* out; long long* cyc; cudaMalloc(&out, 4 * 32 * 148); cudaMalloc(&cyc, 8 * 148);
  k<MODE><<<148, 32>>>(out, cyc, 1000, 0.999f, 0.001f, -1e30f, 1e30f); cudaDeviceSynchronize();
  k<MODE><<<148, 32>>>(out, cyc, 10000, 0.999f, 0.001f, -1e30f, 1e30f); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
  printf("%-44s %6.2f cycles per group (%d dependent ops) -> %5.2f per op\n", name, h[0] / (10000.0 * 16), ops, h[0] / (10000.0 * 16 * ops));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA -> FFMA", 1);
  run<6>("FMNMX -> FMNMX", 1);
  run<1>("FMNMX -> FMNMX -> FMUL", 3);
  run<5>("FFMA -> FFMA (row without clamp)", 2);
  run<2>("FFMA -> FMNMX -> FMNMX -> FFMA (row, scalar)", 4);
  run<3>("FFMA -> FMNMX -> FMNMX -> FFMA2 (row, f32x2)", 4);
  run<4>("FFMA2 -> FFMA2", 1);
  return 0;
}
