// Micro-benchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) per SM sub-partition on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_rate fma_rate.cu && ./fma_rate
// Prints cycles per warp-instruction per SMSP for 1..8 resident warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: FFMA (32 FMA / warp-inst), 1: FFMA2 (64 FMA / warp-inst), 2: mix FFMA2 + FMNMX (alu pipe)
__global__ void k(float* out, long long* cyc, int iters) {
  float2 a[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  float2 w = make_float2(1.0001f, 0.9999f), x = make_float2(0.5f, 0.25f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        if (MODE == 0) {
          a[i].x = fmaf(w.x, x.x, a[i].x);
          a[i].y = fmaf(w.y, x.y, a[i].y);
        } else {
          a[i] = __ffma2_rn(w, x, a[i]);
          if (MODE == 2 && (i & 3) == 0) x.x = fmaxf(x.x, a[i].y);
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x.x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  const int iters = 2000;
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int wps = 1; wps <= 8; wps *= 2) {        // warps per SMSP
    const int threads = wps * 4 * 32;
    k<MODE><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double inst_per_warp = (double)iters * 8 * 12 * (MODE == 0 ? 2 : 1);
    const double fma_per_clk_sm = (double)iters * 8 * 12 * 64 * wps * 4 / avg;
    printf("%-6s warps/SMSP=%d  cycles/warp-inst/SMSP=%.2f  FMA/clk/SM=%.1f\n", name, wps, avg / (inst_per_warp * wps),
           fma_per_clk_sm);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA");
  run<1>("FFMA2");
  run<2>("F2+MNMX");
  return 0;
}
