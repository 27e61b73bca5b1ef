// Microbenchmark: fp64 FMA latency / throughput per SM sub-partition on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = __fma_rn(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// mixed: one DFMA + M independent integer ops, to see whether fp64 issue blocks the other pipes
template <int ILP, int M>
__global__ void kmix(double* out, long long* cyc, int iters, double a, double b, unsigned c) {
  double x[ILP];
  unsigned y[M > 0 ? M : 1];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < M; ++i) y[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        x[i] = __fma_rn(x[i], a, b);
#pragma unroll
        for (int j = 0; j < M; ++j) y[j] = (y[j] ^ c) + (unsigned)j * 3u + (y[j] >> 3);
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < M; ++i) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps) {
  double* out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  int iters = 2048;
  k<ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
  k<ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double n = (double)iters * 8 * ILP;  // DFMAs per thread
  printf("ILP %2d warps/SM %2d: %7.2f cycles per DFMA per warp; SMSP cycles per warp-DFMA %6.2f\n", ILP, warps,
         h / n, h / (n * warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}
template <int ILP, int M>
void runmix(int warps) {
  double* out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  int iters = 1024;
  kmix<ILP, M><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9, 0x55u);
  kmix<ILP, M><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9, 0x55u);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double n = (double)iters * 8 * ILP;
  printf("mix ILP %2d +%d int-ops/DFMA warps/SM %2d: SMSP cycles per warp-DFMA %6.2f\n", ILP, M, warps,
         h / (n * warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(4); run<2>(4); run<4>(4); run<8>(4); run<16>(4);
  run<1>(8); run<4>(8); run<8>(8); run<16>(8);
  run<1>(16); run<4>(16); run<8>(16);
  run<4>(32); run<8>(32);
  runmix<8, 0>(8); runmix<8, 1>(8); runmix<8, 2>(8); runmix<8, 3>(8);
  runmix<8, 1>(16); runmix<8, 2>(16); runmix<8, 3>(16);
  return 0;
}
