// Microbenchmark: shared-memory broadcast read throughput for the P-table access patterns.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int WIDTH>
__global__ void k(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double tab[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = i * 0.5;
  __syncthreads();
  unsigned lane = threadIdx.x & 31;
  unsigned sel = MODE == 0 ? 0 : MODE == 1 ? (lane & 3) : MODE == 2 ? (lane >> 3) : lane;
  const char* base = reinterpret_cast<const char*>(tab) + sel * 144;
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (WIDTH == 16) {
        double2 v; unsigned ad = (unsigned)__cvta_generic_to_shared(base + ((it & 7) * 1152 + j * 16));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ad));
        acc0 += v.x; acc1 += v.y;
      } else if (WIDTH == 8) {
        double v, w; unsigned ad = (unsigned)__cvta_generic_to_shared(base + ((it & 7) * 1152 + j * 16));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(ad));
        asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(w) : "r"(ad));
        acc0 += v; acc1 += w;
      } else {
        float4 v; unsigned ad = (unsigned)__cvta_generic_to_shared(base + ((it & 7) * 1152 + j * 16));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ad));
        acc2 += v.x; acc3 += v.w;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}
template <int MODE, int WIDTH>
void run(const char* name) {
  double* out; long long* cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  int iters = 4096, warps = 16;
  k<MODE, WIDTH><<<148, warps * 32>>>(out, cyc, iters);
  k<MODE, WIDTH><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // 16 bytes per lane per inner step
  double steps = (double)iters * 8 * warps;
  printf("%-40s %8.2f SM-cycles per warp-level 16B-per-lane read\n", name, h / steps);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 16>("LDS.128 all lanes same address");
  run<1, 16>("LDS.128 k=lane%4 (4 addresses)");
  run<2, 16>("LDS.128 k=lane/8 (1 per quarter)");
  run<3, 16>("LDS.128 32 distinct (stride 144B)");
  run<0, 8>("2xLDS.64 all lanes same address");
  run<1, 8>("2xLDS.64 k=lane%4");
  run<2, 8>("2xLDS.64 k=lane/8");
  run<0, 4>("LDS.128(f4) same address");
  return 0;
}
