// rdk_program_inst.cu -- the instantiations of clv_program_kernel for ONE number of rate
// categories (compiled once per K with -DRDK_INST_K=K: the translation units build in
// parallel, which is what keeps the library's build time near a minute).
#define RDK_PROGRAM_KERNEL_ONLY 1
#include "rdk_kernels.cuh"

#ifndef RDK_INST_K
#error "compile with -DRDK_INST_K=<rate categories>"
#endif

namespace rdk {
namespace {

template <int K, int E, int MAXT, int MINB, bool TS>
cudaError_t launch_inst(const ProgArgs &a, int grid, int threads, cudaStream_t st) {
  // shared memory: program window + per warp: double-buffered P / tip tables of both
  // children and two mbarriers
  const int warps = threads / 32;
#if RDK_TABLES_L1
  (void)warps;
  const size_t smem = sizeof(Instr) * kProgWindow;  // the tables are read through L1
#else
  const size_t smem = sizeof(Instr) * kProgWindow + (size_t)warps * (sizeof(double) * 2 * 2 * kTabDoubles * K + 16);
#endif
  static size_t configured = 0;  // per template instantiation
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(clv_program_kernel<K, E, MAXT, MINB, TS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured = smem;
  }
  clv_program_kernel<K, E, MAXT, MINB, TS><<<dim3((unsigned)grid, a.n_chunks > 1 ? a.n_chunks : 1u), threads, smem, st>>>(a);
  return cudaSuccess;
}

}  // namespace

// elements per thread E trades registers (occupancy) for fewer per-instruction preambles
// per element; ts: the tail-skip instantiation (choose_tail_skip in rdk_abi.cu)
template <>
cudaError_t launch_program<RDK_INST_K>(const ProgArgs &a, int grid, int threads, int E, bool ts, cudaStream_t st) {
  constexpr int K = RDK_INST_K;
  if (E != 1) threads = threads < 128 ? threads : 128;
  switch (E) {
    case 1: return launch_inst<K, 1, 256, 3, false>(a, grid, threads, st);
    case 4:
      return ts ? launch_inst<K, 4, 128, 2, true>(a, grid, threads, st)
                : launch_inst<K, 4, 128, 2, false>(a, grid, threads, st);
    default:
      return ts ? launch_inst<K, 2, 128, RDK_MINB2, true>(a, grid, threads, st)
                : launch_inst<K, 2, 128, RDK_MINB2, false>(a, grid, threads, st);
  }
}

}  // namespace rdk
