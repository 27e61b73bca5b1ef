/*
 * tests/oracle_shim/rdk.h -- TEST INFRASTRUCTURE.
 *
 * Lets the host sources (root_digger_b200/host/*.cpp, which only know the
 * rdk_* C ABI) be compiled against the CPU oracle: every rdk_* name used by the
 * host maps to the oracle's rdo_* function of the same shape.  The resulting
 * tests/_build/librd_host_oracle.so is the oracle-backed model_t the parity
 * tests compare the CUDA-backed model_t with.  Never part of the product.
 */
#ifndef RDK_H_
#define RDK_H_
#include "rd_oracle.h"

#define RDK_SUCCESS RDO_SUCCESS
#define RDK_FAILURE RDO_FAILURE
#define RDK_SCALE_BUFFER_NONE RDO_SCALE_BUFFER_NONE
#define RDK_GAMMA_RATES_MEAN RDO_GAMMA_RATES_MEAN
#define RDK_GAMMA_RATES_MEDIAN RDO_GAMMA_RATES_MEDIAN
#define RDK_ATTRIB_SITE_REPEATS RDO_ATTRIB_SITE_REPEATS
#define RDK_ATTRIB_NONREV RDO_ATTRIB_NONREV

typedef rdo_state_t     rdk_state_t;
typedef rdo_operation_t rdk_operation_t;
typedef rdo_partition_t rdk_partition_t;

#define rdk_errmsg rdo_errmsg
#define rdk_errno rdo_errno
#define rdk_map_nt rdo_map_nt
#define rdk_partition_create rdo_partition_create
#define rdk_partition_destroy rdo_partition_destroy
#define rdk_set_tip_states rdo_set_tip_states
#define rdk_set_pattern_weights rdo_set_pattern_weights
#define rdk_set_subst_params rdo_set_subst_params
#define rdk_set_frequencies rdo_set_frequencies
#define rdk_set_category_rates rdo_set_category_rates
#define rdk_set_category_weights rdo_set_category_weights
#define rdk_update_invariant_sites rdo_update_invariant_sites
#define rdk_update_invariant_sites_proportion rdo_update_invariant_sites_proportion
#define rdk_update_prob_matrices rdo_update_prob_matrices
#define rdk_update_clvs rdo_update_clvs
#define rdk_compute_root_loglikelihood rdo_compute_root_loglikelihood
#define rdk_compute_gamma_cats rdo_compute_gamma_cats
#define rdk_msa_empirical_frequencies rdo_msa_empirical_frequencies

/* rdk_sweep_root_placements restated as the call sequence it stands for */
static inline int rdk_sweep_root_placements(rdo_partition_t *p, unsigned int placements,
                                            const unsigned int *params_indices,
                                            const unsigned int *freqs_indices,
                                            const unsigned int *pm_offsets,
                                            const unsigned int *matrix_indices,
                                            const double *branch_lengths, const unsigned int *op_offsets,
                                            const rdo_operation_t *operations, unsigned int root_clv_index,
                                            int root_scaler_index, double *out_lnl) {
  for (unsigned int q = 0; q < placements; ++q) {
    if (rdo_update_prob_matrices(p, params_indices, matrix_indices + pm_offsets[q],
                                 branch_lengths + pm_offsets[q],
                                 pm_offsets[q + 1] - pm_offsets[q]) != RDO_SUCCESS)
      return RDO_FAILURE;
    rdo_update_clvs(p, operations + op_offsets[q], op_offsets[q + 1] - op_offsets[q]);
    out_lnl[q] = rdo_compute_root_loglikelihood(p, root_clv_index, root_scaler_index, freqs_indices, 0);
  }
  return RDO_SUCCESS;
}
/* the flags only concern what is stored, never the values returned */
#define RDK_SWEEP_KEEP_ROOT 1u
#define RDK_SWEEP_DISCARD 2u
static inline int rdk_sweep_root_placements_ex(rdo_partition_t *p, unsigned int placements,
                                               const unsigned int *params_indices,
                                               const unsigned int *freqs_indices,
                                               const unsigned int *pm_offsets,
                                               const unsigned int *matrix_indices,
                                               const double *branch_lengths, const unsigned int *op_offsets,
                                               const rdo_operation_t *operations, unsigned int root_clv_index,
                                               int root_scaler_index, unsigned int flags, double *out_lnl) {
  (void)flags;
  return rdk_sweep_root_placements(p, placements, params_indices, freqs_indices, pm_offsets, matrix_indices,
                                   branch_lengths, op_offsets, operations, root_clv_index, root_scaler_index,
                                   out_lnl);
}
/* chunks may always run in order; the hint is > 1 for small alignments so that the CPU
 * tests walk the chunked schedules of the host mirror too */
#define RDK_SWEEP_MAX_CHUNKS 16u
#define RDK_SHARD_ALIGN 256u
static inline unsigned int rdk_sweep_chunk_hint(unsigned int sites, unsigned int rate_cats) {
  (void)rate_cats;
  return sites < 4096u ? 3u : 1u;
}
static inline int rdk_sweep_root_placements_chunks(rdo_partition_t *p, unsigned int placements,
                                                   const unsigned int *params_indices,
                                                   const unsigned int *freqs_indices,
                                                   const unsigned int *pm_offsets,
                                                   const unsigned int *matrix_indices,
                                                   const double *branch_lengths,
                                                   const unsigned int *op_offsets,
                                                   const rdo_operation_t *operations,
                                                   unsigned int root_clv_index, int root_scaler_index,
                                                   unsigned int flags, unsigned int n_chunks,
                                                   const unsigned int *chunk_offsets, double *out_lnl) {
  (void)flags;
  (void)n_chunks;
  (void)chunk_offsets;
  return rdk_sweep_root_placements(p, placements, params_indices, freqs_indices, pm_offsets, matrix_indices,
                                   branch_lengths, op_offsets, operations, root_clv_index, root_scaler_index,
                                   out_lnl);
}
#endif
