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

#include <stdlib.h>
#include <string.h>

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

/* rdk_root_loglikelihood_multi restated as the call sequence it stands for (include/rdk.h:
 * count x { update_prob_matrices(2) ; update_clvs(root_op) ; compute_root_loglikelihood }), with
 * the root CLV, the root scaler and the two P-matrices put back afterwards -- the engine leaves
 * partition state untouched */
static inline int rdk_root_loglikelihood_multi(rdo_partition_t *p, const rdo_operation_t *root_op,
                                               const unsigned int *params_indices,
                                               const unsigned int *freqs_indices,
                                               const double *branch_lengths, unsigned int count,
                                               double *out_lnl) {
  const size_t       clv_n = (size_t)p->sites * p->rate_cats * p->states;
  const size_t       pm_n = (size_t)p->rate_cats * p->states * p->states;
  const unsigned int mi[2] = {root_op->child1_matrix_index, root_op->child2_matrix_index};
  const int          sc = root_op->parent_scaler_index;
  double            *keep_clv = (double *)malloc(sizeof(double) * clv_n);
  double            *keep_pm = (double *)malloc(sizeof(double) * 2 * pm_n);
  unsigned int      *keep_sc = sc >= 0 ? (unsigned int *)malloc(sizeof(unsigned int) * p->sites) : 0;
  int                rc = RDO_SUCCESS;
  memcpy(keep_clv, p->clv[root_op->parent_clv_index], sizeof(double) * clv_n);
  memcpy(keep_pm, p->pmatrix[mi[0]], sizeof(double) * pm_n);
  memcpy(keep_pm + pm_n, p->pmatrix[mi[1]], sizeof(double) * pm_n);
  if (keep_sc) memcpy(keep_sc, p->scale_buffer[sc], sizeof(unsigned int) * p->sites);
  for (unsigned int b = 0; b < count && rc == RDO_SUCCESS; ++b) {
    rc = rdo_update_prob_matrices(p, params_indices, mi, branch_lengths + 2 * b, 2);
    if (rc != RDO_SUCCESS) break;
    rdo_update_clvs(p, root_op, 1);
    out_lnl[b] = rdo_compute_root_loglikelihood(p, root_op->parent_clv_index, sc, freqs_indices, 0);
  }
  memcpy(p->clv[root_op->parent_clv_index], keep_clv, sizeof(double) * clv_n);
  memcpy(p->pmatrix[mi[0]], keep_pm, sizeof(double) * pm_n);
  memcpy(p->pmatrix[mi[1]], keep_pm + pm_n, sizeof(double) * pm_n);
  if (keep_sc) memcpy(p->scale_buffer[sc], keep_sc, sizeof(unsigned int) * p->sites);
  free(keep_clv);
  free(keep_pm);
  free(keep_sc);
  return rc;
}

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
