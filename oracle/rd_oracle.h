/*
 * rd_oracle.h -- CPU ORACLE for the RootDigger likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under root_digger_b200/ (the product) may
 * include, link or dlopen this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * PARITY STATUS: "parity unpinned".  The arithmetic of this path lives in
 * coraxlib (https://codeberg.org/Exelixis-Lab/coraxlib.git @
 * 668ebca236472d79cb10ef411807844f8a1cd804, built with CORAX_NONREV=ON), an
 * un-vendored submodule that is absent from /root/reference (lib/coraxlib is an
 * empty directory) and the reference tests pin no numeric log-likelihood
 * (test/src/model.cpp checks invariants only).  This file restates the published
 * libpll-2/coraxlib algorithm and is anchored on
 *   - the reference call sites: src/model.cpp:159-168 (partition shape),
 *     :185,:205,:209,:244,:310-313,:324,:337 (setters), :367,:432,:842
 *     (update_prob_matrices), :402,:440,:461,:851 (update_clvs), :406,:441,:466
 *     (root log-likelihood), :239-272 (gamma cats), :329 (empirical freqs);
 *   - the reference invariants test/src/model.cpp:59-75, :271-288, :367-387;
 *   - scipy.linalg.expm / mpmath (expm), scipy.stats.gamma (gamma categories).
 *
 *
 * ARITHMETIC SPEC v2 -- the checker followed the implementation, and says so: the 4x4 mat-vec of
 * the CLV update and the dot products of the root log-likelihood are explicit FMA chains here
 * (rd_oracle.c "CLV update") because that is what the CUDA kernels issue; v1 rounded every
 * product and sum separately.  Neither order is coraxlib's (its AVX2 kernels use their own), and
 * the choice is only legitimate because the result stays bounded by independent arithmetic:
 * tests/test_oracle.py::test_oracle_against_exact_arithmetic re-evaluates the whole chain
 * (Q -> expm -> pruning -> weighted root logL) in 60-digit mpmath and requires 1e-12 relative in
 * both arithmetic modes -- three orders inside the 1e-9 parity budget of the north star.
 * The H1 conventions of the Q builder (pi-multiplication, slot order, normalisation) are
 * run-time switches (rdo_set_q_convention): pinning on a coraxlib number, if one ever becomes
 * available, is a matter of selecting the variant that reproduces it.
 *
 * Every entry point has the same shape as the corax_* function RootDigger
 * calls, with prefix rdo_.
 */
#ifndef RD_ORACLE_H_
#define RD_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RDO_SUCCESS 1
#define RDO_FAILURE 0

#define RDO_SCALE_BUFFER_NONE (-1)

#define RDO_GAMMA_RATES_MEAN 0
#define RDO_GAMMA_RATES_MEDIAN 1

/* attribute bits: accepted for call-site compatibility (src/model.cpp:145-157) */
#define RDO_ATTRIB_ARCH_CPU 0u
#define RDO_ATTRIB_ARCH_SSE (1u << 0)
#define RDO_ATTRIB_ARCH_AVX (1u << 1)
#define RDO_ATTRIB_ARCH_AVX2 (1u << 2)
#define RDO_ATTRIB_SITE_REPEATS (1u << 10)
#define RDO_ATTRIB_NONREV (1u << 11)

/* arithmetic mode of rdo_compute_root_loglikelihood:
 *   REFERENCE: libm log(), per-site terms summed serially in site order
 *              (libpll core_root_loglikelihood)
 *   ENGINE   : the spec'd software log (DESIGN.md "rd_log") and the canonical
 *              balanced pairwise tree over the zero-padded site index space;
 *              bit-for-bit what the CUDA engine must produce.
 * The two agree to ~1e-15 relative; tests state both bars. */
#define RDO_MODE_REFERENCE 0
#define RDO_MODE_ENGINE 1

typedef unsigned long long rdo_state_t;

/* field order/meaning = libpll/corax operation (fields used at
 * src/tree.cpp:401-410) */
typedef struct rdo_operation {
  unsigned int parent_clv_index;
  int          parent_scaler_index;
  unsigned int child1_clv_index;
  unsigned int child1_matrix_index;
  int          child1_scaler_index;
  unsigned int child2_clv_index;
  unsigned int child2_matrix_index;
  int          child2_scaler_index;
} rdo_operation_t;

typedef struct rdo_partition {
  unsigned int tips;
  unsigned int clv_buffers;
  unsigned int states;
  unsigned int sites;
  unsigned int rate_matrices;
  unsigned int prob_matrices;
  unsigned int rate_cats;
  unsigned int scale_buffers;
  unsigned int attributes;

  double       **clv;            /* [tips + clv_buffers][sites*rate_cats*states] */
  double       **pmatrix;        /* [prob_matrices][rate_cats*states*states]     */
  unsigned int **scale_buffer;   /* [scale_buffers][sites]                        */
  double       **subst_params;   /* [rate_matrices][states*states-states]         */
  double       **frequencies;    /* [rate_matrices][states]                       */
  double        *rates;          /* [rate_cats] */
  double        *rate_weights;   /* [rate_cats] */
  double        *prop_invar;     /* [rate_matrices] */
  unsigned int  *pattern_weights;/* [sites] */
  int           *invariant;      /* [sites] or NULL */
} rdo_partition_t;

extern int  rdo_errno;
extern char rdo_errmsg[200];
extern const rdo_state_t rdo_map_nt[256];

rdo_partition_t *rdo_partition_create(unsigned int tips,
                                      unsigned int clv_buffers,
                                      unsigned int states,
                                      unsigned int sites,
                                      unsigned int rate_matrices,
                                      unsigned int prob_matrices,
                                      unsigned int rate_cats,
                                      unsigned int scale_buffers,
                                      unsigned int attributes);
void rdo_partition_destroy(rdo_partition_t *p);

int  rdo_set_tip_states(rdo_partition_t *p, unsigned int tip_index,
                        const rdo_state_t *map, const char *sequence);
void rdo_set_pattern_weights(rdo_partition_t *p, const unsigned int *w);
void rdo_set_subst_params(rdo_partition_t *p, unsigned int params_index,
                          const double *params);
void rdo_set_frequencies(rdo_partition_t *p, unsigned int params_index,
                         const double *freqs);
void rdo_set_category_rates(rdo_partition_t *p, const double *rates);
void rdo_set_category_weights(rdo_partition_t *p, const double *weights);
int  rdo_update_invariant_sites(rdo_partition_t *p);
int  rdo_update_invariant_sites_proportion(rdo_partition_t *p,
                                           unsigned int params_index,
                                           double prop_invar);

int rdo_update_prob_matrices(rdo_partition_t *p,
                             const unsigned int *params_indices,
                             const unsigned int *matrix_indices,
                             const double *branch_lengths,
                             unsigned int count);
void rdo_update_clvs(rdo_partition_t *p, const rdo_operation_t *ops,
                     unsigned int count);
double rdo_compute_root_loglikelihood(rdo_partition_t *p,
                                      unsigned int clv_index,
                                      int scaler_index,
                                      const unsigned int *freqs_indices,
                                      double *persite_lnl);
/* same, with explicit arithmetic mode (RDO_MODE_*) */
double rdo_compute_root_loglikelihood_mode(rdo_partition_t *p,
                                           unsigned int clv_index,
                                           int scaler_index,
                                           const unsigned int *freqs_indices,
                                           double *persite_lnl,
                                           int mode);
/* process-wide default mode used by rdo_compute_root_loglikelihood */
void rdo_set_default_mode(int mode);

int     rdo_compute_gamma_cats(double alpha, unsigned int categories,
                               double *output_rates, int rates_mode);
double *rdo_msa_empirical_frequencies(rdo_partition_t *p);

/* building blocks exposed for unit tests */
/* H1 switches of the Q builder (rd_oracle.c, "Q matrix"): 0 = the convention of SURVEY Appendix A-2
 * (the one the CUDA engine implements); any combination of the flags selects another reading of
 * coraxlib's non-reversible builder.  Process-global; affects P-matrices built afterwards. */
#define RDO_Q_NO_PI 1
#define RDO_Q_SLOTS_COLUMN_MAJOR 2
#define RDO_Q_NO_NORMALISATION 4
void rdo_set_q_convention(int flags);
int  rdo_get_q_convention(void);
void   rdo_build_q_nonrev(const double *subst_params, const double *freqs,
                          double *Q /*16*/);
void   rdo_expm4(const double *A /*16*/, double *E /*16*/);
double rdo_log(double x); /* the spec'd software log (ENGINE mode) */
double rdo_pairwise_sum(const double *v, unsigned long n);

/* OpenMP site-parallel variants used only by the bench CPU baseline
 * ("all-cores site-parallel oracle"); results are
 * identical to the serial functions for update_clvs; the log-likelihood uses
 * per-thread partial sums. */
void   rdo_update_clvs_mt(rdo_partition_t *p, const rdo_operation_t *ops,
                          unsigned int count, int threads);
double rdo_compute_root_loglikelihood_mt(rdo_partition_t *p,
                                         unsigned int clv_index,
                                         int scaler_index, int threads);

#ifdef __cplusplus
}
#endif
#endif
