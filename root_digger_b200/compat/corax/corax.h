/*
 * corax/corax.h -- compatibility header: the subset of coraxlib's C API that RootDigger's own
 * sources use (src/model.cpp, src/tree.cpp, src/msa.cpp, src/checkpoint.cpp),
 * served by the B200 likelihood engine.
 *
 *   - the partition / likelihood calls (corax_partition_create ... corax_compute_root_loglikelihood,
 *     src/model.cpp:159-168,185,205-209,244-297,310-347,367,402-466) are the engine's C ABI under
 *     their corax names: rdk.h entry points have the same argument order and meaning, so these are
 *     plain name mappings, no wrapper code;
 *   - the unrooted-tree module (corax_unode_t / corax_utree_t and the seven corax_utree_* calls of
 *     src/tree.cpp) and the alignment readers (corax_msa_t, PHYLIP / FASTA, pattern compression,
 *     src/msa.cpp:18-88,621-632) are host-only C code in corax_compat.cpp.
 *
 * With this directory first on the include path, RootDigger's src/ (minus main.cpp's CLI) compiles
 * UNCHANGED against the engine: tests/ref_build does exactly that with the files where they lie
 * under the reference checkout.  Include order: <rdk.h> is found through the -I list, so the same
 * header also serves a build against any other implementation of that ABI.
 */
#ifndef CORAX_COMPAT_CORAX_H_
#define CORAX_COMPAT_CORAX_H_

#include <rdk.h>

/* coraxlib's header pulls these in, and RootDigger's sources rely on that (memcpy, sqrt, ...) */
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
#include <cmath>
#include <cstring>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status, attributes, modes ------------------------------------------------------------ */
/* (an implementation of the ABI that selects no CPU kernel flavour may leave the ARCH bits out) */
#ifndef RDK_ATTRIB_ARCH_CPU
#define RDK_ATTRIB_ARCH_CPU 0u
#endif
#ifndef RDK_ATTRIB_ARCH_SSE
#define RDK_ATTRIB_ARCH_SSE (1u << 0)
#endif
#ifndef RDK_ATTRIB_ARCH_AVX
#define RDK_ATTRIB_ARCH_AVX (1u << 1)
#endif
#ifndef RDK_ATTRIB_ARCH_AVX2
#define RDK_ATTRIB_ARCH_AVX2 (1u << 2)
#endif
#define CORAX_SUCCESS RDK_SUCCESS
#define CORAX_FAILURE RDK_FAILURE
#define CORAX_SCALE_BUFFER_NONE RDK_SCALE_BUFFER_NONE
#define CORAX_ATTRIB_ARCH_CPU RDK_ATTRIB_ARCH_CPU
#define CORAX_ATTRIB_ARCH_SSE RDK_ATTRIB_ARCH_SSE
#define CORAX_ATTRIB_ARCH_AVX RDK_ATTRIB_ARCH_AVX
#define CORAX_ATTRIB_ARCH_AVX2 RDK_ATTRIB_ARCH_AVX2
#define CORAX_ATTRIB_SITE_REPEATS RDK_ATTRIB_SITE_REPEATS
#define CORAX_ATTRIB_NONREV RDK_ATTRIB_NONREV
#define CORAX_GAMMA_RATES_MEAN RDK_GAMMA_RATES_MEAN
#define CORAX_GAMMA_RATES_MEDIAN RDK_GAMMA_RATES_MEDIAN
/* src/model.cpp:145-157 picks a CPU kernel flavour; the engine has none to pick */
#define CORAX_HAS_CPU_FEATURE(x) 0

#define corax_errno rdk_errno
#define corax_errmsg rdk_errmsg

/* ---- partition + likelihood: the engine's ABI under its corax names ---------------------------- */
typedef rdk_partition_t corax_partition_t;
typedef rdk_operation_t corax_operation_t;
typedef rdk_state_t     corax_state_t;
#define corax_map_nt rdk_map_nt
#define corax_partition_create rdk_partition_create
#define corax_partition_destroy rdk_partition_destroy
#define corax_set_tip_states rdk_set_tip_states
#define corax_set_pattern_weights rdk_set_pattern_weights
#define corax_set_subst_params rdk_set_subst_params
#define corax_set_frequencies rdk_set_frequencies
#define corax_set_category_rates rdk_set_category_rates
#define corax_set_category_weights rdk_set_category_weights
#define corax_update_invariant_sites rdk_update_invariant_sites
#define corax_update_invariant_sites_proportion rdk_update_invariant_sites_proportion
#define corax_update_prob_matrices rdk_update_prob_matrices
#define corax_update_clvs rdk_update_clvs
#define corax_compute_root_loglikelihood rdk_compute_root_loglikelihood
#define corax_compute_gamma_cats rdk_compute_gamma_cats
#define corax_msa_empirical_frequencies rdk_msa_empirical_frequencies

/* ---- unrooted trees (src/tree.cpp) ------------------------------------------------------------- */
typedef struct corax_unode_s {
  char                 *label;
  double                length;
  unsigned int          node_index;
  unsigned int          clv_index;
  int                   scaler_index;
  unsigned int          pmatrix_index;
  struct corax_unode_s *next;
  struct corax_unode_s *back;
  void                 *data;
} corax_unode_t;

typedef struct corax_utree_s {
  unsigned int    tip_count;
  unsigned int    inner_count;
  unsigned int    edge_count;
  int             binary;
  corax_unode_t **nodes; /* tips first (by clv index), then one unode per inner node */
  corax_unode_t  *vroot;
} corax_utree_t;

#define CORAX_TREE_TRAVERSE_POSTORDER 1
#define CORAX_UTREE_SHOW_LABEL (1 << 0)
#define CORAX_UTREE_SHOW_BRANCH_LENGTH (1 << 1)

/* src/tree.cpp:12 */
corax_utree_t *corax_utree_parse_newick_unroot(const char *filename);
/* src/tree.cpp:30,66 */
corax_utree_t *corax_utree_clone(const corax_utree_t *tree);
/* src/tree.cpp:49 */
void corax_utree_destroy(corax_utree_t *tree, void (*cb_destroy)(void *));
/* src/tree.cpp:247,262,603 */
int corax_utree_traverse(corax_unode_t *root, int traversal, int (*cbtrav)(corax_unode_t *),
                         corax_unode_t **outbuffer, unsigned int *trav_size);
/* src/tree.cpp:387,620 */
void corax_utree_create_operations(corax_unode_t *const *trav_buffer, unsigned int trav_buffer_size,
                                   double *branches, unsigned int *pmatrix_indices,
                                   corax_operation_t *ops, unsigned int *matrix_count,
                                   unsigned int *ops_count);
/* src/tree.cpp:488 (the result and every string the callback returns are malloc'd; both are freed
 * with free(), the callback's by this function) */
char *corax_utree_export_newick(const corax_unode_t *root,
                                char *(*cb_serialize)(const corax_unode_t *));
/* src/tree.cpp:495 */
void corax_utree_show_ascii(const corax_unode_t *tree, int options);

/* ---- alignments (src/msa.cpp) ------------------------------------------------------------------- */
typedef struct corax_msa_s {
  int    count;
  int    length;
  char **sequence;
  char **label;
} corax_msa_t;
typedef struct corax_phylip_s corax_phylip_t;
typedef struct corax_fasta_s  corax_fasta_t;
extern const unsigned int     corax_map_generic[256];
extern const unsigned int     corax_map_fasta[256];

corax_phylip_t *corax_phylip_open(const char *filename, const unsigned int *map);
int             corax_phylip_rewind(corax_phylip_t *fd);
void            corax_phylip_close(corax_phylip_t *fd);
corax_msa_t    *corax_phylip_parse_interleaved(corax_phylip_t *fd);
corax_msa_t    *corax_phylip_parse_sequential(corax_phylip_t *fd);
corax_fasta_t  *corax_fasta_open(const char *filename, const unsigned int *map);
int             corax_fasta_getnext(corax_fasta_t *fd, char **head, long *head_len, char **seq,
                                    long *seq_len, long *seqno);
void            corax_fasta_close(corax_fasta_t *fd);
void            corax_msa_destroy(corax_msa_t *msa);
/* src/msa.hpp:30, src/msa.cpp:624: identical columns (compared through `map`) are merged, the
 * surviving patterns come out sorted, *length becomes their number, the malloc'd result holds
 * their weights */
unsigned int *corax_compress_site_patterns(char **sequence, const corax_state_t *map, int count,
                                           int *length);

#ifdef __cplusplus
}
#endif
#endif
