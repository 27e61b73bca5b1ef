/*
 * rd_oracle.c -- CPU ORACLE (test infrastructure, see rd_oracle.h).
 *
 * "parity unpinned": coraxlib itself is absent from /root/reference; this is a
 * restatement of the libpll-2/coraxlib algorithm as documented in SURVEY.md
 * Appendix A, anchored on RootDigger's call sites (cited per function).
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fno-fast-math (see
 * oracle/Makefile).  Arithmetic specification v2: the 4x4 mat-vec of the CLV
 * update and the dot products of the root log-likelihood are explicit fma()
 * chains (coraxlib's AVX2 kernels are FMA kernels too); every other +,-,*,/ is
 * individually rounded (-ffp-contract=off: no implicit contraction), in the
 * order written here.  The CUDA engine follows the same order with
 * __fma_rn/__dmul_rn/__dadd_rn so that ENGINE mode results can be compared bit
 * for bit.
 */
#include "rd_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int  rdo_errno = 0;
char rdo_errmsg[200] = {0};

static int g_default_mode = RDO_MODE_REFERENCE;
void rdo_set_default_mode(int mode) { g_default_mode = mode; }

/* 2^-256 underflow threshold and its inverse (libpll PLL_SCALE_THRESHOLD /
 * PLL_SCALE_FACTOR). */
#define RDO_SCALE_THRESHOLD 0x1p-256
#define RDO_SCALE_FACTOR 0x1p+256
/* ln(2^-256) = -256 ln 2, correctly rounded */
#define RDO_LOG_SCALE_THRESHOLD (-177.44567822334599)

/* ------------------------------------------------------------------------ */
/* nucleotide map (corax_map_nt; used at src/model.cpp:312 via msa.map())    */
/* A=1 C=2 G=4 T/U=8, IUPAC unions, gap/N/X/O/? = 15, everything else 0      */
/* ------------------------------------------------------------------------ */
#define NT(ch, v) [ch] = v, [ch + 32] = v
const rdo_state_t rdo_map_nt[256] = {
    ['-'] = 15, ['?'] = 15,
    NT('A', 1),  NT('B', 14), NT('C', 2),  NT('D', 13), NT('G', 4),
    NT('H', 11), NT('K', 12), NT('M', 3),  NT('N', 15), NT('O', 15),
    NT('R', 5),  NT('S', 6),  NT('T', 8),  NT('U', 8),  NT('V', 7),
    NT('W', 9),  NT('X', 15), NT('Y', 10),
};
#undef NT

/* ------------------------------------------------------------------------ */
/* partition (corax_partition_create, src/model.cpp:159-168; Appendix A-1)   */
/* ------------------------------------------------------------------------ */
static void set_error(const char *msg) {
  rdo_errno = 1;
  snprintf(rdo_errmsg, sizeof(rdo_errmsg), "%s", msg);
}

rdo_partition_t *rdo_partition_create(unsigned int tips,
                                      unsigned int clv_buffers,
                                      unsigned int states, unsigned int sites,
                                      unsigned int rate_matrices,
                                      unsigned int prob_matrices,
                                      unsigned int rate_cats,
                                      unsigned int scale_buffers,
                                      unsigned int attributes) {
  if (states != 4) {
    set_error("oracle supports 4 states only");
    return NULL;
  }
  if (rate_cats == 0 || rate_matrices == 0) {
    set_error("rate_cats and rate_matrices must be positive");
    return NULL;
  }
  rdo_partition_t *p = (rdo_partition_t *)calloc(1, sizeof(*p));
  if (!p) return NULL;
  p->tips = tips;
  p->clv_buffers = clv_buffers;
  p->states = states;
  p->sites = sites;
  p->rate_matrices = rate_matrices;
  p->prob_matrices = prob_matrices;
  p->rate_cats = rate_cats;
  p->scale_buffers = scale_buffers;
  p->attributes = attributes;

  size_t clv_len = (size_t)sites * rate_cats * states;
  p->clv = (double **)calloc(tips + clv_buffers, sizeof(double *));
  for (unsigned i = 0; i < tips + clv_buffers; ++i)
    p->clv[i] = (double *)calloc(clv_len ? clv_len : 1, sizeof(double));
  p->pmatrix = (double **)calloc(prob_matrices, sizeof(double *));
  for (unsigned i = 0; i < prob_matrices; ++i)
    p->pmatrix[i] = (double *)calloc((size_t)rate_cats * 16, sizeof(double));
  p->scale_buffer = (unsigned int **)calloc(scale_buffers, sizeof(unsigned *));
  for (unsigned i = 0; i < scale_buffers; ++i)
    p->scale_buffer[i] = (unsigned *)calloc(sites ? sites : 1, sizeof(unsigned));
  p->subst_params = (double **)calloc(rate_matrices, sizeof(double *));
  p->frequencies = (double **)calloc(rate_matrices, sizeof(double *));
  for (unsigned i = 0; i < rate_matrices; ++i) {
    p->subst_params[i] = (double *)calloc(12, sizeof(double));
    p->frequencies[i] = (double *)calloc(4, sizeof(double));
    for (int j = 0; j < 12; ++j) p->subst_params[i][j] = 1.0;
    for (int j = 0; j < 4; ++j) p->frequencies[i][j] = 0.25;
  }
  p->rates = (double *)calloc(rate_cats, sizeof(double));
  p->rate_weights = (double *)calloc(rate_cats, sizeof(double));
  for (unsigned i = 0; i < rate_cats; ++i) {
    p->rates[i] = 1.0;
    p->rate_weights[i] = 1.0 / rate_cats;
  }
  p->prop_invar = (double *)calloc(rate_matrices, sizeof(double));
  p->pattern_weights = (unsigned *)calloc(sites ? sites : 1, sizeof(unsigned));
  for (unsigned i = 0; i < sites; ++i) p->pattern_weights[i] = 1;
  p->invariant = NULL;
  return p;
}

void rdo_partition_destroy(rdo_partition_t *p) {
  if (!p) return;
  for (unsigned i = 0; i < p->tips + p->clv_buffers; ++i) free(p->clv[i]);
  free(p->clv);
  for (unsigned i = 0; i < p->prob_matrices; ++i) free(p->pmatrix[i]);
  free(p->pmatrix);
  for (unsigned i = 0; i < p->scale_buffers; ++i) free(p->scale_buffer[i]);
  free(p->scale_buffer);
  for (unsigned i = 0; i < p->rate_matrices; ++i) {
    free(p->subst_params[i]);
    free(p->frequencies[i]);
  }
  free(p->subst_params);
  free(p->frequencies);
  free(p->rates);
  free(p->rate_weights);
  free(p->prop_invar);
  free(p->pattern_weights);
  free(p->invariant);
  free(p);
}

/* corax_set_tip_states (src/model.cpp:310-313): tip CLV entries are 0/1 from
 * the state bit mask, replicated over rate categories; unknown character ->
 * failure. */
int rdo_set_tip_states(rdo_partition_t *p, unsigned int tip_index,
                       const rdo_state_t *map, const char *sequence) {
  if (tip_index >= p->tips) {
    set_error("tip index out of range");
    return RDO_FAILURE;
  }
  double  *clv = p->clv[tip_index];
  unsigned K = p->rate_cats;
  for (unsigned s = 0; s < p->sites; ++s) {
    rdo_state_t st = map[(unsigned char)sequence[s]];
    if (!st) {
      snprintf(rdo_errmsg, sizeof(rdo_errmsg),
               "Illegal state code in tip \"%c\"", sequence[s]);
      rdo_errno = 1;
      return RDO_FAILURE;
    }
    for (unsigned k = 0; k < K; ++k)
      for (unsigned j = 0; j < 4; ++j)
        clv[((size_t)s * K + k) * 4 + j] = (double)((st >> j) & 1ULL);
  }
  return RDO_SUCCESS;
}

void rdo_set_pattern_weights(rdo_partition_t *p, const unsigned int *w) {
  memcpy(p->pattern_weights, w, sizeof(unsigned) * p->sites);
}
void rdo_set_subst_params(rdo_partition_t *p, unsigned int idx,
                          const double *params) {
  memcpy(p->subst_params[idx], params, sizeof(double) * 12);
}
void rdo_set_frequencies(rdo_partition_t *p, unsigned int idx,
                         const double *freqs) {
  memcpy(p->frequencies[idx], freqs, sizeof(double) * 4);
}
void rdo_set_category_rates(rdo_partition_t *p, const double *rates) {
  memcpy(p->rates, rates, sizeof(double) * p->rate_cats);
}
void rdo_set_category_weights(rdo_partition_t *p, const double *weights) {
  memcpy(p->rate_weights, weights, sizeof(double) * p->rate_cats);
}

/* corax_update_invariant_sites (src/model.cpp:294): mark columns whose tips
 * share a state.  RootDigger never sets a non-zero proportion (Appendix B-5),
 * so the marks do not enter the likelihood. */
int rdo_update_invariant_sites(rdo_partition_t *p) {
  if (!p->invariant) p->invariant = (int *)malloc(sizeof(int) * (p->sites + 1));
  unsigned K = p->rate_cats;
  for (unsigned s = 0; s < p->sites; ++s) {
    unsigned mask = 15;
    for (unsigned t = 0; t < p->tips; ++t) {
      unsigned m = 0;
      for (unsigned j = 0; j < 4; ++j)
        if (p->clv[t][((size_t)s * K) * 4 + j] != 0.0) m |= 1u << j;
      mask &= m;
    }
    /* single shared state -> its index, otherwise -1 */
    int inv = -1;
    if (mask && !(mask & (mask - 1))) inv = __builtin_ctz(mask);
    p->invariant[s] = inv;
  }
  return RDO_SUCCESS;
}
int rdo_update_invariant_sites_proportion(rdo_partition_t *p, unsigned int idx,
                                          double prop_invar) {
  if (prop_invar < 0.0 || prop_invar >= 1.0) {
    set_error("Invalid proportion of invariant sites");
    return RDO_FAILURE;
  }
  p->prop_invar[idx] = prop_invar;
  return RDO_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* Q matrix (Appendix A-2): Q_ij = r_(ij) * pi_j, 12 r's in row-major         */
/* off-diagonal order AC,AG,AT,CA,CG,CT,GA,GC,GT,TA,TC,TG; diagonal = -row    */
/* sum; normalised so that -sum_i pi_i Q_ii = 1.                              */
/*                                                                            */
/* H1: the three choices above are the UNVERIFIED reading   */
/* of coraxlib's non-reversible builder (the library is absent).  They are     */
/* switchable at run time -- rdo_set_q_convention -- so that the day a coraxlib*/
/* number is available, pinning the oracle is a matter of selecting the        */
/* variant that reproduces it.  The default (0) is the convention the CUDA     */
/* engine implements; the engine must be changed together with the default.    */
/* ------------------------------------------------------------------------ */
static int g_q_convention = 0;
void rdo_set_q_convention(int flags) { g_q_convention = flags; }
int  rdo_get_q_convention(void) { return g_q_convention; }

void rdo_build_q_nonrev(const double *r, const double *pi, double *Q) {
  const int no_pi = g_q_convention & RDO_Q_NO_PI;         /* Q_ij = r_(ij), frequencies only at the root */
  const int col_major = g_q_convention & RDO_Q_SLOTS_COLUMN_MAJOR; /* r's listed column by column       */
  const int no_norm = g_q_convention & RDO_Q_NO_NORMALISATION;     /* rates taken as absolute            */
  int k = 0;
  if (!col_major) {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        if (i != j) Q[i * 4 + j] = no_pi ? r[k++] : r[k++] * pi[j];
  } else {
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i)
        if (i != j) Q[i * 4 + j] = no_pi ? r[k++] : r[k++] * pi[j];
  }
  for (int i = 0; i < 4; ++i) {
    double s = 0.0;
    int first = 1;
    for (int j = 0; j < 4; ++j) {
      if (j == i) continue;
      if (first) {
        s = Q[i * 4 + j];
        first = 0;
      } else
        s = s + Q[i * 4 + j];
    }
    Q[i * 4 + i] = -s;
  }
  if (no_norm) return;
  double mu = pi[0] * (-Q[0]);
  mu = mu + pi[1] * (-Q[5]);
  mu = mu + pi[2] * (-Q[10]);
  mu = mu + pi[3] * (-Q[15]);
  for (int i = 0; i < 16; ++i) Q[i] = Q[i] / mu;
}

/* ------------------------------------------------------------------------ */
/* expm, 4x4: Higham (2005) scaling-and-squaring with Pade approximants of   */
/* degree 3/5/7/9/13 (the algorithm behind scipy.linalg.expm).  Operation     */
/* order is part of the spec (DESIGN.md "expm4").                             */
/* ------------------------------------------------------------------------ */
static void mm4(const double *A, const double *B, double *C) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = A[i * 4 + 0] * B[0 * 4 + j];
      s = s + A[i * 4 + 1] * B[1 * 4 + j];
      s = s + A[i * 4 + 2] * B[2 * 4 + j];
      s = s + A[i * 4 + 3] * B[3 * 4 + j];
      C[i * 4 + j] = s;
    }
}

static double norm1_4(const double *A) {
  double best = 0.0;
  for (int j = 0; j < 4; ++j) {
    double s = fabs(A[0 * 4 + j]);
    s = s + fabs(A[1 * 4 + j]);
    s = s + fabs(A[2 * 4 + j]);
    s = s + fabs(A[3 * 4 + j]);
    if (s > best) best = s;
  }
  return best;
}

/* X = M^-1 N by Gaussian elimination with partial pivoting (first maximal
 * pivot), row operations in ascending order, then back substitution. */
static void solve4(double *M, double *N, double *X) {
  for (int c = 0; c < 4; ++c) {
    int    piv = c;
    double best = fabs(M[c * 4 + c]);
    for (int r = c + 1; r < 4; ++r) {
      double v = fabs(M[r * 4 + c]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    if (piv != c) {
      for (int j = 0; j < 4; ++j) {
        double t = M[c * 4 + j];
        M[c * 4 + j] = M[piv * 4 + j];
        M[piv * 4 + j] = t;
        t = N[c * 4 + j];
        N[c * 4 + j] = N[piv * 4 + j];
        N[piv * 4 + j] = t;
      }
    }
    for (int r = c + 1; r < 4; ++r) {
      double f = M[r * 4 + c] / M[c * 4 + c];
      for (int j = c + 1; j < 4; ++j)
        M[r * 4 + j] = M[r * 4 + j] - f * M[c * 4 + j];
      for (int j = 0; j < 4; ++j) N[r * 4 + j] = N[r * 4 + j] - f * N[c * 4 + j];
    }
  }
  for (int r = 3; r >= 0; --r)
    for (int j = 0; j < 4; ++j) {
      double s = N[r * 4 + j];
      for (int q = r + 1; q < 4; ++q) s = s - M[r * 4 + q] * X[q * 4 + j];
      X[r * 4 + j] = s / M[r * 4 + r];
    }
}

static const double PADE3[4] = {120., 60., 12., 1.};
static const double PADE5[6] = {30240., 15120., 3360., 420., 30., 1.};
static const double PADE7[8] = {17297280., 8648640., 1995840., 277200.,
                                25200.,    1512.,    56.,      1.};
static const double PADE9[10] = {17643225600., 8821612800., 2075673600.,
                                 302702400.,   30270240.,   2162160.,
                                 110880.,      3960.,       90.,
                                 1.};
static const double PADE13[14] = {64764752532480000.,
                                  32382376266240000.,
                                  7771770303897600.,
                                  1187353796428800.,
                                  129060195264000.,
                                  10559470521600.,
                                  670442572800.,
                                  33522128640.,
                                  1323241920.,
                                  40840800.,
                                  960960.,
                                  16380.,
                                  182.,
                                  1.};

#define TH3 1.495585217958292e-2
#define TH5 2.539398330063230e-1
#define TH7 9.504178996162932e-1
#define TH9 2.097847961257068e0
#define TH13 5.371920351148152e0

/* Upoly = sum_{odd} b[2m+1] A^{2m}, Vpoly = sum_{even} b[2m] A^{2m} for the
 * low degrees: terms are accumulated from the highest power down to the
 * identity term: ((b_hi*A_hi + ... ) + b_lo*A2) + b0*I. */
static void pade_low(const double *A, const double *b, int deg, double *U,
                     double *V) {
  double A2[16], A4[16], A6[16], A8[16], W[16];
  mm4(A, A, A2);
  if (deg >= 5) mm4(A2, A2, A4);
  if (deg >= 7) mm4(A4, A2, A6);
  if (deg >= 9) mm4(A6, A2, A8);
  for (int i = 0; i < 16; ++i) {
    double id = (i % 5 == 0) ? 1.0 : 0.0;
    double w, v;
    if (deg == 3) {
      w = b[3] * A2[i];
      v = b[2] * A2[i];
    } else if (deg == 5) {
      w = b[5] * A4[i];
      w = w + b[3] * A2[i];
      v = b[4] * A4[i];
      v = v + b[2] * A2[i];
    } else if (deg == 7) {
      w = b[7] * A6[i];
      w = w + b[5] * A4[i];
      w = w + b[3] * A2[i];
      v = b[6] * A6[i];
      v = v + b[4] * A4[i];
      v = v + b[2] * A2[i];
    } else {
      w = b[9] * A8[i];
      w = w + b[7] * A6[i];
      w = w + b[5] * A4[i];
      w = w + b[3] * A2[i];
      v = b[8] * A8[i];
      v = v + b[6] * A6[i];
      v = v + b[4] * A4[i];
      v = v + b[2] * A2[i];
    }
    W[i] = w + b[1] * id;
    V[i] = v + b[0] * id;
  }
  mm4(A, W, U);
}

static void pade13(const double *A, double *U, double *V) {
  const double *b = PADE13;
  double A2[16], A4[16], A6[16], W1[16], W2[16], Z1[16], Z2[16], W[16];
  mm4(A, A, A2);
  mm4(A2, A2, A4);
  mm4(A4, A2, A6);
  for (int i = 0; i < 16; ++i) {
    double id = (i % 5 == 0) ? 1.0 : 0.0;
    double t = b[13] * A6[i];
    t = t + b[11] * A4[i];
    t = t + b[9] * A2[i];
    W1[i] = t;
    t = b[7] * A6[i];
    t = t + b[5] * A4[i];
    t = t + b[3] * A2[i];
    t = t + b[1] * id;
    W2[i] = t;
    t = b[12] * A6[i];
    t = t + b[10] * A4[i];
    t = t + b[8] * A2[i];
    Z1[i] = t;
    t = b[6] * A6[i];
    t = t + b[4] * A4[i];
    t = t + b[2] * A2[i];
    t = t + b[0] * id;
    Z2[i] = t;
  }
  double T[16];
  mm4(A6, W1, T);
  for (int i = 0; i < 16; ++i) W[i] = T[i] + W2[i];
  mm4(A, W, U);
  mm4(A6, Z1, T);
  for (int i = 0; i < 16; ++i) V[i] = T[i] + Z2[i];
}

void rdo_expm4(const double *Ain, double *E) {
  double A[16], U[16], V[16], M[16], N[16];
  memcpy(A, Ain, sizeof(A));
  double n1 = norm1_4(A);
  int    s = 0;
  if (n1 <= TH3)
    pade_low(A, PADE3, 3, U, V);
  else if (n1 <= TH5)
    pade_low(A, PADE5, 5, U, V);
  else if (n1 <= TH7)
    pade_low(A, PADE7, 7, U, V);
  else if (n1 <= TH9)
    pade_low(A, PADE9, 9, U, V);
  else {
    /* smallest s >= 0 with n1 * 2^-s <= theta13 (halving is exact) */
    while (n1 > TH13) {
      n1 = n1 * 0.5;
      ++s;
    }
    double sc = ldexp(1.0, -s);
    for (int i = 0; i < 16; ++i) A[i] = A[i] * sc;
    pade13(A, U, V);
  }
  for (int i = 0; i < 16; ++i) {
    M[i] = V[i] - U[i];
    N[i] = V[i] + U[i];
  }
  solve4(M, N, E);
  for (int q = 0; q < s; ++q) {
    double T[16];
    mm4(E, E, T);
    memcpy(E, T, sizeof(T));
  }
}

/* corax_update_prob_matrices (src/model.cpp:367,432,842):
 * P_{b,k} = expm(Q * rate_k * t_b / (1 - p_inv)), layout [cat][i][j]. */
int rdo_update_prob_matrices(rdo_partition_t *p,
                             const unsigned int *params_indices,
                             const unsigned int *matrix_indices,
                             const double *branch_lengths,
                             unsigned int count) {
  unsigned K = p->rate_cats;
  for (unsigned b = 0; b < count; ++b) {
    unsigned mi = matrix_indices[b];
    if (mi >= p->prob_matrices) {
      set_error("matrix index out of range");
      return RDO_FAILURE;
    }
    double t = branch_lengths[b];
    if (!(t >= 0.0) || !isfinite(t)) {
      set_error("branch length must be finite and non-negative");
      return RDO_FAILURE;
    }
    for (unsigned k = 0; k < K; ++k) {
      unsigned pi = params_indices ? params_indices[k] : 0;
      double   Q[16], A[16];
      rdo_build_q_nonrev(p->subst_params[pi], p->frequencies[pi], Q);
      double c = (p->rates[k] * t) / (1.0 - p->prop_invar[pi]);
      for (int i = 0; i < 16; ++i) A[i] = Q[i] * c;
      rdo_expm4(A, p->pmatrix[mi] + (size_t)k * 16);
    }
  }
  return RDO_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* corax_update_clvs (src/model.cpp:402,440,461,851; Appendix A-3)           */
/* ------------------------------------------------------------------------ */
static inline void clv_site(const double *P1, const double *P2,
                            const double *c1, const double *c2, double *out,
                            unsigned K, int *all_small) {
  int small = 1;
  for (unsigned k = 0; k < K; ++k) {
    const double *p1 = P1 + k * 16, *p2 = P2 + k * 16;
    const double *a = c1 + k * 4, *b = c2 + k * 4;
    for (int i = 0; i < 4; ++i) {
      /* fused multiply-add chain, j ascending (arithmetic spec v2: the 4x4
       * mat-vec and the root dot products are FMA chains, as in coraxlib's
       * AVX2/FMA kernels; everything else is individually rounded) */
      double x = p1[i * 4 + 0] * a[0];
      x = fma(p1[i * 4 + 1], a[1], x);
      x = fma(p1[i * 4 + 2], a[2], x);
      x = fma(p1[i * 4 + 3], a[3], x);
      double y = p2[i * 4 + 0] * b[0];
      y = fma(p2[i * 4 + 1], b[1], y);
      y = fma(p2[i * 4 + 2], b[2], y);
      y = fma(p2[i * 4 + 3], b[3], y);
      double v = x * y;
      out[k * 4 + i] = v;
      if (!(v < RDO_SCALE_THRESHOLD)) small = 0;
    }
  }
  *all_small = small;
}

static void update_clv_range(rdo_partition_t *p, const rdo_operation_t *op,
                             unsigned s0, unsigned s1) {
  unsigned        K = p->rate_cats;
  const double   *P1 = p->pmatrix[op->child1_matrix_index];
  const double   *P2 = p->pmatrix[op->child2_matrix_index];
  const double   *c1 = p->clv[op->child1_clv_index];
  const double   *c2 = p->clv[op->child2_clv_index];
  double         *out = p->clv[op->parent_clv_index];
  unsigned       *ps = op->parent_scaler_index == RDO_SCALE_BUFFER_NONE
                           ? NULL
                           : p->scale_buffer[op->parent_scaler_index];
  const unsigned *s1b = op->child1_scaler_index == RDO_SCALE_BUFFER_NONE
                            ? NULL
                            : p->scale_buffer[op->child1_scaler_index];
  const unsigned *s2b = op->child2_scaler_index == RDO_SCALE_BUFFER_NONE
                            ? NULL
                            : p->scale_buffer[op->child2_scaler_index];
  size_t span = (size_t)K * 4;
  for (unsigned s = s0; s < s1; ++s) {
    int small;
    clv_site(P1, P2, c1 + s * span, c2 + s * span, out + s * span, K, &small);
    if (ps) {
      unsigned cnt = (s1b ? s1b[s] : 0u) + (s2b ? s2b[s] : 0u);
      if (small) {
        for (size_t q = 0; q < span; ++q)
          out[s * span + q] = out[s * span + q] * RDO_SCALE_FACTOR;
        cnt += 1;
      }
      ps[s] = cnt;
    }
  }
}

void rdo_update_clvs(rdo_partition_t *p, const rdo_operation_t *ops,
                     unsigned int count) {
  for (unsigned o = 0; o < count; ++o) update_clv_range(p, &ops[o], 0, p->sites);
}

void rdo_update_clvs_mt(rdo_partition_t *p, const rdo_operation_t *ops,
                        unsigned int count, int threads) {
  if (threads < 1) threads = 1;
  unsigned S = p->sites;
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    extern int omp_get_thread_num(void);
    extern int omp_get_num_threads(void);
    int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int tid = 0, nt = 1;
#endif
    unsigned s0 = (unsigned)((unsigned long long)S * tid / nt);
    unsigned s1 = (unsigned)((unsigned long long)S * (tid + 1) / nt);
    /* sites are independent: each thread walks the whole op list on its own
     * site range, no barrier needed */
    for (unsigned o = 0; o < count; ++o) update_clv_range(p, &ops[o], s0, s1);
  }
}

/* ------------------------------------------------------------------------ */
/* the spec'd software log (ENGINE mode): argument reduction to              */
/* [sqrt(2)/2, sqrt(2)) and the classic degree-14 minimax in s=f/(2+f)       */
/* (the fdlibm/musl formulation), every operation individually rounded.      */
/* ------------------------------------------------------------------------ */
double rdo_log(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01,
                      ln2_lo = 1.90821492927058770002e-10,
                      Lg1 = 6.666666666666735130e-01,
                      Lg2 = 3.999999999940941908e-01,
                      Lg3 = 2.857142874366239149e-01,
                      Lg4 = 2.222219843214978396e-01,
                      Lg5 = 1.818357216161805012e-01,
                      Lg6 = 1.531383769920937332e-01,
                      Lg7 = 1.479819860511658591e-01;
  union {
    double   f;
    uint64_t i;
  } u = {x};
  uint32_t hx = (uint32_t)(u.i >> 32);
  int      k = 0;
  if (hx < 0x00100000 || hx >> 31) {
    if ((u.i << 1) == 0) return -INFINITY; /* log(+-0) */
    if (hx >> 31) return NAN;              /* log(-#)  */
    /* subnormal: scale up by 2^54 */
    k -= 54;
    x = x * 0x1p54;
    u.f = x;
    hx = (uint32_t)(u.i >> 32);
  } else if (hx >= 0x7ff00000) {
    return x; /* inf or nan */
  } else if (hx == 0x3ff00000 && (u.i << 32) == 0) {
    return 0.0;
  }
  hx += 0x3ff00000 - 0x3fe6a09e;
  k += (int)(hx >> 20) - 0x3ff;
  hx = (hx & 0x000fffff) + 0x3fe6a09e;
  u.i = ((uint64_t)hx << 32) | (u.i & 0xffffffffULL);
  x = u.f;

  double f = x - 1.0;
  double hfsq = (0.5 * f) * f;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double dk = (double)k;
  /* ((((s*(hfsq+R)) + dk*ln2_lo) - hfsq) + f) + dk*ln2_hi */
  double r = s * (hfsq + R);
  r = r + dk * ln2_lo;
  r = r - hfsq;
  r = r + f;
  r = r + dk * ln2_hi;
  return r;
}

/* canonical reduction (ENGINE mode): balanced binary tree over contiguous
 * halves of the site array zero-padded to the next power of two. */
static double pairwise_rec(const double *v, unsigned long lo, unsigned long n,
                           unsigned long len) {
  if (lo >= len) return 0.0;
  if (n == 1) return v[lo];
  unsigned long h = n >> 1;
  double a = pairwise_rec(v, lo, h, len);
  double b = pairwise_rec(v, lo + h, h, len);
  return a + b;
}
double rdo_pairwise_sum(const double *v, unsigned long n) {
  if (n == 0) return 0.0;
  unsigned long N = 1;
  while (N < n) N <<= 1;
  return pairwise_rec(v, 0, N, n);
}

/* corax_compute_root_loglikelihood (src/model.cpp:406,441,466; Appendix A-4) */
static inline double site_term(const rdo_partition_t *p, const double *clv,
                               const double *pi, unsigned s) {
  unsigned K = p->rate_cats;
  double   term = 0.0;
  for (unsigned k = 0; k < K; ++k) {
    const double *c = clv + ((size_t)s * K + k) * 4;
    double        t = pi[0] * c[0];
    t = fma(pi[1], c[1], t);
    t = fma(pi[2], c[2], t);
    t = fma(pi[3], c[3], t);
    if (k == 0)
      term = p->rate_weights[0] * t;
    else
      term = fma(p->rate_weights[k], t, term);
  }
  return term;
}

double rdo_compute_root_loglikelihood_mode(rdo_partition_t *p,
                                           unsigned int clv_index,
                                           int scaler_index,
                                           const unsigned int *freqs_indices,
                                           double *persite_lnl, int mode) {
  const double   *clv = p->clv[clv_index];
  const unsigned *sb = scaler_index == RDO_SCALE_BUFFER_NONE
                           ? NULL
                           : p->scale_buffer[scaler_index];
  const double *pi = p->frequencies[freqs_indices ? freqs_indices[0] : 0];
  unsigned      S = p->sites;
  double       *tmp = NULL;
  if (mode == RDO_MODE_ENGINE) tmp = (double *)malloc(sizeof(double) * (S + 1));
  double logl = 0.0;
  for (unsigned s = 0; s < S; ++s) {
    double term = site_term(p, clv, pi, s);
    /* prop_invar is always 0 in RootDigger (Appendix B-5): no +I branch */
    double l = (mode == RDO_MODE_ENGINE) ? rdo_log(term) : log(term);
    if (sb) l = l + (double)sb[s] * RDO_LOG_SCALE_THRESHOLD;
    l = l * (double)p->pattern_weights[s];
    if (persite_lnl) persite_lnl[s] = l;
    if (tmp)
      tmp[s] = l;
    else
      logl = logl + l;
  }
  if (tmp) {
    logl = rdo_pairwise_sum(tmp, S);
    free(tmp);
  }
  return logl;
}

double rdo_compute_root_loglikelihood(rdo_partition_t *p,
                                      unsigned int clv_index, int scaler_index,
                                      const unsigned int *freqs_indices,
                                      double *persite_lnl) {
  return rdo_compute_root_loglikelihood_mode(p, clv_index, scaler_index,
                                             freqs_indices, persite_lnl,
                                             g_default_mode);
}

double rdo_compute_root_loglikelihood_mt(rdo_partition_t *p,
                                         unsigned int clv_index,
                                         int scaler_index, int threads) {
  const double   *clv = p->clv[clv_index];
  const unsigned *sb = scaler_index == RDO_SCALE_BUFFER_NONE
                           ? NULL
                           : p->scale_buffer[scaler_index];
  const double *pi = p->frequencies[0];
  long          S = (long)p->sites;
  double        logl = 0.0;
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) reduction(+ : logl) schedule(static)
  for (long s = 0; s < S; ++s) {
    double l = log(site_term(p, clv, pi, (unsigned)s));
    if (sb) l = l + (double)sb[s] * RDO_LOG_SCALE_THRESHOLD;
    logl += l * (double)p->pattern_weights[s];
  }
  return logl;
}

/* ------------------------------------------------------------------------ */
/* corax_msa_empirical_frequencies (src/model.cpp:329; Appendix A-6)         */
/* ------------------------------------------------------------------------ */
double *rdo_msa_empirical_frequencies(rdo_partition_t *p) {
  double  *f = (double *)calloc(4, sizeof(double));
  unsigned K = p->rate_cats;
  if (g_default_mode == RDO_MODE_ENGINE) {
    /* ENGINE arithmetic: exact integer histogram of the weighted state masks,
     * f_j = sum over masks m containing j (ascending m) of H[m]/popcount(m),
     * divided by the total count -- order independent, hence identical on any
     * number of shards */
    unsigned long long hist[16] = {0};
    for (unsigned t = 0; t < p->tips; ++t)
      for (unsigned s = 0; s < p->sites; ++s) {
        const double *c = p->clv[t] + ((size_t)s * K) * 4;
        unsigned      m = 0;
        for (int j = 0; j < 4; ++j)
          if (c[j] != 0.0) m |= 1u << j;
        hist[m] += p->pattern_weights[s];
      }
    double total = 0.0;
    for (int m = 1; m < 16; ++m) total += (double)hist[m];
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int m = 1; m < 16; ++m)
        if (m & (1 << j)) s += (double)hist[m] / (double)__builtin_popcount(m);
      f[j] = s / total;
    }
    return f;
  }
  double   wsum = 0.0;
  for (unsigned s = 0; s < p->sites; ++s) wsum += (double)p->pattern_weights[s];
  for (unsigned t = 0; t < p->tips; ++t)
    for (unsigned s = 0; s < p->sites; ++s) {
      const double *c = p->clv[t] + ((size_t)s * K) * 4;
      double        tot = ((c[0] + c[1]) + c[2]) + c[3];
      double        w = (double)p->pattern_weights[s];
      for (int j = 0; j < 4; ++j) f[j] += w * c[j] / tot;
    }
  double denom = wsum * (double)p->tips;
  for (int j = 0; j < 4; ++j) f[j] /= denom;
  return f;
}

/* ------------------------------------------------------------------------ */
/* corax_compute_gamma_cats (src/model.cpp:239-272; Appendix A-5): Yang 1994 */
/* discrete Gamma, with the PAML routines libpll uses (AS 91 / AS 239 / 245). */
/* ------------------------------------------------------------------------ */
static double ln_gamma(double x) {
  double f = 0.0, z;
  if (x < 7.0) {
    f = 1.0;
    z = x - 1.0;
    while (++z < 7.0) f *= z;
    x = z;
    f = -log(f);
  }
  z = 1.0 / (x * x);
  return f + (x - 0.5) * log(x) - x + .918938533204673 +
         (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) *
              z +
          .083333333333333) /
             x;
}

static double incomplete_gamma(double x, double alpha, double ln_gamma_alpha) {
  double p = alpha, g = ln_gamma_alpha;
  double accurate = 1e-8, overflow = 1e30;
  double factor, gin = 0, rn = 0, a = 0, b = 0, an = 0, dif = 0, term = 0;
  double pn[6];
  if (x == 0) return 0;
  if (x < 0 || p <= 0) return -1;
  factor = exp(p * log(x) - x - g);
  if (x > 1 && x >= p) {
    /* continued fraction */
    a = 1 - p;
    b = a + x + 1;
    term = 0;
    pn[0] = 1;
    pn[1] = x;
    pn[2] = x + 1;
    pn[3] = x * b;
    gin = pn[2] / pn[3];
    for (;;) {
      a++;
      b += 2;
      term++;
      an = a * term;
      for (int i = 0; i < 2; ++i) pn[i + 4] = b * pn[i + 2] - an * pn[i];
      if (pn[5] != 0) {
        rn = pn[4] / pn[5];
        dif = fabs(gin - rn);
        if (dif <= accurate && dif <= accurate * rn) break;
        gin = rn;
      }
      for (int i = 0; i < 4; ++i) pn[i] = pn[i + 2];
      if (fabs(pn[4]) >= overflow)
        for (int i = 0; i < 4; ++i) pn[i] /= overflow;
    }
    gin = 1 - factor * gin;
  } else {
    /* series expansion */
    gin = 1;
    term = 1;
    rn = p;
    do {
      rn++;
      term *= x / rn;
      gin += term;
    } while (term > accurate);
    gin *= factor / p;
  }
  return gin;
}

static double point_normal(double prob) {
  double a0 = -.322232431088, a1 = -1, a2 = -.342242088547,
         a3 = -.0204231210245, a4 = -.453642210148e-4, b0 = .0993484626060,
         b1 = .588581570495, b2 = .531103462366, b3 = .103537752850,
         b4 = .0038560700634;
  double y, z, p = prob, p1;
  p1 = (p < 0.5 ? p : 1 - p);
  if (p1 < 1e-20) return -9999;
  y = sqrt(log(1 / (p1 * p1)));
  z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) /
              ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
  return (p < 0.5 ? -z : z);
}

static double point_chi2(double prob, double v) {
  double e = .5e-6, aa = .6931471805, p = prob, g;
  double xx, c, ch, a = 0, q = 0, p1 = 0, p2 = 0, t = 0, x = 0, b = 0;
  double s1, s2, s3, s4, s5, s6;
  if (p < .000002 || p > .999998 || v <= 0) return -1;
  g = ln_gamma(v / 2);
  xx = v / 2;
  c = xx - 1;
  if (v < -1.24 * log(p)) {
    ch = pow((p * xx * exp(g + xx * aa)), 1 / xx);
    if (ch - e < 0) return ch;
  } else if (v <= .32) {
    ch = 0.4;
    a = log(1 - p);
    do {
      q = ch;
      p1 = 1 + ch * (4.67 + ch);
      p2 = ch * (6.73 + ch * (6.66 + ch));
      t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
      ch -= (1 - exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
    } while (fabs(q / ch - 1) - .01 > 0);
  } else {
    x = point_normal(p);
    p1 = 0.222222 / v;
    ch = v * pow((x * sqrt(p1) + 1 - p1), 3.0);
    if (ch > 2.2 * v + 6) ch = -2 * log(1 - p) - c * log(.5 * ch) + g;
  }
  do {
    q = ch;
    p1 = .5 * ch;
    if ((t = incomplete_gamma(p1, xx, g)) < 0) return -1;
    p2 = p - t;
    t = p2 * exp(xx * aa + g + p1 - c * log(ch));
    b = t / ch;
    a = 0.5 * t - b * c;
    s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
    s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
    s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
    s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
    s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
    s6 = (120 + c * (346 + 127 * c)) / 5040;
    ch += t * (1 + 0.5 * t * s1 -
               b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
  } while (fabs(q / ch - 1) > e);
  return ch;
}

#define POINT_GAMMA(prob, alpha, beta) (point_chi2(prob, 2.0 * (alpha)) / (2.0 * (beta)))

int rdo_compute_gamma_cats(double alpha, unsigned int categories,
                           double *output_rates, int rates_mode) {
  if (alpha < 0.02) {
    set_error("Invalid alpha value (must be >= 0.02)");
    return RDO_FAILURE;
  }
  if (categories == 0) {
    set_error("Number of categories must be positive");
    return RDO_FAILURE;
  }
  if (categories == 1) {
    output_rates[0] = 1.0;
    return RDO_SUCCESS;
  }
  double alfa = alpha, beta = alpha;
  double factor = alfa / beta * categories;
  if (rates_mode == RDO_GAMMA_RATES_MEDIAN) {
    double middle = 1.0 / (2.0 * categories), t = 0.0;
    for (unsigned i = 0; i < categories; ++i)
      output_rates[i] = POINT_GAMMA((double)(i * 2 + 1) * middle, alfa, beta);
    for (unsigned i = 0; i < categories; ++i) t += output_rates[i];
    for (unsigned i = 0; i < categories; ++i)
      output_rates[i] /= (t / (double)categories);
  } else if (rates_mode == RDO_GAMMA_RATES_MEAN) {
    double *gp = (double *)malloc(sizeof(double) * categories);
    double  lnga1 = ln_gamma(alfa + 1);
    for (unsigned i = 0; i < categories - 1; ++i)
      gp[i] = POINT_GAMMA((i + 1.0) / categories, alfa, beta);
    for (unsigned i = 0; i < categories - 1; ++i)
      gp[i] = incomplete_gamma(gp[i] * beta, alfa + 1, lnga1);
    output_rates[0] = gp[0] * factor;
    output_rates[categories - 1] = (1 - gp[categories - 2]) * factor;
    for (unsigned i = 1; i < categories - 1; ++i)
      output_rates[i] = (gp[i] - gp[i - 1]) * factor;
    free(gp);
  } else {
    set_error("Unknown gamma rates mode");
    return RDO_FAILURE;
  }
  return RDO_SUCCESS;
}
