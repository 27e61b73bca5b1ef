// rdk_kernels.cuh -- sm_100a kernels of the RootDigger likelihood engine.
//
// Replaces the arithmetic behind corax_update_prob_matrices,
// corax_update_clvs and corax_compute_root_loglikelihood (call sites:
// reference src/model.cpp:367,432,842 / :402,440,461,851 / :406,441,466).
//
// Arithmetic contract (DESIGN.md "Arithmetic specification", v2): fp64; the 4x4
// mat-vec of the CLV update and the dot products of the root log-likelihood are
// explicit FMA chains (__fma_rn, j ascending); every other +,-,*,/ is
// individually rounded (round-to-nearest-even, no contraction), in the order
// written.  All arithmetic goes through the __d*_rn / __fma_rn intrinsics so the
// contract holds whatever -fmad says.
//
// Memory-bound by design: 0.61 flop/B, no tensor cores (a 4x4 mat-vec is not a
// dense contraction).  What matters here is coalesced 32-byte-per-thread
// accesses, enough bytes in flight, and never re-reading a CLV from HBM.
#pragma once
#include <cuda_runtime.h>
#ifndef RDK_KSLOW
// 1: CLVs are stored in blocks of 32 elements (= 32/K sites) laid out [cat][site in block][state]
//    and lane l of a warp holds category l / (32/K): global accesses stay 1 KiB-contiguous per warp
//    while the lanes of a quarter warp share their category (one shared-memory address per
//    quarter warp when reading P).  0: natural [site][cat][state] layout, category = l % K.
#define RDK_KSLOW 0  // measured on B200: no gain from the blocked layout (the LSU pipe is not the limiter)
#endif
#ifndef RDK_MINB2
#define RDK_MINB2 4  // resident CTAs of 128 threads the E = 2 program kernel is compiled for
#endif
#ifndef RDK_FAST_KINDS
// 1: per-kind compile-time copies of the instruction body (K = 4; E = 2, 4).  Measured on B200
// (cfg2 step): 24 % fewer warp instructions, but the same time at E = 4 (80.4 k vs 81-82 k
// placements/s) and +4 % at E = 2 (66.4 k vs 63.8 k): the walk is bound by the latency of each
// warp's dependent chain, not by issue slots, and the copies cost instruction-cache misses and
// 2 more minutes of compile time.  (2: also for the short tail passes -- 64.7 k at 100 k sites,
// no gain on a 12.5 k-site shard.)  Off by default.
#define RDK_FAST_KINDS 0
#endif
#ifndef RDK_TABLES_L1
// 1: the P / tip tables of an instruction are read straight from global memory through L1
//    (ld.global.nc; the lines are prefetched into L1 one instruction ahead) -- no shared-memory
//    staging, no mbarrier wait, no per-warp copy issue per instruction.
// 0: every warp stages them in its own shared-memory double buffer with cp.async.bulk + mbarrier.
// Measured on B200 (cfg2 step, E = 4): 63.7 k placements/s through L1 against 81-82 k with the
// shared-memory staging -- the L1 hit latency sits on every instruction's critical path.
#define RDK_TABLES_L1 0
#endif
#ifndef RDK_FWD_STATIC
// 1: the instruction body is compiled twice, for "the next instruction forwards my values" (they
//    are produced directly in its child-2 operand registers) and for "it does not".
// 0: one copy; forwarded values are moved into the next instruction's operand registers under a
//    run-time test (8 E register moves), which halves the code the warps of an SM walk through.
#define RDK_FWD_STATIC 1
#endif
#ifndef RDK_L2_PREFETCH_DIST
// > 0: while instruction i computes, the CLV operands of instruction i + DIST are pulled into
// L2 (prefetch.global.L2, no registers), so that the register loads issued one instruction
// ahead find them there instead of paying the DRAM latency.  Measured on B200 (cfg2 step and
// its 12.5 k-site shard, distances 2 / 3 / 5): no change -- the walk is not waiting on DRAM.
#define RDK_L2_PREFETCH_DIST 0
#endif
// timing experiments only (tools/build_variant.sh): each removes one piece of the instruction
// body -- the results are WRONG when any is set
#ifndef RDK_X_NOEVAL
#define RDK_X_NOEVAL 0
#endif
#ifndef RDK_X_NOWAIT
#define RDK_X_NOWAIT 0
#endif
#ifndef RDK_X_NOLOAD
#define RDK_X_NOLOAD 0
#endif
#ifndef RDK_X_NOSTORE
#define RDK_X_NOSTORE 0
#endif
#ifndef RDK_LD256
#define RDK_LD256 1  // 256-bit global loads/stores of CLV elements
#endif
#include <stdint.h>
#include <type_traits>

namespace rdk {

constexpr int kMaxCats = 32;  // rate categories handled by the fast path: K | 32

// ---------------------------------------------------------------------------
// individually rounded arithmetic
// ---------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
// the 4x4 mat-vec of the CLV update and the root dot products are explicit FMA chains
// (arithmetic spec v2); everything else stays individually rounded
__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// 2^-256 underflow threshold, 2^256 rescale factor, ln(2^-256)
#define RDK_SCALE_THRESHOLD 0x1p-256
#define RDK_SCALE_FACTOR 0x1p+256
#define RDK_LOG_SCALE_THRESHOLD (-177.44567822334599)

// ---------------------------------------------------------------------------
// rd_log: software natural logarithm.  Reduction of x to 2^k (1+f) with
// sqrt(2)/2 <= 1+f < sqrt(2), then the degree-14 minimax in s = f/(2+f)
// (fdlibm / musl formulation), < 1 ulp.  Written with integer bit operations
// and individually rounded fp64 ops only, so that it produces the same bits on
// any IEEE-754 machine; libdevice's log() is NOT used because its result may
// differ from a host libm in the last bit, which would make root-branch
// derivatives (difference quotients with h = 1e-8, reference
// src/model.cpp:481-519) irreproducible.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double rd_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  uint64_t ix = (uint64_t)__double_as_longlong(x);
  uint32_t hx = (uint32_t)(ix >> 32);
  int      k = 0;
  if (hx < 0x00100000u || (hx >> 31)) {
    if ((ix << 1) == 0) return __longlong_as_double(0xfff0000000000000LL);  // -inf
    if (hx >> 31) return __longlong_as_double(0x7ff8000000000000LL);        // nan
    k -= 54;
    x = dmul(x, 0x1p54);
    ix = (uint64_t)__double_as_longlong(x);
    hx = (uint32_t)(ix >> 32);
  } else if (hx >= 0x7ff00000u) {
    return x;
  } else if (hx == 0x3ff00000u && (ix << 32) == 0) {
    return 0.0;
  }
  hx += 0x3ff00000u - 0x3fe6a09eu;
  k += (int)(hx >> 20) - 0x3ff;
  hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
  ix = ((uint64_t)hx << 32) | (ix & 0xffffffffULL);
  x = __longlong_as_double((long long)ix);

  double f = dsub(x, 1.0);
  double hfsq = dmul(dmul(0.5, f), f);
  double s = ddiv(f, dadd(2.0, f));
  double z = dmul(s, s);
  double w = dmul(z, z);
  double t1 = dmul(w, dadd(Lg2, dmul(w, dadd(Lg4, dmul(w, Lg6)))));
  double t2 = dmul(z, dadd(Lg1, dmul(w, dadd(Lg3, dmul(w, dadd(Lg5, dmul(w, Lg7)))))));
  double R = dadd(t2, t1);
  double dk = (double)k;
  double r = dmul(s, dadd(hfsq, R));
  r = dadd(r, dmul(dk, ln2_lo));
  r = dsub(r, hfsq);
  r = dadd(r, f);
  r = dadd(r, dmul(dk, ln2_hi));
  return r;
}

// ---------------------------------------------------------------------------
// expm4: exp of a 4x4 matrix, Higham (2005) scaling and squaring with Pade
// approximants of degree 3/5/7/9/13.  One thread per matrix.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mm4(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = dmul(A[i * 4 + 0], B[0 * 4 + j]);
      s = dadd(s, dmul(A[i * 4 + 1], B[1 * 4 + j]));
      s = dadd(s, dmul(A[i * 4 + 2], B[2 * 4 + j]));
      s = dadd(s, dmul(A[i * 4 + 3], B[3 * 4 + j]));
      C[i * 4 + j] = s;
    }
}

__device__ __forceinline__ double norm1_4(const double* A) {
  double best = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double s = fabs(A[0 * 4 + j]);
    s = dadd(s, fabs(A[1 * 4 + j]));
    s = dadd(s, fabs(A[2 * 4 + j]));
    s = dadd(s, fabs(A[3 * 4 + j]));
    if (s > best) best = s;
  }
  return best;
}

// X = M^-1 N, Gaussian elimination with partial pivoting (first maximal pivot)
__device__ inline void solve4(double* M, double* N, double* X) {
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
      double f = ddiv(M[r * 4 + c], M[c * 4 + c]);
      for (int j = c + 1; j < 4; ++j) M[r * 4 + j] = dsub(M[r * 4 + j], dmul(f, M[c * 4 + j]));
      for (int j = 0; j < 4; ++j) N[r * 4 + j] = dsub(N[r * 4 + j], dmul(f, N[c * 4 + j]));
    }
  }
  for (int r = 3; r >= 0; --r)
    for (int j = 0; j < 4; ++j) {
      double s = N[r * 4 + j];
      for (int q = r + 1; q < 4; ++q) s = dsub(s, dmul(M[r * 4 + q], X[q * 4 + j]));
      X[r * 4 + j] = ddiv(s, M[r * 4 + r]);
    }
}

__device__ inline void expm4(const double* Ain, double* E) {
  const double TH3 = 1.495585217958292e-2, TH5 = 2.539398330063230e-1,
               TH7 = 9.504178996162932e-1, TH9 = 2.097847961257068e0,
               TH13 = 5.371920351148152e0;
  double A[16], U[16], V[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] = Ain[i];
  double n1 = norm1_4(A);
  int    s = 0;
  int    deg = 13;
  if (n1 <= TH3)
    deg = 3;
  else if (n1 <= TH5)
    deg = 5;
  else if (n1 <= TH7)
    deg = 7;
  else if (n1 <= TH9)
    deg = 9;
  if (deg == 13) {
    while (n1 > TH13) {
      n1 = dmul(n1, 0.5);
      ++s;
    }
    // 2^-s, exact
    double sc = __longlong_as_double((long long)(uint64_t)(1023 - s) << 52);
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] = dmul(A[i], sc);
  }
  double A2[16], A4[16], A6[16];
  mm4(A, A, A2);
  if (deg == 13) {
    const double b0 = 64764752532480000., b1 = 32382376266240000., b2 = 7771770303897600.,
                 b3 = 1187353796428800., b4 = 129060195264000., b5 = 10559470521600.,
                 b6 = 670442572800., b7 = 33522128640., b8 = 1323241920., b9 = 40840800.,
                 b10 = 960960., b11 = 16380., b12 = 182., b13 = 1.;
    double W1[16], Z1[16], W[16], T[16];
    mm4(A2, A2, A4);
    mm4(A4, A2, A6);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double t = dmul(b13, A6[i]);
      t = dadd(t, dmul(b11, A4[i]));
      t = dadd(t, dmul(b9, A2[i]));
      W1[i] = t;
      t = dmul(b12, A6[i]);
      t = dadd(t, dmul(b10, A4[i]));
      t = dadd(t, dmul(b8, A2[i]));
      Z1[i] = t;
    }
    mm4(A6, W1, T);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double id = (i % 5 == 0) ? 1.0 : 0.0;
      double t = dmul(b7, A6[i]);
      t = dadd(t, dmul(b5, A4[i]));
      t = dadd(t, dmul(b3, A2[i]));
      t = dadd(t, dmul(b1, id));
      W[i] = dadd(T[i], t);
    }
    mm4(A, W, U);
    mm4(A6, Z1, T);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      double id = (i % 5 == 0) ? 1.0 : 0.0;
      double t = dmul(b6, A6[i]);
      t = dadd(t, dmul(b4, A4[i]));
      t = dadd(t, dmul(b2, A2[i]));
      t = dadd(t, dmul(b0, id));
      V[i] = dadd(T[i], t);
    }
  } else {
    double A8[16], W[16];
    if (deg >= 5) mm4(A2, A2, A4);
    if (deg >= 7) mm4(A4, A2, A6);
    if (deg >= 9) mm4(A6, A2, A8);
    for (int i = 0; i < 16; ++i) {
      double id = (i % 5 == 0) ? 1.0 : 0.0;
      double w, v;
      if (deg == 3) {
        w = dmul(1., A2[i]);
        v = dmul(12., A2[i]);
        W[i] = dadd(w, dmul(60., id));
        V[i] = dadd(v, dmul(120., id));
      } else if (deg == 5) {
        w = dmul(1., A4[i]);
        w = dadd(w, dmul(420., A2[i]));
        v = dmul(30., A4[i]);
        v = dadd(v, dmul(3360., A2[i]));
        W[i] = dadd(w, dmul(15120., id));
        V[i] = dadd(v, dmul(30240., id));
      } else if (deg == 7) {
        w = dmul(1., A6[i]);
        w = dadd(w, dmul(1512., A4[i]));
        w = dadd(w, dmul(277200., A2[i]));
        v = dmul(56., A6[i]);
        v = dadd(v, dmul(25200., A4[i]));
        v = dadd(v, dmul(1995840., A2[i]));
        W[i] = dadd(w, dmul(8648640., id));
        V[i] = dadd(v, dmul(17297280., id));
      } else {
        w = dmul(1., A8[i]);
        w = dadd(w, dmul(3960., A6[i]));
        w = dadd(w, dmul(2162160., A4[i]));
        w = dadd(w, dmul(302702400., A2[i]));
        v = dmul(90., A8[i]);
        v = dadd(v, dmul(110880., A6[i]));
        v = dadd(v, dmul(30270240., A4[i]));
        v = dadd(v, dmul(2075673600., A2[i]));
        W[i] = dadd(w, dmul(8821612800., id));
        V[i] = dadd(v, dmul(17643225600., id));
      }
    }
    mm4(A, W, U);
  }
  double M[16], N[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    M[i] = dsub(V[i], U[i]);
    N[i] = dadd(V[i], U[i]);
  }
  solve4(M, N, E);
  for (int q = 0; q < s; ++q) {
    double T[16];
    mm4(E, E, T);
#pragma unroll
    for (int i = 0; i < 16; ++i) E[i] = T[i];
  }
}

// Q_ij = r_(ij) pi_j (12 r's row-major off-diagonal), diagonal = -row sum,
// normalised to unit mean rate -sum_i pi_i Q_ii = 1.
__device__ inline void build_q_nonrev(const double* r, const double* pi, double* Q) {
  int k = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (i != j) Q[i * 4 + j] = dmul(r[k++], pi[j]);
  for (int i = 0; i < 4; ++i) {
    double s = 0.0;
    bool   first = true;
    for (int j = 0; j < 4; ++j) {
      if (j == i) continue;
      if (first) {
        s = Q[i * 4 + j];
        first = false;
      } else
        s = dadd(s, Q[i * 4 + j]);
    }
    Q[i * 4 + i] = -s;
  }
  double mu = dmul(pi[0], -Q[0]);
  mu = dadd(mu, dmul(pi[1], -Q[5]));
  mu = dadd(mu, dmul(pi[2], -Q[10]));
  mu = dadd(mu, dmul(pi[3], -Q[15]));
  for (int i = 0; i < 16; ++i) Q[i] = ddiv(Q[i], mu);
}

// ---------------------------------------------------------------------------
// Tip states are stored on the device as 4-bit CODES, a permutation of the
// 4-bit state masks chosen so that the unambiguous states A,C,G,T get codes
// 0..3 (they then hit distinct shared-memory banks in the tip tables).
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned tip_mask_of_code(unsigned code) {
  // code: 0  1  2  3  4  5  6  7  8  9  10 11 12 13 14 15
  // mask: 1  2  4  8  3  5  6  7  9  10 11 12 13 14 15 0
  return (unsigned)((0x0FEDCBA976538421ULL >> (4 * code)) & 15ULL);
}
__host__ __device__ __forceinline__ unsigned tip_code_of_mask(unsigned mask) {
  // mask: 0  1  2  3  4  5  6  7  8  9  10 11 12 13 14 15
  // code: 15 0  1  4  2  5  6  7  3  8  9  10 11 12 13 14
  return (unsigned)((0xEDCBA9837652410FULL >> (4 * mask)) & 15ULL);
}

// ---------------------------------------------------------------------------
// pmat_expm_nonrev: one thread per (branch entry, rate category).
// Replaces corax_update_prob_matrices.  A pool slot holds, for one branch,
//   P : [cat][i*4+j (+2 pad)]   18*K doubles  (the transition matrices; the pad
//                               makes a category's row start 144 B apart so that the
//                               128-bit shared-memory reads of a warp's K categories
//                               fall into distinct banks)
//   T : [tip code][cat][i]      64*K doubles  (tip lookup table)
// T[code][k][i] = ((P_i0 c_0 + P_i1 c_1) + P_i2 c_2) + P_i3 c_3 with c the 0/1
// vector of the tip state: exactly the expression a tip child contributes to a
// CLV update, evaluated once per branch instead of once per site.
// The slot layout IS the shared-memory layout: the program kernel moves P or T
// with one bulk async copy.
// ---------------------------------------------------------------------------
constexpr int kPTabDoubles = 18;    // per category
constexpr int kTipTabDoubles = 64;  // per category
constexpr int kTabDoubles = 64;     // per category: the larger of the two
constexpr int kSlotDoubles = kPTabDoubles + kTipTabDoubles;  // per category

struct PmatEntry {
  unsigned slot;  // physical slot in the P-matrix pool
  unsigned pad;
  double   t;  // branch length
};

constexpr int kPmatInline = 16;
struct PmatArgs {
  const PmatEntry* entries;  // used when n > kPmatInline
  int              n;
  int              K;
  double           r[12];
  double           pi[4];
  double           pinv;
  double           rates[kMaxCats];
  double*          pool;
  PmatEntry        inl[kPmatInline];
};

#ifndef RDK_PROGRAM_KERNEL_ONLY
__global__ void __launch_bounds__(64) pmat_expm_nonrev_kernel(const __grid_constant__ PmatArgs a) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= a.n * a.K) return;
  int       e = tid / a.K, k = tid - e * a.K;
  PmatEntry ent = (a.n <= kPmatInline) ? a.inl[e] : a.entries[e];
  double    Q[16], A[16], E[16];
  double    r[12], pi[4];
  for (int i = 0; i < 12; ++i) r[i] = a.r[i];
  for (int i = 0; i < 4; ++i) pi[i] = a.pi[i];
  build_q_nonrev(r, pi, Q);
  double c = ddiv(dmul(a.rates[k], ent.t), dsub(1.0, a.pinv));
  for (int i = 0; i < 16; ++i) A[i] = dmul(Q[i], c);
  expm4(A, E);
  double* slot = a.pool + (size_t)ent.slot * a.K * kSlotDoubles;
  double* out = slot + (size_t)k * kPTabDoubles;
  for (int i = 0; i < 16; ++i) out[i] = E[i];
  out[16] = out[17] = 0.0;
  double* tab = slot + (size_t)a.K * kPTabDoubles;
  for (int i = 0; i < 4; ++i)
    for (unsigned code = 0; code < 16; ++code) {
      unsigned m = tip_mask_of_code(code);
      double   x = dmul(E[i * 4 + 0], (m & 1u) ? 1.0 : 0.0);
      x = dadd(x, dmul(E[i * 4 + 1], (m & 2u) ? 1.0 : 0.0));
      x = dadd(x, dmul(E[i * 4 + 2], (m & 4u) ? 1.0 : 0.0));
      x = dadd(x, dmul(E[i * 4 + 3], (m & 8u) ? 1.0 : 0.0));
      tab[((size_t)code * a.K + k) * 4 + i] = x;
    }
}

#endif  // RDK_PROGRAM_KERNEL_ONLY

// ---------------------------------------------------------------------------
// The likelihood program kernel.
//
// A "program" is a list of instructions executed in order for every
// (site, category) element.  Elements are independent of each other (the
// dependency between a parent CLV and its children is per site), so each warp
// owns a contiguous range of elements and walks the WHOLE program on it with
// no grid synchronisation: a full post-order traversal (n-1 CLV
// operations + the root log-likelihood), a root move, or an entire placement
// sweep is ONE launch.  Child CLVs produced a few instructions earlier by the
// same warp are still in L2 (or in registers), so HBM sees each CLV written
// once and mostly not read back.
//
// Thread mapping: element e = site*K + k.  A warp iteration `it` covers the 32/K
// sites [it*32/K, (it+1)*32/K) = 32 consecutive 32-byte (4 x fp64) vectors = 1 KiB
// contiguous per CLV per warp access.  Lane l handles site it*32/K + l % (32/K),
// category k = l / (32/K): the lanes of a quarter warp share k, so their 128-bit
// shared-memory reads of P hit ONE address per quarter warp (measured on B200:
// 2 LSU cycles per read instead of 4 for the interleaved k = l % K mapping).
// ---------------------------------------------------------------------------
enum : unsigned {
  kTip1 = 1u,      // child1 is a tip: 1 byte per site (4-bit state code)
  kTip2 = 2u,      // child2 is a tip
  kWrite = 4u,     // store the parent CLV (and parent scaler if present)
  kEval = 8u,      // evaluate the root log-likelihood of the parent values
  kLoadOnly = 16u, // no CLV arithmetic: parent values := CLV at c1 (root logL of
                   // a stored CLV, corax_compute_root_loglikelihood)
  kScale = 32u,    // parent has a scale buffer: apply 2^256 rescaling
  // pre-decoded by the host when the program is finalised (finalize_program):
  kFwd1 = 64u,     // c1 is the CLV the previous instruction produced: take it from registers
  kFwd2 = 128u,    // same for c2
  kLdS1 = 256u,    // load child1's scaler counts from memory
  kLdS2 = 512u,    // load child2's scaler counts from memory
  kFwdS1 = 1024u,  // child1's scaler is the one the previous instruction produced
  kFwdS2 = 2048u,  // same for child2
  kEvalScaler = 4096u,  // the evaluated root has a scale buffer (adds cnt * ln 2^-256)
};

// The flags that select code in the arithmetic of an instruction (the rest only steer the
// operand loads).  For the combinations below -- the ones a post-order traversal and the
// directed placement sweep are made of -- the K = 4, E = 2 kernel carries copies of the
// instruction body compiled with the flags as CONSTANTS (no flag tests), once for "the next
// instruction takes my result as its child 2" (the values are then produced directly in the
// registers of the next instruction's operand: forwarding costs no register move) and once
// for "it does not".  The host stores 2 * (index + 1) + forwards-out in Instr::kind
// (finalize_program; 0 / 1 = flags decoded at run time), after ordering the two children
// canonically -- a tip first, a forwarded CLV last -- which the commutative product
// (P1 c1) o (P2 c2) allows without changing a bit of the result.
constexpr unsigned kKindMask = kTip1 | kTip2 | kWrite | kEval | kLoadOnly | kScale | kFwd1 | kEvalScaler;
constexpr unsigned kFastKinds[] = {
    kWrite | kScale,                         // inner x inner
    kWrite | kScale | kTip1,                 // tip x inner
    kWrite | kScale | kTip1 | kTip2,         // tip x tip
    kEval | kEvalScaler | kScale,            // sweep placement: evaluated in registers, nothing stored
    kEval | kEvalScaler | kScale | kTip1,    // sweep placement on a tip branch
};
constexpr int kNumFastKinds = (int)(sizeof(kFastKinds) / sizeof(kFastKinds[0]));
inline constexpr unsigned fast_kind_of(unsigned flags) {
  for (int i = 0; i < kNumFastKinds; ++i)
    if ((flags & kKindMask) == kFastKinds[i]) return (unsigned)i + 1u;
  return 0u;
}

struct alignas(16) Instr {
  double*         parent;
  const void*     c1;
  const void*     c2;
  unsigned*       pscale;
  const unsigned* c1scale;
  const unsigned* c2scale;
  const double*   P1;  // child1's table in its branch's pool slot: P (inner child) or T (tip child)
  const double*   P2;
  unsigned        flags;
  unsigned        slot;  // eval slot (row of the partial-sum buffer)
  unsigned        kind;  // 2 * (index into kFastKinds + 1, or 0) + (the next instruction forwards my result)
  unsigned        tx;    // bytes of the two tables: b1 | b2 << 16 (0: no tables, kLoadOnly)
};
static_assert(sizeof(Instr) == 80, "Instr layout");

constexpr int kProgInline = 8;
#ifndef RDK_PROG_WINDOW
#define RDK_PROG_WINDOW 256
#endif
constexpr int kProgWindow = RDK_PROG_WINDOW;  // instructions staged in shared memory at a time
constexpr int kMaxChunks = 16;  // independent sub-programs one launch can run side by side
struct ProgArgs {
  const Instr*    prog;  // used when n_instr > kProgInline
  int             n_instr;
  unsigned        nelem;    // sites * K on this shard
  unsigned        n_witer;  // ceil(nelem / 32)
  const unsigned* weights;  // pattern weights [sites]
  double*         partials; // [slots][partial_stride], one value per warp iteration
  unsigned        partial_stride;
  double*         persite;  // optional, eval slot 0 only
  double          pi[4];
  double          w[kMaxCats];
  // n_chunks > 1: the program is n_chunks INDEPENDENT sub-programs (chunk c = instructions
  // [chunk_off[c], chunk_off[c+1]) of prog); blockIdx.y selects the one a CTA walks, so a
  // small shard fills the device with (site range) x (chunk) warps instead of leaving a
  // handful of warps per SM to walk one long chain
  unsigned        n_chunks;
  unsigned        chunk_off[kMaxChunks + 1];
  Instr           inl[kProgInline];
};

struct d4 {
  double v[4];
};

__device__ __forceinline__ d4 ld_clv(const double* base, unsigned e) {
  // one 256-bit load per element (sm_100 LDG.E.256): a warp access covers 1 KiB
  // contiguous with every 32-byte sector used by exactly one lane.  L2-coherent
  // (.cg): CLVs are read and written within one launch and never re-read by the
  // same SM soon enough for L1 to help.
  const char* q = reinterpret_cast<const char*>(base) + (size_t)e * 32u;
  d4          r;
#if RDK_LD256
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];\n"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
               : "l"(q));
#else
  const double2* q2 = reinterpret_cast<const double2*>(q);
  double2        a = __ldcg(q2), b = __ldcg(q2 + 1);
  r.v[0] = a.x;
  r.v[1] = a.y;
  r.v[2] = b.x;
  r.v[3] = b.y;
#endif
  return r;
}
__device__ __forceinline__ void st_clv(double* base, unsigned e, const d4& x) {
  char* q = reinterpret_cast<char*>(base) + (size_t)e * 32u;
#if RDK_LD256
  asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};\n" ::"l"(q), "d"(x.v[0]), "d"(x.v[1]), "d"(x.v[2]),
               "d"(x.v[3])
               : "memory");
#else
  double2* q2 = reinterpret_cast<double2*>(q);
  __stcg(q2, make_double2(x.v[0], x.v[1]));
  __stcg(q2 + 1, make_double2(x.v[2], x.v[3]));
#endif
}

// a 16-byte read of a P / tip table entry
__device__ __forceinline__ double2 ld_tab(const double2* p) {
#if RDK_TABLES_L1
  return __ldg(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p));
}

// ---- mbarrier + bulk async copy (one elected thread moves a whole table) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// operands of one instruction for the E elements of a thread
template <int E>
struct Operands {
  d4       c1[E], c2[E];  // inner children: the 4 state likelihoods
  unsigned m1[E], m2[E];  // tip children: the state code
  unsigned cnt1[E], cnt2[E];  // the children's scaler counts (kept apart so that nothing
                              // waits on the loads before the instruction that needs them)
};

// Each WARP owns a contiguous range of warp iterations and walks the whole
// program over it, one "pass" of E iterations at a time, independently of every
// other warp (no CTA barrier per instruction).
//  * the program is staged in shared memory in windows (the only CTA-level
//    synchronisation: once per kProgWindow instructions, and only when the
//    program does not fit one window);
//  * per instruction and child, either the transition matrices P (inner child)
//    or the tip table T (tip child) of the child's branch is moved into the
//    warp's own shared double buffer by lane 0 with a bulk async copy
//    (cp.async.bulk + mbarrier) while the previous instruction computes;
//  * the global operands of instruction i+1 are loaded into registers BEFORE
//    the arithmetic of instruction i (software pipelining, two register sets
//    used in ping-pong), and when a child of i+1 is the CLV instruction i is
//    producing -- the normal case in a post-order schedule -- it is forwarded in
//    registers and never re-read (the host pre-decodes this into kFwd*);
//  * there are no per-lane validity predicates: a lane without a site of its
//    own (tail of the partition) redundantly recomputes the last site and
//    stores the identical values (same warp, same instruction: no race); only
//    the log-likelihood reduction masks it out.  The last pass of a warp whose
//    range is not a multiple of E runs the instruction loop instantiated for
//    the number of iterations it has left (NV < E) when the kernel is
//    instantiated with TS (tail skip); without TS the slots without an iteration
//    of their own redundantly recompute the last iteration of the range.  The
//    host picks TS when no warp has a full last pass (choose_tail_skip): there
//    the short loop pays; otherwise the slowest warps run full passes anyway and
//    the second copy of the loop only costs instruction-cache misses (measured).
template <int K, int E, int MAXT, int MINB, bool TS>
__global__ void __launch_bounds__(MAXT, MINB) clv_program_kernel(const __grid_constant__ ProgArgs a) {
  static_assert(32 % K == 0, "K must divide the warp size");
  // the configuration RootDigger runs DNA data in (4 Gamma categories) carries the
  // per-kind copies of the instruction body; the others decode the flags at run time
  constexpr bool FAST = (K == 4 && (E == 2 || E == 4)) && RDK_FAST_KINDS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr unsigned kTabBytes = kTabDoubles * K * 8;  // one child's slice of a table buffer
  Instr*              s_prog = reinterpret_cast<Instr*>(smem_raw);
  const unsigned      wib = threadIdx.x >> 5;
  const unsigned      wpb = blockDim.x >> 5;
  // per warp: [buf][child][kTabBytes] tables, then (after all warps' tables) [warp][buf] barriers
  unsigned char*      s_tab = smem_raw + sizeof(Instr) * kProgWindow + (size_t)wib * 4 * kTabBytes;
  unsigned long long* s_bar =
      reinterpret_cast<unsigned long long*>(smem_raw + sizeof(Instr) * kProgWindow + (size_t)wpb * 4 * kTabBytes) +
      2 * wib;

  const unsigned tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  constexpr unsigned SPW = 32 / K;  // sites per warp iteration
#if RDK_KSLOW
  constexpr unsigned KSTRIDE = SPW;  // lane distance between two categories of a site
  const unsigned     k = lane / SPW;
  const unsigned     sl = lane % SPW;  // site within the warp iteration
#else
  constexpr unsigned KSTRIDE = 1;
  const unsigned     k = lane % K;
  const unsigned     sl = lane / K;
#endif
  const unsigned lane0 = lane - k * KSTRIDE;  // the k == 0 lane of this lane's site
  unsigned       gmask = 0;                   // the K lanes that hold this lane's site
#pragma unroll
  for (int j = 0; j < K; ++j) gmask |= 1u << (j * KSTRIDE + lane0);

  const unsigned gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
  const unsigned it_begin = (unsigned)(((unsigned long long)a.n_witer * gw) / nw);
  const unsigned it_end = (unsigned)(((unsigned long long)a.n_witer * (gw + 1)) / nw);
  // every warp of the grid runs the same number of passes (the window barriers below are
  // CTA-wide); a warp whose range is exhausted idles through the remaining ones
  const unsigned passes = ((a.n_witer + nw - 1) / nw + E - 1) / E;
  const unsigned last_site = a.nelem / K - 1;
  const Instr*   prog = a.prog;
  int            n_instr = a.n_instr;
  if (a.n_chunks > 1) {
    prog += a.chunk_off[blockIdx.y];
    n_instr = (int)(a.chunk_off[blockIdx.y + 1] - a.chunk_off[blockIdx.y]);
  }
  const bool     multi_window = n_instr > kProgWindow;

#if !RDK_TABLES_L1
  if (lane == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
#endif
  unsigned phase = 0;  // bit b: parity of the next completion of s_bar[b]

#if RDK_TABLES_L1
  // the whole warp: pull the lines of P or T of both children of `in` into L1 (lane l
  // takes the l-th 128-byte line of each table; a tip table is 16 lines at K = 4)
  auto prefetch_tables = [&](const Instr& in, unsigned) {
    const unsigned tx = in.tx;
    if (tx == 0) return;
    const unsigned       b1 = tx & 0xffffu, b2 = tx >> 16;
    const unsigned char* s1 = reinterpret_cast<const unsigned char*>(in.P1);
    const unsigned char* s2 = reinterpret_cast<const unsigned char*>(in.P2);
    for (unsigned off = lane * 128u; off < b1 + 127u; off += 32u * 128u) prefetch_l1(s1 + min(off, b1 - 1u));
    for (unsigned off = lane * 128u; off < b2 + 127u; off += 32u * 128u) prefetch_l1(s2 + min(off, b2 - 1u));
  };
#else
  // one thread: move P or T of both children of `in` into buffer `buf`
  auto prefetch_tables = [&](const Instr& in, unsigned buf) {
    unsigned long long* bar = &s_bar[buf];
    const unsigned      tx = in.tx;  // sizes and table addresses are pre-computed by the host
    if (tx == 0) {
      mbar_arrive(bar);
      return;
    }
    const unsigned b1 = tx & 0xffffu, b2 = tx >> 16;
    mbar_expect_tx(bar, b1 + b2);
    unsigned char* dst = s_tab + (size_t)buf * 2 * kTabBytes;
    bulk_g2s(dst, in.P1, b1, bar);
    bulk_g2s(dst + kTabBytes, in.P2, b2, bar);
  };
#endif

  for (unsigned pass = 0; pass < passes; ++pass) {
    const bool active = it_begin + pass * E < it_end;  // warp-uniform
    unsigned   it[E], e[E], site[E];
#pragma unroll
    for (int u = 0; u < E; ++u) {
      const unsigned i0 = it_begin + pass * E + u;
      it[u] = (i0 < it_end || !active) ? i0 : it_end - 1;
      unsigned st = it[u] * SPW + sl;
      if (st > last_site) st = last_site;
      site[u] = st;
#if RDK_KSLOW
      e[u] = (st / SPW) * 32u + k * SPW + (st % SPW);  // blocked CLV layout [block][cat][site in block]
#else
      e[u] = st * K + k;
#endif
    }

    // issue the global loads of one instruction's operands
    auto load_operands = [&](auto nvc, const Instr& in, unsigned fl, Operands<E>& o) __attribute__((always_inline)) {
      constexpr int NV = decltype(nvc)::value;  // slots with an iteration of their own
      if (fl & kTip1) {
        const unsigned char* t = reinterpret_cast<const unsigned char*>(in.c1);
#pragma unroll
        for (int u = 0; u < NV; ++u) o.m1[u] = __ldg(t + site[u]);
      } else if (!(fl & kFwd1) && (fl & (kLoadOnly | kFwd2)) != (kLoadOnly | kFwd2)) {
        const double* g = reinterpret_cast<const double*>(in.c1);
#pragma unroll
        for (int u = 0; u < NV; ++u) o.c1[u] = ld_clv(g, e[u]);
      }
      if (fl & kTip2) {
        const unsigned char* t = reinterpret_cast<const unsigned char*>(in.c2);
#pragma unroll
        for (int u = 0; u < NV; ++u) o.m2[u] = __ldg(t + site[u]);
      } else if (!(fl & (kFwd2 | kLoadOnly))) {
        const double* g = reinterpret_cast<const double*>(in.c2);
#pragma unroll
        for (int u = 0; u < NV; ++u) o.c2[u] = ld_clv(g, e[u]);
      }
#pragma unroll
      for (int u = 0; u < NV; ++u) o.cnt1[u] = o.cnt2[u] = 0;
      if (fl & kLdS1) {
        const unsigned* s1 = in.c1scale;
#pragma unroll
        for (int u = 0; u < NV; ++u) o.cnt1[u] = __ldcg(s1 + site[u]);
      }
      if (fl & kLdS2) {
        const unsigned* s2 = in.c2scale;
#pragma unroll
        for (int u = 0; u < NV; ++u) o.cnt2[u] = __ldcg(s2 + site[u]);
      }
    };

    // pattern weights of this pass's sites (the same for every instruction of the program)
    unsigned wgt[E];
#pragma unroll
    for (int u = 0; u < E; ++u) wgt[u] = __ldg(a.weights + site[u]);

    // one instruction: `cur` holds its operands, the operands of the next instruction are
    // loaded into `nxt`.  A child forwarded from the previous instruction is always child 2
    // (canonical order) and is already in cur.c2: the previous instruction produced its
    // values there.  flc: the instruction's kKindMask flags as a compile-time constant (a
    // fast kind), or < 0: read at run time.  fwdc: the next instruction takes this one's
    // values as its child 2; `v` is then nxt.c2, otherwise a scratch array.  bufc: the
    // parity of the instruction's position in the window = its table buffer.
    auto step = [&](auto nvc, auto flc, auto fwdc, auto bufc, int ii, int wn, Operands<E>& cur, Operands<E>& nxt,
                    d4(&v)[E]) __attribute__((always_inline)) {
      constexpr int      NV = decltype(nvc)::value;
      constexpr int      F = decltype(flc)::value;
      constexpr int      FWDMODE = decltype(fwdc)::value;  // 0: not forwarded, 1: forwarded (v is nxt.c2), 2: run time
      constexpr bool     FWDOUT = FWDMODE == 1;
      constexpr unsigned buf = decltype(bufc)::value;
      const bool more = ii + 1 < wn;
#if RDK_TABLES_L1
      if (more) prefetch_tables(s_prog[ii + 1], 0);
#else
      __syncwarp();  // every lane has finished instruction ii-1 (frees the other table buffer)
      if (more && lane == 0) prefetch_tables(s_prog[ii + 1], buf ^ 1u);
#endif
      const Instr&   in = s_prog[ii];
      const unsigned fl = F < 0 ? in.flags : (unsigned)F;
      unsigned       nfl = 0;
      if (FWDOUT || more) {  // FWDOUT implies a next instruction
        const Instr& nx = s_prog[ii + 1];
        nfl = FWDMODE == 2 ? nx.flags : (FWDOUT ? (nx.flags | kFwd2) : (nx.flags & ~kFwd2));
        if (!RDK_X_NOLOAD) load_operands(nvc, nx, nfl, nxt);
      }
#if RDK_L2_PREFETCH_DIST > 0
      if (ii + RDK_L2_PREFETCH_DIST < wn) {
        const Instr&   fx = s_prog[ii + RDK_L2_PREFETCH_DIST];
        const unsigned ffl = fx.flags;
        if (!(ffl & (kTip1 | kFwd1)) && (ffl & (kLoadOnly | kFwd2)) != (kLoadOnly | kFwd2)) {
          const char* g = reinterpret_cast<const char*>(fx.c1);
#pragma unroll
          for (int u = 0; u < NV; ++u) prefetch_l2(g + (size_t)e[u] * 32u);
        }
        if (!(ffl & (kTip2 | kFwd2 | kLoadOnly))) {
          const char* g = reinterpret_cast<const char*>(fx.c2);
#pragma unroll
          for (int u = 0; u < NV; ++u) prefetch_l2(g + (size_t)e[u] * 32u);
        }
      }
#endif
#if RDK_TABLES_L1
      const unsigned char* tab1 = reinterpret_cast<const unsigned char*>(in.P1);
      const unsigned char* tab2 = reinterpret_cast<const unsigned char*>(in.P2);
#else
#if !RDK_X_NOWAIT
      mbar_wait(&s_bar[buf], (phase >> buf) & 1u);  // tables(ii) have landed
#endif
      phase ^= 1u << buf;
      const unsigned char* tab1 = s_tab + (size_t)buf * 2 * kTabBytes;
      const unsigned char* tab2 = tab1 + kTabBytes;
#endif

      d4 a1[E], a2[E];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        a2[u] = cur.c2[u];
        a1[u] = cur.c1[u];
        if (F < 0 && (fl & kFwd1)) a1[u] = cur.c2[u];  // both children are the forwarded CLV
      }
      unsigned cnt[E];
#pragma unroll
      for (int u = 0; u < NV; ++u) cnt[u] = cur.cnt1[u] + cur.cnt2[u];
      if (fl & kLoadOnly) {
#pragma unroll
        for (int u = 0; u < NV; ++u) v[u] = (fl & kFwd2) ? a2[u] : a1[u];  // kFwd2: the CLV just produced
      } else {
        // child 1
        if (fl & kTip1) {
#pragma unroll
          for (int u = 0; u < NV; ++u) {
            const double2* t = reinterpret_cast<const double2*>(tab1 + (cur.m1[u] * K + k) * 32u);
            const double2  lo = ld_tab(t), hi = ld_tab(t + 1);
            v[u].v[0] = lo.x;
            v[u].v[1] = lo.y;
            v[u].v[2] = hi.x;
            v[u].v[3] = hi.y;
          }
        } else {
          const double2* p = reinterpret_cast<const double2*>(tab1 + k * (kPTabDoubles * 8));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double2 p01 = ld_tab(p + i * 2), p23 = ld_tab(p + i * 2 + 1);
#pragma unroll
            for (int u = 0; u < NV; ++u) {
              double s = dmul(p01.x, a1[u].v[0]);
              s = dfma(p01.y, a1[u].v[1], s);
              s = dfma(p23.x, a1[u].v[2], s);
              s = dfma(p23.y, a1[u].v[3], s);
              v[u].v[i] = s;
            }
          }
        }
        // child 2
        if (fl & kTip2) {
#pragma unroll
          for (int u = 0; u < NV; ++u) {
            const double2* t = reinterpret_cast<const double2*>(tab2 + (cur.m2[u] * K + k) * 32u);
            const double2  lo = ld_tab(t), hi = ld_tab(t + 1);
            v[u].v[0] = dmul(v[u].v[0], lo.x);
            v[u].v[1] = dmul(v[u].v[1], lo.y);
            v[u].v[2] = dmul(v[u].v[2], hi.x);
            v[u].v[3] = dmul(v[u].v[3], hi.y);
          }
        } else {
          const double2* p = reinterpret_cast<const double2*>(tab2 + k * (kPTabDoubles * 8));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double2 p01 = ld_tab(p + i * 2), p23 = ld_tab(p + i * 2 + 1);
#pragma unroll
            for (int u = 0; u < NV; ++u) {
              double s = dmul(p01.x, a2[u].v[0]);
              s = dfma(p01.y, a2[u].v[1], s);
              s = dfma(p23.x, a2[u].v[2], s);
              s = dfma(p23.y, a2[u].v[3], s);
              v[u].v[i] = dmul(v[u].v[i], s);
            }
          }
        }
      }
      if (fl & kScale) {
        // (a branch-free form -- all E ballots first, then a multiplication by 2^256 or 1.0 --
        // was measured on B200: 3 % faster on a 12.5 k-site shard, 1.5 % slower at 100 k)
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          const bool small = (v[u].v[0] < RDK_SCALE_THRESHOLD) && (v[u].v[1] < RDK_SCALE_THRESHOLD) &&
                             (v[u].v[2] < RDK_SCALE_THRESHOLD) && (v[u].v[3] < RDK_SCALE_THRESHOLD);
          const unsigned m = __ballot_sync(0xffffffffu, small);
          if ((m & gmask) == gmask) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[u].v[i] = dmul(v[u].v[i], RDK_SCALE_FACTOR);
            cnt[u] += 1;
          }
        }
      }
      if ((fl & kWrite) && !RDK_X_NOSTORE) {
        double* par = in.parent;
#pragma unroll
        for (int u = 0; u < NV; ++u) st_clv(par, e[u], v[u]);
        unsigned* ps = in.pscale;
        if (ps && k == 0) {
#pragma unroll
          for (int u = 0; u < NV; ++u) __stcg(ps + site[u], cnt[u]);
        }
      }
      if (FWDMODE == 2 && (nfl & kFwd2)) {
#pragma unroll
        for (int u = 0; u < NV; ++u) nxt.c2[u] = v[u];
      }
      // the scaler counts of a forwarded child of instruction ii+1 (its values are `v`)
      if (nfl & kFwdS1) {
#pragma unroll
        for (int u = 0; u < NV; ++u) nxt.cnt1[u] = cnt[u];
      }
      if (nfl & kFwdS2) {
#pragma unroll
        for (int u = 0; u < NV; ++u) nxt.cnt2[u] = cnt[u];
      }
      if ((fl & kEval) && !RDK_X_NOEVAL) {
        // every lane of a site gathers the K category terms in category order
        double term[E];
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          double t = dmul(a.pi[0], v[u].v[0]);
          t = dfma(a.pi[1], v[u].v[1], t);
          t = dfma(a.pi[2], v[u].v[2], t);
          t = dfma(a.pi[3], v[u].v[3], t);
          double tm = dmul(a.w[0], __shfl_sync(0xffffffffu, t, lane0));
#pragma unroll
          for (int kk = 1; kk < K; ++kk) {
            double tk = __shfl_sync(0xffffffffu, t, lane0 + kk * KSTRIDE);
            tm = dfma(a.w[kk], tk, tm);
          }
          term[u] = tm;
        }
        if constexpr (NV <= K) {
          // the K lanes of a site hold the same term: lane k takes the logarithm of the
          // site of slot u = k, so that ONE pass through rd_log serves all NV slots
          double   x = term[0];
          unsigned cn = cnt[0], st = site[0], iw = it[0], wg = wgt[0];
#pragma unroll
          for (int u = 1; u < NV; ++u)
            if (k == (unsigned)u) {
              x = term[u];
              cn = cnt[u];
              st = site[u];
              iw = it[u];
              wg = wgt[u];
            }
          double l = 0.0;
          if (k < (unsigned)NV && iw * SPW + sl <= last_site) {
            l = rd_log(x);
            if (fl & kEvalScaler) l = dadd(l, dmul((double)cn, RDK_LOG_SCALE_THRESHOLD));
            l = dmul(l, (double)wg);
            if (a.persite && in.slot == 0) a.persite[st] = l;
          }
          // canonical tree over the 32/K sites of a warp iteration (the lanes of equal k)
#pragma unroll
          for (int off2 = 1; off2 < (int)SPW; off2 <<= 1)
            l = dadd(l, __shfl_xor_sync(0xffffffffu, l, off2 * (KSTRIDE == 1 ? K : 1)));
          if (sl == 0 && k < (unsigned)NV) a.partials[(size_t)in.slot * a.partial_stride + iw] = l;
        } else {
#pragma unroll
          for (int u = 0; u < NV; ++u) {
            double l = 0.0;
            if (k == 0 && it[u] * SPW + sl <= last_site) {
              l = rd_log(term[u]);
              if (fl & kEvalScaler) l = dadd(l, dmul((double)cnt[u], RDK_LOG_SCALE_THRESHOLD));
              l = dmul(l, (double)wgt[u]);
              if (a.persite && in.slot == 0) a.persite[site[u]] = l;
            }
#pragma unroll
            for (int off2 = 1; off2 < (int)SPW; off2 <<= 1)
              l = dadd(l, __shfl_xor_sync(0xffffffffu, l, off2 * (KSTRIDE == 1 ? K : 1)));
            if (lane == 0) a.partials[(size_t)in.slot * a.partial_stride + it[u]] = l;
          }
        }
      }
    };

    // run instruction ii through the copy of the body compiled for its kind
    auto dispatch = [&](auto nvc, auto bufc, int ii, int wn, Operands<E>& cur, Operands<E>& nxt)
                        __attribute__((always_inline)) {
      using IC = std::integral_constant<int, -1>;
      using std::integral_constant;
      using no_fwd = integral_constant<int, 0>;
      using fwd = integral_constant<int, 1>;
      d4             scratch[E];
      const unsigned kind = s_prog[ii].kind;
      // (short tail passes, NV < E, run the run-time decoded body: they are rare)
      if constexpr (FAST && (decltype(nvc)::value == E || RDK_FAST_KINDS >= 2)) {
        switch (kind) {
          case 2: step(nvc, integral_constant<int, (int)kFastKinds[0]>{}, no_fwd{}, bufc, ii, wn, cur, nxt, scratch); break;
          case 3: step(nvc, integral_constant<int, (int)kFastKinds[0]>{}, fwd{}, bufc, ii, wn, cur, nxt, nxt.c2); break;
          case 4: step(nvc, integral_constant<int, (int)kFastKinds[1]>{}, no_fwd{}, bufc, ii, wn, cur, nxt, scratch); break;
          case 5: step(nvc, integral_constant<int, (int)kFastKinds[1]>{}, fwd{}, bufc, ii, wn, cur, nxt, nxt.c2); break;
          case 6: step(nvc, integral_constant<int, (int)kFastKinds[2]>{}, no_fwd{}, bufc, ii, wn, cur, nxt, scratch); break;
          case 7: step(nvc, integral_constant<int, (int)kFastKinds[2]>{}, fwd{}, bufc, ii, wn, cur, nxt, nxt.c2); break;
          case 8: step(nvc, integral_constant<int, (int)kFastKinds[3]>{}, no_fwd{}, bufc, ii, wn, cur, nxt, scratch); break;
          case 10: step(nvc, integral_constant<int, (int)kFastKinds[4]>{}, no_fwd{}, bufc, ii, wn, cur, nxt, scratch); break;
          case 1: step(nvc, IC{}, fwd{}, bufc, ii, wn, cur, nxt, nxt.c2); break;
          default: step(nvc, IC{}, no_fwd{}, bufc, ii, wn, cur, nxt, scratch); break;
        }
      } else if constexpr (RDK_FWD_STATIC) {
        if (kind & 1u)
          step(nvc, IC{}, fwd{}, bufc, ii, wn, cur, nxt, nxt.c2);
        else
          step(nvc, IC{}, no_fwd{}, bufc, ii, wn, cur, nxt, scratch);
      } else {
        step(nvc, IC{}, integral_constant<int, 2>{}, bufc, ii, wn, cur, nxt, scratch);
      }
    };

    // the instruction loop over one staged window, for NV slots
    auto run_window = [&](auto nvc, int wn) __attribute__((always_inline)) {
      using std::integral_constant;
#if RDK_TABLES_L1
      prefetch_tables(s_prog[0], 0);
#else
      __syncwarp();
      if (lane == 0) prefetch_tables(s_prog[0], 0);
#endif
      Operands<E> opA, opB;
#pragma unroll
      for (int u = 0; u < E; ++u) {
#pragma unroll
        for (int i = 0; i < 4; ++i) opA.c1[u].v[i] = opA.c2[u].v[i] = opB.c1[u].v[i] = opB.c2[u].v[i] = 0.0;
        opA.m1[u] = opA.m2[u] = opB.m1[u] = opB.m2[u] = 15u;  // code 15: all-zero table row
        opA.cnt1[u] = opA.cnt2[u] = opB.cnt1[u] = opB.cnt2[u] = 0;
      }
      // the first instruction of a window never carries kFwd* (finalize_program)
      load_operands(nvc, s_prog[0], s_prog[0].flags, opA);
      int ii = 0;
      for (; ii + 1 < wn; ii += 2) {
        dispatch(nvc, integral_constant<unsigned, 0>{}, ii, wn, opA, opB);
        dispatch(nvc, integral_constant<unsigned, 1>{}, ii + 1, wn, opB, opA);
      }
      if (ii < wn) dispatch(nvc, integral_constant<unsigned, 0>{}, ii, wn, opA, opB);
    };

    for (int w0 = 0; w0 < n_instr; w0 += kProgWindow) {
      const int wn = min(kProgWindow, n_instr - w0);
      if (multi_window || pass == 0) {
        if (multi_window) __syncthreads();  // every warp is done with the previous window
        const int4* src = reinterpret_cast<const int4*>(a.n_instr <= kProgInline ? a.inl : prog + w0);
        int4*       dst = reinterpret_cast<int4*>(s_prog);
        for (unsigned c = tid; c < (unsigned)wn * (sizeof(Instr) / 16); c += blockDim.x) dst[c] = src[c];
        __syncthreads();
      }
      if (!active) continue;
      // iterations of its own this warp has in this pass (warp-uniform)
      const unsigned nvalid = min((unsigned)E, it_end - (it_begin + pass * E));
      if (TS && E >= 2 && nvalid == 1)
        run_window(std::integral_constant<int, 1>{}, wn);
      else if (TS && E >= 4 && nvalid == 2)
        run_window(std::integral_constant<int, (TS && E >= 4 ? 2 : E)>{}, wn);
      else
        run_window(std::integral_constant<int, E>{}, wn);
    }
  }
}

// Launch of the program kernel for K rate categories: one explicit specialisation per K,
// each in its own translation unit (rdk_program_inst.cu).  Returns the CUDA status of the
// launch configuration (the launch itself is checked by the caller with cudaGetLastError).
template <int K>
cudaError_t launch_program(const ProgArgs& a, int grid, int threads, int E, bool tail_skip, cudaStream_t st);
template <> cudaError_t launch_program<1>(const ProgArgs&, int, int, int, bool, cudaStream_t);
template <> cudaError_t launch_program<2>(const ProgArgs&, int, int, int, bool, cudaStream_t);
template <> cudaError_t launch_program<4>(const ProgArgs&, int, int, int, bool, cudaStream_t);
template <> cudaError_t launch_program<8>(const ProgArgs&, int, int, int, bool, cudaStream_t);
template <> cudaError_t launch_program<16>(const ProgArgs&, int, int, int, bool, cudaStream_t);
template <> cudaError_t launch_program<32>(const ProgArgs&, int, int, int, bool, cudaStream_t);

#ifndef RDK_PROGRAM_KERNEL_ONLY
// ---------------------------------------------------------------------------
// Canonical reduction: balanced binary tree over contiguous halves of the
// zero-padded (to a power of two) GLOBAL site index space.  The program kernel
// produced the tree nodes that cover one warp iteration (32/K sites); this
// kernel continues the same tree.  One block per (slot, node-range).
//   in : [slots][in_stride], n_in valid leaves per slot, leaf index offset
//        `leaf0` within the padded tree (sharded partitions)
//   out: [slots][out_stride]; out[slot][b] = tree node over leaves
//        [b*span, (b+1)*span) (span = power of two)
// With span >= n_in and one block per slot this yields the final sum.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tree_reduce_kernel(const double* __restrict__ in,
                                                           unsigned in_stride, unsigned n_in,
                                                           unsigned span, double* __restrict__ out,
                                                           unsigned out_stride, unsigned out_offset) {
  __shared__ double sm[256];
  const unsigned slot = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
  const double*  p = in + (size_t)slot * in_stride;
  const unsigned R = span < 256u ? span : 256u;  // power of two
  const unsigned L = span / R;                   // power of two
  double         res = 0.0;
  if (t < R) {
    const unsigned long long base = (unsigned long long)b * span + (unsigned long long)t * L;
    if (base < n_in) {
      double st[32];
      int    top = 0;
      for (unsigned i = 0; i < L; ++i) {
        unsigned long long idx = base + i;
        double             s = idx < n_in ? p[idx] : 0.0;
        for (unsigned j = i; j & 1u; j >>= 1) s = dadd(st[--top], s);
        st[top++] = s;
      }
      res = st[0];
    }
  }
  sm[t] = res;
  __syncthreads();
  for (unsigned w = 1; w < R; w <<= 1) {
    if (t < R && (t % (2 * w)) == 0) sm[t] = dadd(sm[t], sm[t + w]);
    __syncthreads();
  }
  if (t == 0) out[(size_t)slot * out_stride + out_offset + b] = sm[0];
}

// ---------------------------------------------------------------------------
// Weighted histogram of tip state masks (exact integer arithmetic), the input
// of the empirical base frequencies (corax_msa_empirical_frequencies).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tip_hist_kernel(const unsigned char* __restrict__ tips,
                                                        size_t tip_stride, unsigned ntips,
                                                        unsigned sites,
                                                        const unsigned* __restrict__ weights,
                                                        unsigned long long* __restrict__ hist) {
  __shared__ unsigned long long sh[16];
  if (threadIdx.x < 16) sh[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long local[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) local[i] = 0;
  const size_t total = (size_t)ntips * sites;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    unsigned tip = (unsigned)(idx / sites), s = (unsigned)(idx - (size_t)tip * sites);
    unsigned m = tip_mask_of_code(tips[(size_t)tip * tip_stride + s] & 15u);
    unsigned w = weights[s];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (m == (unsigned)i) local[i] += w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (local[i]) atomicAdd(&sh[i], local[i]);
  __syncthreads();
  if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// expand a tip's byte masks to a full 0/1 CLV (rdk_get_clv on a tip index)
__global__ void tip_expand_kernel(const unsigned char* __restrict__ tip, unsigned sites, int K,
                                  double* __restrict__ out) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)sites * K) return;
  unsigned m = tip_mask_of_code(tip[e / K] & 15u);
  for (int j = 0; j < 4; ++j) out[e * 4 + j] = ((m >> j) & 1u) ? 1.0 : 0.0;
}

#endif  // RDK_PROGRAM_KERNEL_ONLY

}  // namespace rdk
