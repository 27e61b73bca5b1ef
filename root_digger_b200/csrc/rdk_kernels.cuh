// rdk_kernels.cuh -- sm_100a kernels of the RootDigger likelihood engine.
//
// Replaces the arithmetic behind corax_update_prob_matrices,
// corax_update_clvs and corax_compute_root_loglikelihood (call sites:
// reference src/model.cpp:367,432,842 / :402,440,461,851 / :406,441,466).
//
// Arithmetic contract (DESIGN.md "Arithmetic specification"): fp64, every
// +,-,*,/ individually rounded (round-to-nearest-even, no FMA contraction), in
// the order written.  All arithmetic goes through the __d*_rn intrinsics so the
// contract holds whatever -fmad says.
//
// Memory-bound by design: 0.61 flop/B, no tensor cores (a 4x4 mat-vec is not a
// dense contraction).  What matters here is coalesced 32-byte-per-thread
// accesses, enough bytes in flight, and never re-reading a CLV from HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rdk {

constexpr int kMaxCats = 32;  // rate categories handled by the fast path: K | 32

// ---------------------------------------------------------------------------
// individually rounded arithmetic
// ---------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

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
// pmat_expm_nonrev: one thread per (branch entry, rate category).
// Replaces corax_update_prob_matrices.  Output layout [slot][cat][i][j].
// ---------------------------------------------------------------------------
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
  double* out = a.pool + ((size_t)ent.slot * a.K + k) * 16;
  for (int i = 0; i < 16; ++i) out[i] = E[i];
}

// ---------------------------------------------------------------------------
// The likelihood program kernel.
//
// A "program" is a list of instructions executed in order for every
// (site, category) element.  Elements are independent of each other (the
// dependency between a parent CLV and its children is per site), so each warp
// owns a contiguous range of elements and walks the WHOLE program on it with
// no block or grid synchronisation: a full post-order traversal (n-1 CLV
// operations + the root log-likelihood), a root move, or an entire placement
// sweep is ONE launch.  Child CLVs produced a few instructions earlier by the
// same warp are still in L2, so HBM sees each CLV written once and mostly not
// read back.
//
// Thread mapping: element e = site*K + k.  Lane l of a warp iteration `it`
// handles e = 32*it + l, i.e. 32 consecutive 32-byte (4 x fp64) vectors = 1 KiB
// contiguous per CLV per warp access; k = l % K is fixed per thread (K | 32), so
// the thread's two 4x4 P-matrices live in registers for E iterations.
// ---------------------------------------------------------------------------
enum : unsigned {
  kTip1 = 1u,      // child1 is a tip: 1 byte per site (4-bit state mask)
  kTip2 = 2u,      // child2 is a tip
  kWrite = 4u,     // store the parent CLV (and parent scaler if present)
  kEval = 8u,      // evaluate the root log-likelihood of the parent values
  kLoadOnly = 16u, // no CLV arithmetic: parent values := CLV at c1 (root logL of
                   // a stored CLV, corax_compute_root_loglikelihood)
  kScale = 32u,    // parent has a scale buffer: apply 2^256 rescaling
};

struct alignas(16) Instr {
  double*         parent;
  const void*     c1;
  const void*     c2;
  unsigned*       pscale;
  const unsigned* c1scale;
  const unsigned* c2scale;
  const double*   P1;
  const double*   P2;
  unsigned        flags;
  unsigned        slot;  // eval slot (row of the partial-sum buffer)
  unsigned long long pad;
};
static_assert(sizeof(Instr) == 80, "Instr layout");

constexpr int kProgInline = 8;
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
  Instr           inl[kProgInline];
};

struct d4 {
  double v[4];
};

__device__ __forceinline__ d4 ld_clv(const double* p) {
  // 32 B per thread as 2 x 128-bit loads; plain (coherent) loads because CLVs
  // are read and written within one launch.
  const double2* q = reinterpret_cast<const double2*>(p);
  double2        a = q[0], b = q[1];
  d4             r;
  r.v[0] = a.x;
  r.v[1] = a.y;
  r.v[2] = b.x;
  r.v[3] = b.y;
  return r;
}
__device__ __forceinline__ void st_clv(double* p, const d4& x) {
  double2* q = reinterpret_cast<double2*>(p);
  q[0] = make_double2(x.v[0], x.v[1]);
  q[1] = make_double2(x.v[2], x.v[3]);
}
__device__ __forceinline__ d4 tip_vec(unsigned m) {
  d4 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = ((m >> j) & 1u) ? 1.0 : 0.0;
  return r;
}
__device__ __forceinline__ void ld_p(const double* P, double* out) {
  const double2* q = reinterpret_cast<const double2*>(P);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double2 t = __ldg(q + i);
    out[2 * i] = t.x;
    out[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ d4 matvec(const double* P, const d4& c) {
  d4 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double s = dmul(P[i * 4 + 0], c.v[0]);
    s = dadd(s, dmul(P[i * 4 + 1], c.v[1]));
    s = dadd(s, dmul(P[i * 4 + 2], c.v[2]));
    s = dadd(s, dmul(P[i * 4 + 3], c.v[3]));
    r.v[i] = s;
  }
  return r;
}

template <int K, int E>
__global__ void __launch_bounds__(256) clv_program_kernel(const __grid_constant__ ProgArgs a) {
  static_assert(32 % K == 0, "K must divide the warp size");
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned k = lane % K;
  const bool     k0 = (k == 0);
  // the K lanes of this thread's site
  const unsigned gmask = (K == 32) ? 0xffffffffu : (((1u << K) - 1u) << (lane - k));

  const unsigned it_begin = (unsigned)(((unsigned long long)a.n_witer * warp) / nwarps);
  const unsigned it_end = (unsigned)(((unsigned long long)a.n_witer * (warp + 1)) / nwarps);
  const bool     inl = a.n_instr <= kProgInline;

  for (unsigned it0 = it_begin; it0 < it_end; it0 += E) {
    for (int ii = 0; ii < a.n_instr; ++ii) {
      const Instr in = inl ? a.inl[ii] : a.prog[ii];
      const unsigned fl = in.flags;
      double         P1[16], P2[16];
      if (!(fl & kLoadOnly)) {
        ld_p(in.P1 + k * 16, P1);
        ld_p(in.P2 + k * 16, P2);
      }
      d4       c1[E], c2[E];
      unsigned cnt[E];
      bool     valid[E];
      // phase 1: issue every load of this instruction
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const unsigned it = it0 + u;
        const unsigned e = it * 32u + lane;
        valid[u] = (it < it_end) && (e < a.nelem);
        cnt[u] = 0;
        if (valid[u]) {
          const unsigned site = e / K;
          if (fl & kTip1)
            c1[u] = tip_vec(__ldg(reinterpret_cast<const unsigned char*>(in.c1) + site));
          else
            c1[u] = ld_clv(reinterpret_cast<const double*>(in.c1) + (size_t)e * 4);
          if (!(fl & kLoadOnly)) {
            if (fl & kTip2)
              c2[u] = tip_vec(__ldg(reinterpret_cast<const unsigned char*>(in.c2) + site));
            else
              c2[u] = ld_clv(reinterpret_cast<const double*>(in.c2) + (size_t)e * 4);
          }
          if (k0) {
            if (in.c1scale) cnt[u] = in.c1scale[site];
            if (in.c2scale) cnt[u] += in.c2scale[site];
          }
        }
      }
      // phase 2: arithmetic, rescaling, stores, log-likelihood
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const unsigned it = it0 + u;
        if (it >= it_end) break;  // warp-uniform
        const unsigned e = it * 32u + lane;
        const unsigned site = e / K;
        d4             v;
        if (fl & kLoadOnly) {
          v = c1[u];
        } else {
          d4 x = matvec(P1, c1[u]);
          d4 y = matvec(P2, c2[u]);
#pragma unroll
          for (int i = 0; i < 4; ++i) v.v[i] = dmul(x.v[i], y.v[i]);
        }
        if (fl & kScale) {
          bool small = valid[u] && (v.v[0] < RDK_SCALE_THRESHOLD) && (v.v[1] < RDK_SCALE_THRESHOLD) &&
                       (v.v[2] < RDK_SCALE_THRESHOLD) && (v.v[3] < RDK_SCALE_THRESHOLD);
          unsigned m = __ballot_sync(0xffffffffu, small);
          if ((m & gmask) == gmask) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v.v[i] = dmul(v.v[i], RDK_SCALE_FACTOR);
            cnt[u] += 1;
          }
        }
        if ((fl & kWrite) && valid[u]) {
          st_clv(in.parent + (size_t)e * 4, v);
          if (k0 && in.pscale) in.pscale[site] = cnt[u];
        }
        if (fl & kEval) {
          double t = dmul(a.pi[0], v.v[0]);
          t = dadd(t, dmul(a.pi[1], v.v[1]));
          t = dadd(t, dmul(a.pi[2], v.v[2]));
          t = dadd(t, dmul(a.pi[3], v.v[3]));
          double term = dmul(a.w[0], t);
#pragma unroll
          for (int kk = 1; kk < K; ++kk) {
            double tk = __shfl_down_sync(0xffffffffu, t, kk);
            term = dadd(term, dmul(a.w[kk], tk));
          }
          double l = 0.0;
          if (valid[u] && k0) {
            l = rd_log(term);
            // scaler term only when a scale buffer is attached to the root
            const bool has_scaler = (fl & kLoadOnly) ? (in.c1scale != nullptr) : ((fl & kScale) != 0);
            if (has_scaler) l = dadd(l, dmul((double)cnt[u], RDK_LOG_SCALE_THRESHOLD));
            l = dmul(l, (double)__ldg(a.weights + site));
            if (a.persite && in.slot == 0) a.persite[site] = l;
          }
          // canonical tree over the 32/K sites of this warp iteration
#pragma unroll
          for (int off = K; off < 32; off <<= 1) l = dadd(l, __shfl_xor_sync(0xffffffffu, l, off));
          if (lane == 0) a.partials[(size_t)in.slot * a.partial_stride + it] = l;
        }
      }
    }
  }
}

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
    unsigned m = tips[(size_t)tip * tip_stride + s] & 15u;
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
  unsigned m = tip[e / K];
  for (int j = 0; j < 4; ++j) out[e * 4 + j] = ((m >> j) & 1u) ? 1.0 : 0.0;
}

}  // namespace rdk
