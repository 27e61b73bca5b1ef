// rdk_lower.hpp -- from recorded CLV operations to the program the walk kernel executes.
//
// Pure host code (no CUDA): buffer indices in, device-instruction descriptions out, so that the
// data flow of a lowered program can be checked on a CPU (tests/test_lowering.py runs it through
// rdk_debug_lower_program and interprets the result symbolically).
//
// What the kernel's warps carry from one instruction to the next is ONE register-resident CLV
// value per element, `v` (with its scaler count): every instruction computes
//     r = A(child 1) o B(child 2)
// where A is a mat-vec on a CLV LOADED from memory or a tip-table lookup, and B is a mat-vec on
// `v` or a tip-table lookup.  A post-order traversal (reference src/tree.cpp:364-413) and the
// directed placement sweep produce, almost always, the second inner child in the instruction just
// before its parent, so B = v needs no memory access; when it does not hold the right CLV the
// lowering inserts a pseudo-instruction "v := load(CLV)".  The product of the two children's terms
// is commutative bit for bit, so the children are ordered canonically (tip first / `v` last).
//
// The second job of the lowering is liveness: a CLV (or scale buffer) store is dropped when no
// later instruction reads the buffer from memory before it is overwritten and the buffer's
// content after the program is not required (scratch buffers of a directed sweep; the buffers of
// a lazily materialised evaluation).
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

namespace rdk {

// ---- recorded operation (what the ABI entry points append) -------------------------------------
enum : unsigned {
  rWrite = 1u,     // store the parent CLV and its scaler counts
  rEval = 2u,      // evaluate the root log-likelihood of the parent values into eval slot `slot`
  rLoadOnly = 4u,  // no CLV arithmetic: the "parent values" are the stored CLV c1 (with c1scale)
};
constexpr unsigned kNoClv = 0xffffffffu;

struct ROp {
  unsigned parent;  // clv index (tips first, as in corax_operation_t) or kNoClv
  int      pscale;  // scale buffer index or -1
  unsigned c1, c2;  // clv indices
  int      c1scale, c2scale;
  unsigned pm1, pm2;  // PHYSICAL P-matrix pool slots of the two child branches
  unsigned flags;
  unsigned slot;
  unsigned id;  // position in the recorded program (set by the engine before lowering; carried into LInstr::src)
};

// ---- device instruction flags (shared with the kernel) --------------------------------------------
enum : unsigned {
  fTip1 = 1u,         // A = tip-table lookup of tip row c1; otherwise mat-vec on the loaded CLV c1
  fTip2 = 2u,         // B = tip-table lookup of tip row c2; otherwise mat-vec on v
  fWrite = 4u,        // store r to `parent`
  fWriteS = 8u,       // store the scaler counts to `pscale`
  fEval = 16u,        // evaluate r and leave v untouched (r is not stored, not forwarded)
  fScale = 32u,       // the parent has a scale buffer: 2^256 rescaling + count
  fLoadV = 64u,       // pseudo-instruction: v := CLV c1 (vcnt := its scaler counts)
  fNop = 128u,        // no main action (only fEvalV)
  fEvalV = 256u,      // after the main action, evaluate v
  fCnt1 = 512u,       // child 1 (or the CLV of fLoadV) has scaler counts in memory (c1scale)
  fCnt2V = 1024u,     // child 2 = v carries scaler counts
  fEvalScaler = 2048u,  // the evaluated values carry scaler counts (adds cnt * ln 2^-256)
  fLoadV2 = 4096u,    // before this instruction runs, v := CLV c2 (loaded after the previous
                      // instruction's stores, while that instruction finishes)
  fCnt2M = 8192u,     // ... and vcnt := its scaler counts from memory (c2scale)
};

struct LInstr {
  unsigned flags;
  unsigned parent;  // clv index stored when fWrite
  int      pscale;  // scale buffer stored when fWriteS
  unsigned c1;      // clv index: tip row (fTip1) or inner CLV loaded from memory
  int      c1scale;
  unsigned c2;      // tip row when fTip2; the inner CLV loaded into v when fLoadV2
  int      c2scale; // its scale buffer (fCnt2M)
  unsigned pm1, pm2;  // pool slots of the tables A / B read
  unsigned slot;
  // where pm1 / pm2 came from: recorded operation `src` (ROp::id), children exchanged or not -- a
  // lowered program kept for the next identical traversal only needs its P slots refreshed
  unsigned src;
  bool     swapped;
};

constexpr unsigned kNoSrc = 0xffffffffu;

struct LowerOptions {
  unsigned tips = 0;
  // content after the program is NOT required for: every written buffer (discard_writes), or the
  // clv indices / scale buffers flagged in these per-index tables (may be null)
  bool                     discard_writes = false;
  const std::vector<char> *scratch_clv = nullptr;
  const std::vector<char> *scratch_scaler = nullptr;
};

struct LowerStats {
  unsigned loadv = 0;         // pseudo-instructions inserted
  unsigned stores_dropped = 0;  // CLV stores removed by the liveness pass
  unsigned forwarded = 0;     // inner children taken from v
};

// `chunk_off` (size >= 2) delimits independent sub-programs of `ops`; `out_chunk_off` receives the
// same boundaries in instructions of `out`.
inline void lower_program(const std::vector<ROp> &ops, const std::vector<unsigned> &chunk_off,
                          const LowerOptions &opt, std::vector<LInstr> &out,
                          std::vector<unsigned> &out_chunk_off, LowerStats *stats = nullptr) {
  out.clear();
  out_chunk_off.clear();
  LowerStats        st;
  const unsigned    tips = opt.tips;
  std::vector<unsigned> bounds = chunk_off;
  if (bounds.size() < 2) bounds = {0u, (unsigned)ops.size()};
  auto is_tip = [&](unsigned c) { return c < tips; };

  for (size_t ch = 0; ch + 1 < bounds.size(); ++ch) {
    out_chunk_off.push_back((unsigned)out.size());
    const size_t chunk_first = out.size();
    unsigned     vptr = kNoClv;  // the CLV held in v
    int          vscale = -1;
    // The operands an instruction loads from memory are fetched while the PREVIOUS instruction
    // computes, i.e. before that instruction's stores are issued: a buffer stored by the
    // instruction just before must not be loaded (it normally is not: its values are in v).
    // The degenerate cases (both children the same CLV; same CLV under another scale buffer)
    // get a no-op in between.
    auto guard_raw = [&](unsigned clv, int scaler) {
      if (out.size() == chunk_first) return;
      const LInstr &pv = out.back();
      const bool    hazard = ((pv.flags & fWrite) && pv.parent == clv) ||
                          ((pv.flags & fWriteS) && scaler >= 0 && pv.pscale == scaler);
      if (!hazard) return;
      LInstr nop{};
      nop.src = kNoSrc;
      nop.flags = fNop;
      nop.parent = kNoClv;
      nop.pscale = -1;
      nop.c1 = kNoClv;
      nop.c1scale = -1;
      nop.c2 = kNoClv;
      nop.c2scale = -1;
      out.push_back(nop);
    };
    for (unsigned i = bounds[ch]; i < bounds[ch + 1]; ++i) {
      const ROp &op = ops[i];
      LInstr     li{};
      li.parent = kNoClv;
      li.pscale = -1;
      li.c2 = kNoClv;
      li.c1scale = li.c2scale = -1;
      li.slot = op.slot;
      li.src = kNoSrc;
      if (op.flags & rLoadOnly) {
        // evaluate the stored CLV c1
        const bool in_v = vptr == op.c1 && vscale == op.c1scale;
        if (in_v) {
          li.flags = fNop | fEvalV;
          li.c1 = kNoClv;
          ++st.forwarded;
        } else {
          guard_raw(op.c1, op.c1scale);
          li.flags = fLoadV | fEvalV | (op.c1scale >= 0 ? fCnt1 : 0u);
          li.c1 = op.c1;
          li.c1scale = op.c1scale;
          vptr = op.c1;
          vscale = op.c1scale;
        }
        if (op.c1scale >= 0) li.flags |= fEvalScaler;
        out.push_back(li);
        continue;
      }
      unsigned c1 = op.c1, c2 = op.c2, pm1 = op.pm1, pm2 = op.pm2;
      int      s1 = op.c1scale, s2 = op.c2scale;
      bool swapped = false;
      auto swap_children = [&]() {
        std::swap(c1, c2);
        std::swap(pm1, pm2);
        std::swap(s1, s2);
        swapped = !swapped;
      };
      unsigned fl = 0;
      bool     load_v = false;  // v does not hold child 2: load it (fLoadV2)
      if (is_tip(c1) && is_tip(c2)) {
        fl |= fTip1 | fTip2;
        if (s2 >= 0 && s1 < 0) swap_children();  // only child 1's counts can come from memory
        s2 = -1;                                 // (the ABI rejects two tips with scale buffers)
      } else if (is_tip(c1) || is_tip(c2)) {
        if (is_tip(c2)) swap_children();  // the tip first
        fl |= fTip1;
        if (!(vptr == c2 && vscale == s2)) load_v = true;
        else ++st.forwarded;
      } else {
        if (vptr == c2 && vscale == s2) {
          ++st.forwarded;
        } else if (vptr == c1 && vscale == s1) {
          swap_children();
          ++st.forwarded;
        } else {
          load_v = true;
        }
      }
      if (load_v) {
        fl |= fLoadV2 | (s2 >= 0 ? fCnt2M : 0u);
        ++st.loadv;
      }
      if (s1 >= 0) fl |= fCnt1;
      if (!(fl & fTip2) && s2 >= 0) fl |= fCnt2V;
      if (op.pscale >= 0) fl |= fScale;
      li.c1 = c1;
      li.c1scale = s1;
      li.c2 = (fl & (fTip2 | fLoadV2)) ? c2 : kNoClv;
      li.c2scale = (fl & fCnt2M) ? s2 : -1;
      li.pm1 = pm1;
      li.pm2 = pm2;
      li.src = op.id;
      li.swapped = swapped;
      if ((op.flags & rEval) && !(op.flags & rWrite)) {
        fl |= fEval;
        if (fl & fScale) fl |= fEvalScaler;
      } else {
        if (op.flags & rWrite) {
          fl |= fWrite;
          li.parent = op.parent;
          if (op.pscale >= 0) {
            fl |= fWriteS;
            li.pscale = op.pscale;
          }
        }
        if (op.flags & rEval) {
          fl |= fEvalV;
          if (fl & fScale) fl |= fEvalScaler;
        }
        vptr = op.parent;
        vscale = op.pscale;
      }
      li.flags = fl;
      guard_raw((fl & fTip1) ? kNoClv : li.c1, li.c1scale);
      out.push_back(li);
    }

    // ---- liveness: drop the stores nobody reads back -----------------------------------------
    // walked backwards over this chunk; live[x] = "a later instruction reads x from memory before
    // x is overwritten, or x must hold its value after the program"
    // dense tables sized by the largest index seen in the chunk
    unsigned max_clv = 0, max_sc = 0;
    for (size_t j = chunk_first; j < out.size(); ++j) {
      const LInstr &x = out[j];
      if (x.parent != kNoClv) max_clv = std::max(max_clv, x.parent + 1);
      if (x.c1 != kNoClv) max_clv = std::max(max_clv, x.c1 + 1);
      if (x.c2 != kNoClv) max_clv = std::max(max_clv, x.c2 + 1);
      if (x.c2scale >= 0) max_sc = std::max(max_sc, (unsigned)x.c2scale + 1);
      if (x.pscale >= 0) max_sc = std::max(max_sc, (unsigned)x.pscale + 1);
      if (x.c1scale >= 0) max_sc = std::max(max_sc, (unsigned)x.c1scale + 1);
    }
    auto out_needed_clv = [&](unsigned c) {
      if (opt.discard_writes) return false;
      if (opt.scratch_clv && c < opt.scratch_clv->size() && (*opt.scratch_clv)[c]) return false;
      return true;
    };
    auto out_needed_sc = [&](unsigned s) {
      if (opt.discard_writes) return false;
      if (opt.scratch_scaler && s < opt.scratch_scaler->size() && (*opt.scratch_scaler)[s]) return false;
      return true;
    };
    std::vector<char> live_clv(max_clv), live_sc(max_sc);
    for (unsigned c = 0; c < max_clv; ++c) live_clv[c] = out_needed_clv(c) ? 1 : 0;
    for (unsigned s = 0; s < max_sc; ++s) live_sc[s] = out_needed_sc(s) ? 1 : 0;
    for (size_t j = out.size(); j-- > chunk_first;) {
      LInstr &x = out[j];
      if (x.flags & fWrite) {
        if (!live_clv[x.parent]) {
          x.flags &= ~fWrite;
          ++st.stores_dropped;
        }
        live_clv[x.parent] = 0;
      }
      if (x.flags & fWriteS) {
        if (!live_sc[x.pscale]) x.flags &= ~fWriteS;
        live_sc[x.pscale] = 0;
      }
      const bool loads_c1 = !(x.flags & (fTip1 | fNop));
      if (loads_c1 && x.c1 != kNoClv) live_clv[x.c1] = 1;
      if ((x.flags & fCnt1) && x.c1scale >= 0) live_sc[x.c1scale] = 1;
      if (x.flags & fLoadV2) live_clv[x.c2] = 1;
      if (x.flags & fCnt2M) live_sc[x.c2scale] = 1;
    }
  }
  out_chunk_off.push_back((unsigned)out.size());
  if (stats) *stats = st;
}

// ---- subtree groups --------------------------------------------------------------------------------
// A post-order traversal is a chain of dependent instructions only along each root-to-tip path:
// disjoint subtrees are independent programs.  On a small shard (few elements per SM) the walk is
// bound by the latency of one instruction, not by throughput, so the engine runs G groups of
// subtrees side by side (one `blockIdx.y` each, as the chunks of a placement sweep) and then the
// operations that join them.  Every operation computes the same value from the same operands as
// in array order, so the results are identical bit for bit; what changes is only which operations
// follow one another.
//
// The rearrangement is offered only for programs that are plainly forests: every CLV / scale
// buffer is written at most once, never read before it is written in the same program, and read
// by at most one later operation (with the scale buffer its producer wrote).  Anything else
// (in-place updates, shared children, repeated evaluations) keeps its order.
struct ForestInfo {
  std::vector<int>      consumer;  // the operation that reads this operation's parent CLV, -1: none
  std::vector<unsigned> weight;    // operations in the subtree below and including this one
  std::vector<char>     joins;     // evaluations and whatever depends on them: never in a group
};

inline bool analyse_forest(const std::vector<ROp> &ops, unsigned tips, ForestInfo &f) {
  const size_t n = ops.size();
  unsigned     max_clv = tips;
  int          max_sc = 0;
  for (const ROp &r : ops) {
    if (r.parent != kNoClv) max_clv = std::max(max_clv, r.parent + 1);
    if (r.c1 != kNoClv) max_clv = std::max(max_clv, r.c1 + 1);
    if (r.c2 != kNoClv) max_clv = std::max(max_clv, r.c2 + 1);
    max_sc = std::max(max_sc, std::max(r.pscale, std::max(r.c1scale, r.c2scale)) + 1);
  }
  std::vector<int>  wclv(max_clv, -1), wsc((size_t)max_sc, -1);  // the operation that wrote a buffer
  std::vector<char> xclv(max_clv, 0), xsc((size_t)max_sc, 0);    // read as it was before the program
  f.consumer.assign(n, -1);
  f.weight.assign(n, 1u);
  f.joins.assign(n, 0);
  for (size_t i = 0; i < n; ++i) {
    const ROp &r = ops[i];
    auto       read = [&](unsigned c, int s) -> bool {
      int prod = -1;
      if (c != kNoClv && c >= tips) {
        prod = wclv[c];
        if (prod < 0) {
          xclv[c] = 1;
        } else {
          if (f.consumer[prod] != -1) return false;  // a shared child: not a forest
          f.consumer[prod] = (int)i;
          f.weight[i] += f.weight[prod];
          if (f.joins[prod]) f.joins[i] = 1;
        }
      }
      if (s >= 0) {
        if (wsc[s] < 0)
          xsc[s] = 1;
        else if (wsc[s] != prod)
          return false;  // counts that did not come with the CLV
      }
      return true;
    };
    if (!read(r.c1, r.c1scale)) return false;
    if (!(r.flags & rLoadOnly) && !read(r.c2, r.c2scale)) return false;
    if ((r.flags & (rEval | rLoadOnly)) || !(r.flags & rWrite)) f.joins[i] = 1;
    if (r.flags & rWrite) {
      if (r.parent == kNoClv || r.parent < tips) return false;
      if (wclv[r.parent] >= 0 || xclv[r.parent]) return false;  // rewritten, or read before written
      wclv[r.parent] = (int)i;
      if (r.pscale >= 0) {
        if (wsc[r.pscale] >= 0 || xsc[r.pscale]) return false;
        wsc[r.pscale] = (int)i;
      }
    }
  }
  return true;
}

// Whole subtrees of at most `cap` operations are dealt to `n_groups` groups (largest first, to the
// least loaded group); the operations above them join.  group[i] = the group of operation i, or -1.
// Returns the number of groups in use (0: nothing to deal); `longest` = operations of the fullest
// group, `n_join` = joining operations.
inline unsigned assign_subtree_groups(const ForestInfo &f, unsigned n_groups, unsigned cap, std::vector<int> &group,
                                      unsigned &longest, unsigned &n_join) {
  const size_t n = f.consumer.size();
  group.assign(n, -1);
  std::vector<int>                           unit(n, -1);
  std::vector<std::pair<unsigned, unsigned>> units;  // (operations, unit id)
  n_join = 0;
  for (size_t i = n; i-- > 0;) {
    const int c = f.consumer[i];
    if (c >= 0 && unit[c] >= 0) {
      unit[i] = unit[c];
    } else if (!f.joins[i] && f.weight[i] <= cap) {
      unit[i] = (int)units.size();
      units.emplace_back(f.weight[i], (unsigned)units.size());
    } else {
      ++n_join;
    }
  }
  longest = 0;
  if (units.size() < 2 || n_groups < 2) return 0;
  std::sort(units.begin(), units.end(), [](const std::pair<unsigned, unsigned> &a, const std::pair<unsigned, unsigned> &b) {
    return a.first != b.first ? a.first > b.first : a.second < b.second;
  });
  const unsigned        g_used = (unsigned)std::min<size_t>(n_groups, units.size());
  std::vector<unsigned> load(g_used, 0u), of_unit(units.size(), 0u);
  for (const auto &u : units) {
    unsigned best = 0;
    for (unsigned g = 1; g < g_used; ++g)
      if (load[g] < load[best]) best = g;
    load[best] += u.first;
    of_unit[u.second] = best;
  }
  for (unsigned g = 0; g < g_used; ++g) longest = std::max(longest, load[g]);
  for (size_t i = 0; i < n; ++i)
    if (unit[i] >= 0) group[i] = (int)of_unit[unit[i]];
  return g_used;
}

// the operations of each group (array order kept inside a group), then the joining operations
inline void split_by_group(const std::vector<ROp> &ops, const std::vector<int> &group, unsigned g_used,
                           std::vector<ROp> &grouped, std::vector<unsigned> &group_off, std::vector<ROp> &join) {
  group_off.assign(g_used + 1, 0u);
  for (size_t i = 0; i < ops.size(); ++i)
    if (group[i] >= 0) ++group_off[(size_t)group[i] + 1];
  for (unsigned g = 0; g < g_used; ++g) group_off[g + 1] += group_off[g];
  grouped.resize(group_off[g_used]);
  join.clear();
  std::vector<unsigned> at(group_off.begin(), group_off.end() - 1);
  for (size_t i = 0; i < ops.size(); ++i) {
    if (group[i] >= 0)
      grouped[at[(size_t)group[i]]++] = ops[i];
    else
      join.push_back(ops[i]);
  }
}

// lower the groups (as independent chunks) and the joining operations.  With discard_writes (a
// lazily materialised evaluation) a group keeps, besides the stores it reads back itself, the
// buffers the joining operations read.
inline void lower_grouped(const std::vector<ROp> &grouped, const std::vector<unsigned> &group_off,
                          const std::vector<ROp> &join, const LowerOptions &opt, std::vector<LInstr> &out_groups,
                          std::vector<unsigned> &out_group_off, std::vector<LInstr> &out_join, LowerStats *stats) {
  LowerOptions      gopt = opt;
  std::vector<char> scr_clv, scr_sc;
  if (opt.discard_writes) {
    unsigned max_clv = 0;
    int      max_sc = 0;
    for (const ROp &r : grouped) {
      if (r.parent != kNoClv) max_clv = std::max(max_clv, r.parent + 1);
      max_sc = std::max(max_sc, r.pscale + 1);
    }
    scr_clv.assign(max_clv, 1);
    scr_sc.assign((size_t)max_sc, 1);
    auto needed = [&](unsigned c, int sc) {
      if (c != kNoClv && c < scr_clv.size()) scr_clv[c] = 0;
      if (sc >= 0 && (size_t)sc < scr_sc.size()) scr_sc[(size_t)sc] = 0;
    };
    for (const ROp &r : join) {
      needed(r.c1, r.c1scale);
      if (!(r.flags & rLoadOnly)) needed(r.c2, r.c2scale);
    }
    gopt.discard_writes = false;
    gopt.scratch_clv = &scr_clv;
    gopt.scratch_scaler = &scr_sc;
  }
  LowerStats a, b;
  lower_program(grouped, group_off, gopt, out_groups, out_group_off, &a);
  std::vector<unsigned> none;
  lower_program(join, std::vector<unsigned>(), opt, out_join, none, &b);
  if (stats) {
    stats->loadv = a.loadv + b.loadv;
    stats->stores_dropped = a.stores_dropped + b.stores_dropped;
    stats->forwarded = a.forwarded + b.forwarded;
  }
}

}  // namespace rdk
