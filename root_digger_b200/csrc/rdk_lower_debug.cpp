// rdk_lower_debug.cpp -- introspection of the lowering (rdk_lower.hpp): recorded operations in,
// the instructions the program kernel would walk out, as plain integer tables.  Pure host code;
// tests/test_lowering.py interprets the result symbolically to check the data flow (which CLV
// is in the warps' registers, which stores were dropped) without a GPU.
#include "../../include/rdk.h"
#include "rdk_lower.hpp"

#include <vector>

namespace {
// every instruction names the recorded operation its P slots came from (what lets the engine
// refresh the slots of a kept program): check the mapping against the slots the lowering wrote
bool sources_consistent(const std::vector<rdk::ROp> &ops, const std::vector<rdk::LInstr> &low) {
  using namespace rdk;
  for (const LInstr &x : low) {
    if (x.flags & (fLoadV | fNop)) {
      if (x.src != kNoSrc) return false;
      continue;
    }
    if (x.src >= ops.size()) return false;
    const ROp &r = ops[x.src];
    if (x.pm1 != (x.swapped ? r.pm2 : r.pm1) || x.pm2 != (x.swapped ? r.pm1 : r.pm2)) return false;
  }
  return true;
}
}  // namespace

extern "C" int rdk_debug_lower_program(unsigned int tips, unsigned int n_ops, const int *ops,
                                       unsigned int n_chunks, const unsigned int *chunk_off,
                                       int discard_writes, const unsigned char *scratch_clv,
                                       unsigned int n_scratch_clv, int *out, unsigned int out_cap,
                                       unsigned int *out_chunk_off) {
  using namespace rdk;
  std::vector<ROp> rops(n_ops);
  for (unsigned i = 0; i < n_ops; ++i) {
    const int *f = ops + 10 * (size_t)i;
    ROp       &r = rops[i];
    r.parent = (unsigned)f[0];
    r.pscale = f[1];
    r.c1 = (unsigned)f[2];
    r.c2 = (unsigned)f[3];
    r.c1scale = f[4];
    r.c2scale = f[5];
    r.pm1 = (unsigned)f[6];
    r.pm2 = (unsigned)f[7];
    r.flags = (unsigned)f[8];
    r.slot = (unsigned)f[9];
    r.id = i;
  }
  std::vector<unsigned> coff;
  if (n_chunks > 1 && chunk_off) coff.assign(chunk_off, chunk_off + n_chunks + 1);
  std::vector<char> scratch;
  if (scratch_clv) scratch.assign(scratch_clv, scratch_clv + n_scratch_clv);
  LowerOptions opt;
  opt.tips = tips;
  opt.discard_writes = discard_writes != 0;
  opt.scratch_clv = scratch_clv ? &scratch : nullptr;
  std::vector<LInstr>   low;
  std::vector<unsigned> lchunk;
  lower_program(rops, coff, opt, low, lchunk);
  if (!sources_consistent(rops, low)) {
    rdk_errno = RDK_ERROR_PARAM;
    return -2;
  }
  if (low.size() > out_cap) {
    rdk_errno = RDK_ERROR_PARAM;
    return -1;
  }
  for (size_t i = 0; i < low.size(); ++i) {
    int          *o = out + 10 * i;
    const LInstr &x = low[i];
    o[0] = (int)x.flags;
    o[1] = (int)x.parent;
    o[2] = x.pscale;
    o[3] = (int)x.c1;
    o[4] = x.c1scale;
    o[5] = (int)x.c2;
    o[6] = (int)x.pm1;
    o[7] = (int)x.pm2;
    o[8] = (int)x.slot;
    o[9] = x.c2scale;
  }
  if (out_chunk_off)
    for (size_t c = 0; c < lchunk.size(); ++c) out_chunk_off[c] = lchunk[c];
  return (int)low.size();
}

// The subtree-group rearrangement (rdk_lower.hpp): operations in, the lowered groups followed by
// the lowered joining program out (10 ints per instruction, as above).  out_group_off receives
// n_groups_used + 1 instruction offsets, *n_group_instr the instructions of all groups.  Returns
// the number of groups in use (0: the program keeps its order, nothing is written), -1 on overflow.
extern "C" int rdk_debug_lower_grouped(unsigned int tips, unsigned int n_ops, const int *ops, unsigned int n_groups,
                                       unsigned int cap, int discard_writes, int *out, unsigned int out_cap,
                                       unsigned int *out_group_off, unsigned int *n_group_instr,
                                       unsigned int *n_total_instr) {
  using namespace rdk;
  std::vector<ROp> rops(n_ops);
  for (unsigned i = 0; i < n_ops; ++i) {
    const int *f = ops + 10 * (size_t)i;
    ROp       &r = rops[i];
    r.parent = (unsigned)f[0];
    r.pscale = f[1];
    r.c1 = (unsigned)f[2];
    r.c2 = (unsigned)f[3];
    r.c1scale = f[4];
    r.c2scale = f[5];
    r.pm1 = (unsigned)f[6];
    r.pm2 = (unsigned)f[7];
    r.flags = (unsigned)f[8];
    r.slot = (unsigned)f[9];
    r.id = i;
  }
  ForestInfo fi;
  if (!analyse_forest(rops, tips, fi)) return 0;
  std::vector<int> group;
  unsigned         longest = 0, n_join = 0;
  const unsigned   used = assign_subtree_groups(fi, n_groups, cap, group, longest, n_join);
  if (used < 2) return 0;
  std::vector<ROp>      grouped, join;
  std::vector<unsigned> goff, lgoff;
  split_by_group(rops, group, used, grouped, goff, join);
  LowerOptions opt;
  opt.tips = tips;
  opt.discard_writes = discard_writes != 0;
  std::vector<LInstr> lg, lj;
  lower_grouped(grouped, goff, join, opt, lg, lgoff, lj, nullptr);
  if (!sources_consistent(rops, lg) || !sources_consistent(rops, lj)) {
    rdk_errno = RDK_ERROR_PARAM;
    return -2;
  }
  if (lg.size() + lj.size() > out_cap) {
    rdk_errno = RDK_ERROR_PARAM;
    return -1;
  }
  size_t k = 0;
  for (const std::vector<LInstr> *vec : {&lg, &lj})
    for (const LInstr &x : *vec) {
      int *o = out + 10 * k++;
      o[0] = (int)x.flags;
      o[1] = (int)x.parent;
      o[2] = x.pscale;
      o[3] = (int)x.c1;
      o[4] = x.c1scale;
      o[5] = (int)x.c2;
      o[6] = (int)x.pm1;
      o[7] = (int)x.pm2;
      o[8] = (int)x.slot;
      o[9] = x.c2scale;
    }
  for (size_t c = 0; c < lgoff.size(); ++c) out_group_off[c] = lgoff[c];
  *n_group_instr = (unsigned)lg.size();
  *n_total_instr = (unsigned)(lg.size() + lj.size());
  return (int)used;
}
