// rdk_lower_debug.cpp -- introspection of the lowering (rdk_lower.hpp): recorded operations in,
// the instructions the program kernel would walk out, as plain integer tables.  Pure host code;
// tests/test_lowering.py interprets the result symbolically to check the data flow (which CLV
// is in the warps' registers, which stores were dropped) without a GPU.
#include "../../include/rdk.h"
#include "rdk_lower.hpp"

#include <vector>

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
