// tree_capi.cpp -- C wrappers around rooted_tree_t for ctypes (tests, bench).
// Everything returns 1 on success, 0 on failure with the message available from
// rdh_last_error().
#include "tree.hpp"

#include <cstring>
#include <string>

static thread_local std::string g_err;

extern "C" const char *rdh_last_error(void) { return g_err.c_str(); }
void                   rdh_set_error(const std::string &s) { g_err = s; }
extern "C" void        rdh_free(void *p) { free(p); }

#define RDH_GUARD(body)                                                                            \
  try {                                                                                            \
    body                                                                                           \
  } catch (const std::exception &e) {                                                              \
    g_err = e.what();                                                                              \
    return 0;                                                                                      \
  }

extern "C" void *rdh_tree_from_newick(const char *text) {
  try {
    return new rooted_tree_t(rooted_tree_t::from_newick(text));
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

extern "C" void *rdh_tree_from_file(const char *path) {
  try {
    return new rooted_tree_t(std::string(path));
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

extern "C" void *rdh_tree_copy(void *t) {
  try {
    return new rooted_tree_t(*reinterpret_cast<rooted_tree_t *>(t));
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

extern "C" void rdh_tree_destroy(void *t) { delete reinterpret_cast<rooted_tree_t *>(t); }

static rooted_tree_t &T(void *t) { return *reinterpret_cast<rooted_tree_t *>(t); }

extern "C" unsigned rdh_tree_tip_count(void *t) { return T(t).tip_count(); }
extern "C" unsigned rdh_tree_inner_count(void *t) { return T(t).inner_count(); }
extern "C" unsigned rdh_tree_branch_count(void *t) { return T(t).branch_count(); }
extern "C" unsigned rdh_tree_root_count(void *t) { return (unsigned)T(t).root_count(); }
extern "C" unsigned rdh_tree_root_clv_index(void *t) { return T(t).root_clv_index(); }
extern "C" int      rdh_tree_root_scaler_index(void *t) { return T(t).root_scaler_index(); }
extern "C" int      rdh_tree_rooted(void *t) { return T(t).rooted() ? 1 : 0; }
extern "C" int      rdh_tree_sanity_check(void *t) { return T(t).sanity_check() ? 1 : 0; }

// root location info: saved branch length, internal flag, label (may be "(null)")
extern "C" int rdh_tree_root_info(void *t, unsigned id, double *saved_brlen, int *is_internal,
                                  char *label, unsigned label_cap) {
  RDH_GUARD({
    auto rl = T(t).root_location((size_t)id);
    if (saved_brlen) *saved_brlen = rl.saved_brlen;
    if (is_internal) *is_internal = rl.is_internal() ? 1 : 0;
    if (label && label_cap) {
      std::string l = rl.label();
      std::strncpy(label, l.c_str(), label_cap - 1);
      label[label_cap - 1] = 0;
    }
    return 1;
  })
}

extern "C" int rdh_tree_root_id_by_label(void *t, const char *label) {
  try {
    return (int)T(t).root_location(std::string(label)).id;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// clv index of the tip with this label, -1 if absent
extern "C" int rdh_tree_tip_index(void *t, const char *label) {
  auto lm = T(t).label_map();
  auto it = lm.find(label);
  return it == lm.end() ? -1 : (int)it->second;
}

// label of the tip with clv index `index`
extern "C" int rdh_tree_tip_label(void *t, unsigned index, char *out, unsigned cap) {
  auto lm = T(t).label_map();
  for (auto &kv : lm)
    if (kv.second == index) {
      std::strncpy(out, kv.first.c_str(), cap - 1);
      out[cap - 1] = 0;
      return 1;
    }
  g_err = "no tip with this index";
  return 0;
}

static int unpack(const rooted_tree_t::op_bundle_t &b, rdk_operation_t *ops, unsigned ops_cap,
                  unsigned *n_ops, unsigned *pm, double *br, unsigned pm_cap, unsigned *n_pm) {
  const auto &o = std::get<0>(b);
  const auto &m = std::get<1>(b);
  const auto &l = std::get<2>(b);
  if (o.size() > ops_cap || m.size() > pm_cap) {
    g_err = "output buffers too small";
    return 0;
  }
  for (size_t i = 0; i < o.size(); ++i) ops[i] = o[i];
  for (size_t i = 0; i < m.size(); ++i) {
    pm[i] = m[i];
    br[i] = l[i];
  }
  *n_ops = (unsigned)o.size();
  *n_pm = (unsigned)m.size();
  return 1;
}

extern "C" int rdh_tree_generate_operations(void *t, unsigned root_id, double ratio,
                                            rdk_operation_t *ops, unsigned ops_cap, unsigned *n_ops,
                                            unsigned *pm, double *br, unsigned pm_cap,
                                            unsigned *n_pm) {
  RDH_GUARD({
    auto rl = T(t).root_location((size_t)root_id);
    rl.brlen_ratio = ratio;
    return unpack(T(t).generate_operations(rl), ops, ops_cap, n_ops, pm, br, pm_cap, n_pm);
  })
}

extern "C" int rdh_tree_generate_derivative_operations(void *t, unsigned root_id, double ratio,
                                                       rdk_operation_t *op, unsigned *pm /*2*/,
                                                       double *br /*2*/) {
  RDH_GUARD({
    auto rl = T(t).root_location((size_t)root_id);
    rl.brlen_ratio = ratio;
    auto r = T(t).generate_derivative_operations(rl);
    *op = std::get<0>(r);
    for (int i = 0; i < 2; ++i) {
      pm[i] = std::get<1>(r)[i];
      br[i] = std::get<2>(r)[i];
    }
    return 1;
  })
}

extern "C" int rdh_tree_generate_root_update_operations(void *t, unsigned root_id, double ratio,
                                                        rdk_operation_t *ops, unsigned ops_cap,
                                                        unsigned *n_ops, unsigned *pm, double *br,
                                                        unsigned pm_cap, unsigned *n_pm) {
  RDH_GUARD({
    auto rl = T(t).root_location((size_t)root_id);
    rl.brlen_ratio = ratio;
    return unpack(T(t).generate_root_update_operations(rl), ops, ops_cap, n_ops, pm, br, pm_cap,
                  n_pm);
  })
}

extern "C" unsigned rdh_tree_sweep_depth_bound(void *t) { return T(t).sweep_depth_bound(); }

// the directed-CLV placement sweep of root positions [begin, end) from the CURRENT root
// (rooted_tree_t::generate_sweep_operations); flat outputs in the shape
// rdk_sweep_root_placements takes.  pm_off / op_off / root_pos hold placements (+1) entries.
extern "C" int rdh_tree_generate_sweep_operations(void *t, unsigned begin, unsigned end, unsigned clv0,
                                                  int scaler0, unsigned pm0, unsigned extra,
                                                  unsigned *n_placements, unsigned *pm_off,
                                                  unsigned *op_off, unsigned *root_pos,
                                                  unsigned placements_cap, unsigned *pm, double *br,
                                                  unsigned pm_cap, rdk_operation_t *ops,
                                                  unsigned ops_cap) {
  RDH_GUARD({
    auto s = T(t).generate_sweep_operations(begin, end, clv0, scaler0, pm0, extra);
    if (s.root_pos.size() > placements_cap || s.mi.size() > pm_cap || s.ops.size() > ops_cap) {
      g_err = "output buffers too small";
      return 0;
    }
    *n_placements = (unsigned)s.root_pos.size();
    for (size_t i = 0; i < s.pm_off.size(); ++i) pm_off[i] = s.pm_off[i];
    for (size_t i = 0; i < s.op_off.size(); ++i) op_off[i] = s.op_off[i];
    for (size_t i = 0; i < s.root_pos.size(); ++i) root_pos[i] = (unsigned)s.root_pos[i];
    for (size_t i = 0; i < s.mi.size(); ++i) {
      pm[i] = s.mi[i];
      br[i] = s.bl[i];
    }
    for (size_t i = 0; i < s.ops.size(); ++i) ops[i] = s.ops[i];
    return 1;
  })
}

extern "C" int rdh_tree_root_by(void *t, unsigned root_id, double ratio) {
  RDH_GUARD({
    auto rl = T(t).root_location((size_t)root_id);
    rl.brlen_ratio = ratio;
    T(t).root_by(rl);
    return 1;
  })
}

extern "C" int rdh_tree_unroot(void *t) {
  RDH_GUARD({
    T(t).unroot();
    return 1;
  })
}

extern "C" char *rdh_tree_newick(void *t, int annotations) {
  try {
    std::string s = T(t).newick(annotations != 0);
    char       *out = (char *)malloc(s.size() + 1);
    std::memcpy(out, s.c_str(), s.size() + 1);
    return out;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

extern "C" int rdh_tree_annotate_branch(void *t, unsigned root_id, const char *key,
                                        const char *value) {
  RDH_GUARD({
    T(t).annotate_branch((size_t)root_id, key, value);
    return 1;
  })
}

extern "C" int rdh_tree_annotate_lwr(void *t, unsigned root_id, double ratio, double lwr, double llh) {
  RDH_GUARD({
    auto rl = T(t).root_location((size_t)root_id);
    rl.brlen_ratio = ratio;
    T(t).annotate_branch(rl, "LWR", std::to_string(lwr));
    T(t).annotate_lh(rl, llh);
    T(t).annotate_ratio(rl, ratio);
    return 1;
  })
}

// which: 0 = midpoint ranking, 1 = modified MAD ranking; writes root ids
extern "C" int rdh_tree_rank_roots(void *t, int which, unsigned *ids_out, unsigned cap) {
  RDH_GUARD({
    auto r = which == 0 ? T(t).rank_midpoints() : T(t).rank_modified_mad();
    if (r.size() > cap) {
      g_err = "output buffer too small";
      return 0;
    }
    for (size_t i = 0; i < r.size(); ++i) ids_out[i] = (unsigned)r[i].id;
    return 1;
  })
}
