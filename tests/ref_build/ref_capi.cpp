// ref_capi.cpp -- C wrappers for ctypes around RootDigger's OWN model_t / rooted_tree_t / msa_t /
// checkpoint_t: this file is compiled together with the reference's unmodified src/model.cpp,
// src/tree.cpp, src/msa.cpp, src/checkpoint.cpp, src/util.cpp (taken from where they lie in the
// reference checkout, never copied) against root_digger_b200/compat/corax/corax.h, i.e. against the
// engine's C ABI (include/rdk.h) or -- for the CPU suite -- the oracle behind the same ABI.
// TEST INFRASTRUCTURE: it proves the drop-in boundary with the reference's own callers
// (tests/test_reference_sources.py) and mirrors the call sequence of src/main.cpp:513-640.
#include "model.hpp"

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <unistd.h>

int __VERBOSE__ = 0;
int __MPI_RANK__ = 0;
int __MPI_NUM_TASKS__ = 1;

namespace {
thread_local std::string g_err;
struct holder_t {
  std::vector<msa_t>            msa;
  std::unique_ptr<model_t>      model;
  std::unique_ptr<checkpoint_t> checkpoint;
  std::string                   ckp_prefix;
  unsigned                      root_count = 0;
};
holder_t &H(void *h) { return *reinterpret_cast<holder_t *>(h); }
}  // namespace

#define REF_TRY(body)                    \
  try {                                  \
    body                                 \
  } catch (const std::exception &e) {    \
    g_err = e.what();                    \
    return 0;                            \
  }

extern "C" const char *rdref_last_error(void) { return g_err.c_str(); }

// src/main.cpp:513-589: alignment (compressed), tree, model, initialize_partitions (uniform
// frequencies when the empirical ones are invalid), initialize
extern "C" void *rdref_create(const char *tree_file, const char *msa_file, unsigned rate_cats, unsigned long long seed,
                              int early_stop, const char *scratch_prefix) {
  try {
    auto h = std::make_unique<holder_t>();
    h->msa.emplace_back(std::string(msa_file));
    for (auto &m : h->msa) m.valid_data();
    rooted_tree_t tree{std::string(tree_file)};
    h->root_count = (unsigned)tree.root_count();
    h->model = std::make_unique<model_t>(tree, h->msa, std::vector<ratehet_opts_t>{ratehet_opts_t{rate_cats}}, false,
                                         (uint64_t)seed, early_stop != 0);
    try {
      h->model->initialize_partitions(h->msa);
    } catch (const invalid_empirical_frequencies_exception &) {
      h->model->initialize_partitions_uniform_freqs(h->msa);
    }
    h->model->initialize();
    h->ckp_prefix = scratch_prefix;
    unlink((h->ckp_prefix + ".ckp").c_str());
    h->checkpoint = std::make_unique<checkpoint_t>(h->ckp_prefix);
    cli_options_t o;
    o.prefix = h->ckp_prefix;
    h->checkpoint->save_options(o);
    return h.release();
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

extern "C" void rdref_destroy(void *h) {
  if (!h) return;
  std::string f = H(h).ckp_prefix + ".ckp";
  delete &H(h);
  unlink(f.c_str());
}

extern "C" unsigned rdref_root_count(void *h) { return H(h).root_count; }

// init_strategy: 0 random, 1 midpoint, 2 modified MAD (src/main.cpp:591-606)
extern "C" int rdref_search(void *h, unsigned min_roots, double root_ratio, double atol, double pgtol, double brtol,
                            double factor, int init_strategy, unsigned *id, double *alpha, double *lh) {
  REF_TRY({
    auto &m = *H(h).model;
    auto  strat = init_strategy == 0   ? initial_root_strategy_t::random
                  : init_strategy == 1 ? initial_root_strategy_t::midpoint
                                       : initial_root_strategy_t::modified_mad;
    m.assign_indicies_by_rank_search(min_roots, root_ratio, 0, 1, strat, *H(h).checkpoint);
    auto r = m.search(min_roots, root_ratio, atol, pgtol, brtol, factor, *H(h).checkpoint);
    *id = (unsigned)r.first.id;
    *alpha = r.first.brlen_ratio;
    *lh = r.second;
    return 1;
  })
}

// src/main.cpp:620-640: every branch optimised; per-root results come from the checkpoint log
extern "C" int rdref_exhaustive(void *h, double atol, double pgtol, double brtol, double factor, unsigned *ids,
                                double *llh, double *alpha, unsigned cap, unsigned *n_out, unsigned *best_id,
                                double *best_lh) {
  REF_TRY({
    auto &m = *H(h).model;
    m.assign_indicies_by_rank_exhaustive(0, 1, *H(h).checkpoint);
    auto r = m.exhaustive_search(atol, pgtol, brtol, factor, *H(h).checkpoint);
    *best_id = (unsigned)r.first.id;
    *best_lh = r.second;
    auto res = H(h).checkpoint->read_results();
    if (res.size() > cap) throw std::runtime_error("output buffers too small");
    for (size_t i = 0; i < res.size(); ++i) {
      ids[i] = (unsigned)res[i].first.root_id;
      llh[i] = res[i].first.llh;
      alpha[i] = res[i].first.alpha;
    }
    *n_out = (unsigned)res.size();
    return 1;
  })
}

extern "C" int rdref_all_root_lh(void *h, double *out, unsigned cap) {
  REF_TRY({
    auto v = H(h).model->compute_all_root_lh();
    if (v.size() > cap) throw std::runtime_error("output buffer too small");
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return 1;
  })
}
