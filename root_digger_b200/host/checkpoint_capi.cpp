// checkpoint_capi.cpp -- C wrappers around checkpoint_t for ctypes (tests, the Python
// launcher).  Return 1 on success, 0 on failure (message from rdh_last_error()).
#include "checkpoint.hpp"

#include <cstring>
#include <memory>

void rdh_set_error(const std::string &s);

#define RDH_TRY(body)                                                                              \
  try {                                                                                            \
    body                                                                                           \
  } catch (const std::exception &e) {                                                              \
    rdh_set_error(e.what());                                                                       \
    return 0;                                                                                      \
  }

namespace {
checkpoint_t &C(void *c) { return *reinterpret_cast<checkpoint_t *>(c); }

// the option fields the wrappers move (the rest keep their defaults)
void fill(cli_options_t &o, const char *msa, const char *tree, const char *prefix, const char *model_string,
          const unsigned long long *rate_cats, unsigned n_rate_cats, unsigned long long seed,
          unsigned long long min_roots, unsigned long long threads, int exhaustive, int early_stop,
          int strategy) {
  o.msa_filename = msa ? msa : "";
  o.tree_filename = tree ? tree : "";
  o.prefix = prefix ? prefix : "";
  o.model_string = model_string ? model_string : "";
  o.rate_cats.clear();
  for (unsigned i = 0; i < n_rate_cats; ++i) o.rate_cats.emplace_back((size_t)rate_cats[i]);
  o.seed = seed;
  o.min_roots = (size_t)min_roots;
  o.threads = (size_t)threads;
  o.exhaustive = exhaustive != 0;
  o.early_stop = initialized_flag_t::from_raw(early_stop);
  o.initial_root_strategy = (initial_root_strategy_t)strategy;
}
}  // namespace

extern "C" void *rdh_ckp_open(const char *prefix) {
  try {
    return prefix ? new checkpoint_t(std::string(prefix)) : new checkpoint_t();
  } catch (const std::exception &e) {
    rdh_set_error(e.what());
    return nullptr;
  }
}

extern "C" void rdh_ckp_close(void *c) { delete reinterpret_cast<checkpoint_t *>(c); }

extern "C" int rdh_ckp_existing(void *c) { return C(c).existing_checkpoint() ? 1 : 0; }

extern "C" int rdh_ckp_filename(void *c, char *out, unsigned cap) {
  std::string s = C(c).get_filename();
  if (s.size() + 1 > cap) return 0;
  memcpy(out, s.c_str(), s.size() + 1);
  return 1;
}

extern "C" int rdh_ckp_save_options(void *c, const char *msa, const char *tree, const char *prefix,
                                    const char *model_string, const unsigned long long *rate_cats,
                                    unsigned n_rate_cats, unsigned long long seed,
                                    unsigned long long min_roots, unsigned long long threads, int exhaustive,
                                    int early_stop, int strategy) {
  RDH_TRY({
    cli_options_t o;
    fill(o, msa, tree, prefix, model_string, rate_cats, n_rate_cats, seed, min_roots, threads, exhaustive,
         early_stop, strategy);
    C(c).save_options(o);
    return 1;
  })
}

// returns 2 when the stored options equal the given ones (the reference's operator==), 1 when
// they differ, 0 on failure; the stored msa filename / seed / rate_cats count are copied out
extern "C" int rdh_ckp_load_options(void *c, const char *msa, const char *tree, const char *prefix,
                                    const char *model_string, const unsigned long long *rate_cats,
                                    unsigned n_rate_cats, unsigned long long seed,
                                    unsigned long long min_roots, unsigned long long threads, int exhaustive,
                                    int early_stop, int strategy, char *msa_out, unsigned cap,
                                    unsigned long long *seed_out, unsigned *n_rate_cats_out) {
  RDH_TRY({
    cli_options_t mine;
    cli_options_t stored;
    fill(mine, msa, tree, prefix, model_string, rate_cats, n_rate_cats, seed, min_roots, threads, exhaustive,
         early_stop, strategy);
    C(c).load_options(stored);
    if (msa_out && stored.msa_filename.size() + 1 <= cap)
      memcpy(msa_out, stored.msa_filename.c_str(), stored.msa_filename.size() + 1);
    if (seed_out) *seed_out = stored.seed;
    if (n_rate_cats_out) *n_rate_cats_out = (unsigned)stored.rate_cats.size();
    return stored == mine ? 2 : 1;
  })
}

// one record: n_parts parameter sets, each 12 rates, 4 freqs, 1 alpha, K weights
extern "C" int rdh_ckp_write(void *c, unsigned long long root_id, double llh, double alpha, unsigned n_parts,
                             const double *rates, const double *freqs, const double *alphas,
                             const double *weights, unsigned K) {
  RDH_TRY({
    std::vector<partition_parameters_t> params(n_parts);
    for (unsigned p = 0; p < n_parts; ++p) {
      params[p].subst_rates.assign(rates + 12 * p, rates + 12 * (p + 1));
      params[p].freqs.assign(freqs + 4 * p, freqs + 4 * (p + 1));
      params[p].gamma_alpha.assign(alphas + p, alphas + p + 1);
      params[p].gamma_weights.assign(weights + (size_t)K * p, weights + (size_t)K * (p + 1));
    }
    C(c).write(rd_result_t{(size_t)root_id, llh, alpha}, params);
    return 1;
  })
}

extern "C" int rdh_ckp_read(void *c, unsigned cap, unsigned long long *ids, double *llh, double *alpha,
                            unsigned *n_parts, unsigned *n_out) {
  RDH_TRY({
    auto res = C(c).read_results();
    *n_out = (unsigned)res.size();
    for (size_t i = 0; i < res.size() && i < cap; ++i) {
      ids[i] = res[i].first.root_id;
      llh[i] = res[i].first.llh;
      alpha[i] = res[i].first.alpha;
      if (n_parts) n_parts[i] = (unsigned)res[i].second.size();
    }
    return 1;
  })
}

// the parameter set `part` of record `index`, flattened as rdh_ckp_write takes it
extern "C" int rdh_ckp_read_params(void *c, unsigned index, unsigned part, double *rates, double *freqs,
                                   double *alpha, double *weights, unsigned K) {
  RDH_TRY({
    auto res = C(c).read_results();
    if (index >= res.size() || part >= res[index].second.size())
      throw std::runtime_error("no such checkpoint record");
    const auto &pp = res[index].second[part];
    // gamma_weights is empty unless the categories are FREE (model_t::random_params)
    if (pp.subst_rates.size() != 12 || pp.freqs.size() != 4 || pp.gamma_alpha.empty() ||
        (pp.gamma_weights.size() != K && !pp.gamma_weights.empty()))
      throw std::runtime_error("unexpected parameter vector sizes in the checkpoint");
    memcpy(rates, pp.subst_rates.data(), 12 * sizeof(double));
    memcpy(freqs, pp.freqs.data(), 4 * sizeof(double));
    *alpha = pp.gamma_alpha[0];
    memset(weights, 0, K * sizeof(double));
    memcpy(weights, pp.gamma_weights.data(), pp.gamma_weights.size() * sizeof(double));
    return 1;
  })
}

extern "C" int rdh_ckp_completed(void *c, unsigned cap, unsigned long long *ids, unsigned *n_out) {
  RDH_TRY({
    auto idx = C(c).completed_indicies();
    *n_out = (unsigned)idx.size();
    for (size_t i = 0; i < idx.size() && i < cap; ++i) ids[i] = idx[i];
    return 1;
  })
}

// 1 = damaged tail, 0 = clean, -1 = failure
extern "C" int rdh_ckp_needs_cleaning(void *c) {
  try {
    return C(c).needs_cleaning() ? 1 : 0;
  } catch (const std::exception &e) {
    rdh_set_error(e.what());
    return -1;
  }
}

extern "C" int rdh_ckp_clean(void *c) {
  RDH_TRY({
    C(c).clean();
    return 1;
  })
}

extern "C" unsigned rdh_ckp_checksum_result(unsigned long long root_id, double llh, double alpha) {
  return checkpoint_checksum(rd_result_t{(size_t)root_id, llh, alpha});
}

extern "C" unsigned rdh_ckp_checksum_params(unsigned n_parts, const double *rates, const double *freqs,
                                            const double *alphas, const double *weights, unsigned K) {
  std::vector<partition_parameters_t> params(n_parts);
  for (unsigned p = 0; p < n_parts; ++p) {
    params[p].subst_rates.assign(rates + 12 * p, rates + 12 * (p + 1));
    params[p].freqs.assign(freqs + 4 * p, freqs + 4 * (p + 1));
    params[p].gamma_alpha.assign(alphas + p, alphas + p + 1);
    params[p].gamma_weights.assign(weights + (size_t)K * p, weights + (size_t)K * (p + 1));
  }
  return checkpoint_checksum(params);
}
