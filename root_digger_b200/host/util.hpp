// util.hpp -- option / parameter types of the model_t mirror (same names and
// meaning as the reference's src/util.hpp:35-125; CLI-only types are omitted).
#ifndef RD_HOST_UTIL_HPP_
#define RD_HOST_UTIL_HPP_

#include <cstddef>
#include <cstdint>
#include <vector>

typedef std::vector<double> model_params_t;

enum class param_type { emperical, estimate, equal, user };
enum class rate_category { MEDIAN, MEAN, FREE };

struct ratehet_opts_t {
  ratehet_opts_t() = default;
  ratehet_opts_t(size_t rc)
      : type{param_type::estimate}, rate_category_type{rate_category::MEAN}, rate_cats{rc},
        alpha_init{false}, alpha{1.0} {}
  param_type    type = param_type::estimate;
  rate_category rate_category_type = rate_category::MEAN;
  size_t        rate_cats = 0;
  bool          alpha_init = false;
  double        alpha = 1.0;
};

enum class initial_root_strategy_t { random, midpoint, modified_mad };

struct partition_parameters_t {
  model_params_t subst_rates;
  model_params_t freqs;
  model_params_t gamma_alpha;
  model_params_t gamma_weights;
};

struct rd_result_t {
  size_t root_id;
  double llh;
  double alpha;
};

struct dlh_t {
  double lh;
  double dlh;
};

#endif
