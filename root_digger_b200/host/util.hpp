// util.hpp -- option / parameter types of the model_t mirror (same names and
// meaning as the reference's src/util.hpp:35-125; CLI-only types are omitted).
#ifndef RD_HOST_UTIL_HPP_
#define RD_HOST_UTIL_HPP_

#include <cstddef>
#include <cstdint>
#include <string>
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

  bool operator==(const ratehet_opts_t &o) const {
    return type == o.type && rate_category_type == o.rate_category_type && rate_cats == o.rate_cats &&
           alpha_init == o.alpha_init && alpha == o.alpha;
  }
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

// tri-state flag of the reference (src/util.hpp:127-160): one 4-byte enum on disk
class initialized_flag_t {
public:
  enum class initial_behavior { uninitalized, initialized_true, initialized_false };
  initialized_flag_t() : value(initial_behavior::uninitalized) {}
  initialized_flag_t(const initial_behavior &v) : value(v) {}
  bool operator==(const initialized_flag_t &rhs) const { return rhs.value == value; }
  bool operator!=(const initialized_flag_t &rhs) const { return rhs.value != value; }
  bool initalized() const { return value != initial_behavior::uninitalized; }
  bool convert_with_default(bool default_value) const {
    if (value == initial_behavior::uninitalized) return default_value;
    return value == initial_behavior::initialized_true;
  }
  int32_t raw() const { return (int32_t)value; }
  static initialized_flag_t from_raw(int32_t r) { return initialized_flag_t((initial_behavior)r); }

private:
  initial_behavior value;
};

// the run options the checkpoint header records (reference src/util.hpp:162-215; the
// std::filesystem::path members are plain strings here -- same bytes on disk).  The seed
// default is fixed instead of std::random_device so that runs are reproducible.
struct cli_options_t {
  std::string                 msa_filename;
  std::string                 tree_filename;
  std::string                 prefix;
  std::string                 prefix_dir;
  std::string                 model_filename;
  std::string                 freqs_filename;
  std::string                 partition_filename;
  std::string                 data_type;
  std::string                 model_string;
  std::vector<ratehet_opts_t> rate_cats = {ratehet_opts_t(1)};
  uint64_t                    seed = 0;
  size_t                      min_roots = 1;
  size_t                      threads = 0;
  double                      root_ratio = 0.01;
  double                      abs_tolerance = 1e-7;
  double                      factor = 1e4;
  double                      br_tolerance = 1e-12;
  double                      bfgs_tol = 1e-7;
  unsigned int                states = 4;
  bool                        silent = false;
  bool                        exhaustive = false;
  bool                        echo = false;
  bool                        invariant_sites = false;
  bool                        clean = false;
  initialized_flag_t          early_stop;
  initial_root_strategy_t     initial_root_strategy = initial_root_strategy_t::modified_mad;

  // the reference's comparison (src/util.hpp:193-211): min_roots, silent and clean do not take part
  bool operator==(const cli_options_t &o) const {
    return msa_filename == o.msa_filename && tree_filename == o.tree_filename && prefix == o.prefix &&
           prefix_dir == o.prefix_dir && model_filename == o.model_filename &&
           freqs_filename == o.freqs_filename && partition_filename == o.partition_filename &&
           data_type == o.data_type && model_string == o.model_string && rate_cats == o.rate_cats &&
           seed == o.seed && threads == o.threads && root_ratio == o.root_ratio &&
           abs_tolerance == o.abs_tolerance && factor == o.factor && br_tolerance == o.br_tolerance &&
           bfgs_tol == o.bfgs_tol && states == o.states && exhaustive == o.exhaustive && echo == o.echo &&
           invariant_sites == o.invariant_sites && early_stop == o.early_stop &&
           initial_root_strategy == o.initial_root_strategy;
  }
  bool operator!=(const cli_options_t &o) const { return !(*this == o); }
};

struct dlh_t {
  double lh;
  double dlh;
};

#endif
