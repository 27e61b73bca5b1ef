// partition_file.hpp -- RAxML-NG style partition files and model strings
// (behavioural reference src/msa.cpp:364-587, option types
// src/util.hpp:37-100, known answers test/src/msa.cpp:40-283).
//
//   <MODEL> , <PARTITION_NAME> = <BEGIN>-<END> [, <BEGIN>-<END>]*
//   MODEL  := SUBST ( '+' OPTION )*
//   OPTION := F | FC | FO | FE | FU{f/f/f/f}        base frequencies
//           | I | IO | IC | IU{p}                    invariant sites
//           | G | Gn | Gn{alpha} | GA                discrete Gamma (mean; GA = median)
//           | Rn | Rn{r/..}{w/..}                    free rates (values ignored)
//           | ASC_LEWIS | ASC_FELS{w} | ASC_STAM{w/..}
//           | M...                                   ignored
//
// Site ranges are 1-based and inclusive.  The parser is a hand-written
// recursive-descent scanner over a bounds-checked cursor; every malformed
// input raises std::runtime_error (the reference's tests only require a throw).
#ifndef RD_HOST_PARTITION_FILE_HPP_
#define RD_HOST_PARTITION_FILE_HPP_

#include "util.hpp"

#include <string>
#include <utility>
#include <vector>

struct freq_opts_t {
  param_type          type = param_type::emperical;
  std::vector<double> user_freqs;  // FU{...}: kept (the reference discards them)
};

struct invar_opts_t {
  param_type type = param_type::estimate;
  bool       present = false;  // +I appeared in the model string
  double     user_prop = 0.0;
};

enum class asc_bias_type { none, lewis, fels, stam };

struct asc_bias_opts_t {
  asc_bias_type       type = asc_bias_type::none;
  double              fels_weight = 0.0;
  std::vector<double> stam_weights;
};

struct model_info_t {
  size_t          states = 4;
  std::string     subst_str;
  freq_opts_t     freq_opts;
  invar_opts_t    invar_opts;
  ratehet_opts_t  ratehet_opts;
  asc_bias_opts_t asc_opts;
};

struct partition_info_t {
  std::vector<std::pair<size_t, size_t>> parts;  // [first, second], 1-based inclusive
  std::string                            model_name;
  std::string                            partition_name;
  model_info_t                           model;
  size_t                                 sites() const;
};

typedef std::vector<partition_info_t> msa_partitions_t;

model_info_t     parse_model_info(const std::string &model_string);
partition_info_t parse_partition_info(const std::string &line);
msa_partitions_t parse_partition_text(const std::string &text);  // one partition per non-blank line
msa_partitions_t parse_partition_file(const std::string &filename);

#endif
