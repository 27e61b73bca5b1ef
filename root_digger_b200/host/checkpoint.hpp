// checkpoint.hpp -- result log used by search()/exhaustive_search().
//
// The reference appends {root_id, llh, alpha} + the partition parameters of
// every finished start root / branch to "<prefix>.ckp" and rank 0 reads them
// all back (src/model.cpp:1107,1120,1215,1238; src/checkpoint.cpp).  The
// on-disk format (Adler-32 framed records, fcntl locks) is outside the hot
// path; this mirror keeps the same member names on
// an in-memory log, which is all model_t needs.
#ifndef RD_HOST_CHECKPOINT_HPP_
#define RD_HOST_CHECKPOINT_HPP_

#include "util.hpp"

#include <mutex>
#include <utility>
#include <vector>

class checkpoint_t {
public:
  typedef std::pair<rd_result_t, std::vector<partition_parameters_t>> record_t;

  void write(const rd_result_t &result, const std::vector<partition_parameters_t> &params) {
    std::lock_guard<std::mutex> lk(_mu);
    _records.emplace_back(result, params);
  }
  std::vector<record_t> read_results() const { return _records; }
  std::vector<rd_result_t> current_progress() const {
    std::vector<rd_result_t> r;
    for (auto &rec : _records) r.push_back(rec.first);
    return r;
  }
  std::vector<size_t> completed_indicies() const {
    std::vector<size_t> r;
    for (auto &rec : _records) r.push_back(rec.first.root_id);
    return r;
  }
  void clear() { _records.clear(); }

private:
  std::vector<record_t> _records;
  std::mutex            _mu;
};

#endif
