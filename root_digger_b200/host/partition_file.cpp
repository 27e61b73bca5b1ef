// partition_file.cpp -- see partition_file.hpp.
#include "partition_file.hpp"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace {

[[noreturn]] void bad(const std::string &what) { throw std::runtime_error("partition file: " + what); }

// bounds-checked cursor over one line / model string
class cursor_t {
public:
  explicit cursor_t(const std::string &s) : _s(s) {}

  bool   done() const { return _i >= _s.size(); }
  char   peek() const { return done() ? '\0' : _s[_i]; }
  char   lower() const { return (char)std::tolower((unsigned char)peek()); }
  size_t pos() const { return _i; }
  void   advance() {
    if (!done()) ++_i;
  }
  void skip_space() {
    while (!done() && std::isspace((unsigned char)_s[_i])) ++_i;
  }
  bool accept(char c) {
    if (std::tolower((unsigned char)peek()) != std::tolower((unsigned char)c) || done()) return false;
    ++_i;
    return true;
  }
  // next non-blank character must be c (case-insensitive); trailing blanks are eaten too
  void expect(char c) {
    skip_space();
    if (!accept(c))
      bad(std::string("expected '") + c + "' but found '" + (done() ? std::string("end of line") : std::string(1, peek())) +
          "'");
    skip_space();
  }
  template <class Pred> std::string take_while(Pred p) {
    size_t b = _i;
    while (!done() && p((unsigned char)_s[_i])) ++_i;
    return _s.substr(b, _i - b);
  }
  std::string since(size_t b) const { return _s.substr(b, _i - b); }

  size_t unsigned_integer(const char *what) {
    std::string d = take_while([](unsigned char c) { return std::isdigit(c) != 0; });
    if (d.empty()) bad(std::string("expected ") + what);
    if (d.size() > 18) bad(std::string(what) + " is too large");
    return (size_t)std::strtoull(d.c_str(), nullptr, 10);
  }
  // digits [. digits] [e [+-] digits]
  double real(const char *what) {
    size_t b = _i;
    take_while([](unsigned char c) { return std::isdigit(c) != 0; });
    if (_i == b) bad(std::string("expected ") + what);
    if (peek() == '.') {
      advance();
      take_while([](unsigned char c) { return std::isdigit(c) != 0; });
    }
    if (lower() == 'e') {
      size_t save = _i;
      advance();
      if (peek() == '+' || peek() == '-') advance();
      size_t d = _i;
      take_while([](unsigned char c) { return std::isdigit(c) != 0; });
      if (_i == d) _i = save;  // a bare 'e' is not part of the number
    }
    return std::strtod(_s.substr(b, _i - b).c_str(), nullptr);
  }
  // '{' real ('/' real)* '}'
  std::vector<double> braced_list(const char *what) {
    std::vector<double> v;
    expect('{');
    for (;;) {
      v.push_back(real(what));
      skip_space();
      if (accept('}')) break;
      expect('/');
    }
    return v;
  }

private:
  const std::string &_s;
  size_t             _i = 0;
};

freq_opts_t freq_options(cursor_t &c) {  // after 'F'
  freq_opts_t f;
  switch (c.lower()) {
    case 'c': c.advance(); f.type = param_type::emperical; break;
    case 'o': c.advance(); f.type = param_type::estimate; break;
    case 'e': c.advance(); f.type = param_type::equal; break;
    case 'u':
      c.advance();
      f.type = param_type::user;
      f.user_freqs = c.braced_list("a base frequency");
      break;
    default: f.type = param_type::emperical; break;
  }
  return f;
}

invar_opts_t invar_options(cursor_t &c) {  // after 'I'
  invar_opts_t v;
  v.present = true;
  switch (c.lower()) {
    case 'o': c.advance(); v.type = param_type::estimate; break;
    case 'c': c.advance(); v.type = param_type::emperical; break;
    case 'u': {
      c.advance();
      v.type = param_type::user;
      auto l = c.braced_list("a proportion of invariant sites");
      if (l.size() != 1) bad("+IU takes exactly one value");
      v.user_prop = l[0];
      break;
    }
    default: v.type = param_type::estimate; break;
  }
  return v;
}

ratehet_opts_t gamma_options(cursor_t &c) {  // after 'G'
  ratehet_opts_t r;
  r.type = param_type::estimate;
  r.rate_category_type = rate_category::MEAN;
  r.rate_cats = 4;
  if (c.lower() == 'a') {  // +GA: median category rates
    c.advance();
    r.rate_category_type = rate_category::MEDIAN;
    return r;
  }
  if (std::isdigit((unsigned char)c.peek())) {
    r.rate_cats = c.unsigned_integer("a number of rate categories");
    if (c.peek() == '{') {
      auto l = c.braced_list("a Gamma shape");
      if (l.size() != 1) bad("+Gn{alpha} takes exactly one value");
      r.alpha = l[0];
      r.alpha_init = true;
      r.type = param_type::user;
    }
  }
  return r;
}

ratehet_opts_t free_rate_options(cursor_t &c) {  // after 'R'
  ratehet_opts_t r;
  r.type = param_type::estimate;
  r.rate_category_type = rate_category::FREE;
  r.rate_cats = c.unsigned_integer("a number of rate categories");
  // user rates / weights are accepted and ignored (not supported by RootDigger)
  while (c.peek() == '{') c.braced_list("a rate or weight");
  return r;
}

asc_bias_opts_t asc_options(cursor_t &c) {  // after 'A'
  asc_bias_opts_t a;
  c.expect('S');
  c.expect('C');
  c.expect('_');
  std::string word = c.take_while([](unsigned char ch) { return std::isalpha(ch) != 0; });
  if (word.empty()) bad("expected an ascertainment bias correction type");
  switch (std::tolower((unsigned char)word[0])) {
    case 'l': a.type = asc_bias_type::lewis; break;
    case 'f': {
      a.type = asc_bias_type::fels;
      auto l = c.braced_list("a weight");
      if (l.size() != 1) bad("+ASC_FELS takes exactly one weight");
      a.fels_weight = l[0];
      break;
    }
    case 's':
      a.type = asc_bias_type::stam;
      a.stam_weights = c.braced_list("a weight");
      break;
    default: bad("unknown ascertainment bias correction '" + word + "'");
  }
  return a;
}

bool model_name_char(unsigned char c) {
  return std::isalnum(c) || c == '+' || c == '{' || c == '}' || c == '/' || c == '.' || c == '_' || c == ':';
}

}  // namespace

size_t partition_info_t::sites() const {
  size_t n = 0;
  for (auto &r : parts) n += r.second - r.first + 1;
  return n;
}

model_info_t parse_model_info(const std::string &model_string) {
  model_info_t m;
  cursor_t     c(model_string);
  c.skip_space();
  m.subst_str = c.take_while([](unsigned char ch) { return std::isalnum(ch) || ch == '_' || ch == ':'; });
  if (m.subst_str.empty()) bad("the model string has no substitution model name");
  for (c.skip_space(); !c.done(); c.skip_space()) {
    c.expect('+');
    const char opt = c.lower();
    c.advance();
    switch (opt) {
      case 'f': m.freq_opts = freq_options(c); break;
      case 'i': m.invar_opts = invar_options(c); break;
      case 'g': m.ratehet_opts = gamma_options(c); break;
      case 'r': m.ratehet_opts = free_rate_options(c); break;
      case 'a': m.asc_opts = asc_options(c); break;
      case 'm':  // +M...: not supported by RootDigger, skipped up to the next option
        c.take_while([](unsigned char ch) { return ch != '+'; });
        break;
      default: bad(std::string("unknown model option '+") + opt + "'");
    }
  }
  return m;
}

partition_info_t parse_partition_info(const std::string &line) {
  partition_info_t p;
  cursor_t         c(line);
  c.skip_space();
  p.model_name = c.take_while(model_name_char);
  if (p.model_name.empty()) bad("the partition has no model name");
  p.model = parse_model_info(p.model_name);
  c.expect(',');
  p.partition_name = c.take_while([](unsigned char ch) { return std::isalnum(ch) || ch == '_'; });
  c.expect('=');
  for (;;) {
    c.skip_space();
    const size_t begin = c.unsigned_integer("the first site of a range");
    if (c.peek() == ',') {  // "<site>," : a single column
      p.parts.emplace_back(begin, begin);
      c.advance();
      continue;
    }
    c.expect('-');
    const size_t end = c.unsigned_integer("the last site of a range");
    if (end < begin) bad("the range " + std::to_string(begin) + "-" + std::to_string(end) + " of partition '" +
                         p.partition_name + "' ends before it begins");
    p.parts.emplace_back(begin, end);
    c.skip_space();
    if (!c.accept(',')) break;
  }
  c.skip_space();
  if (!c.done()) bad(std::string("unexpected '") + c.peek() + "' after the site ranges");
  return p;
}

msa_partitions_t parse_partition_text(const std::string &text) {
  msa_partitions_t   parts;
  std::istringstream in(text);
  for (std::string line; std::getline(in, line);) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    bool blank = true;
    for (char ch : line)
      if (!std::isspace((unsigned char)ch)) blank = false;
    if (blank) continue;
    parts.push_back(parse_partition_info(line));
  }
  return parts;
}

msa_partitions_t parse_partition_file(const std::string &filename) {
  std::ifstream f(filename);
  if (!f) throw std::runtime_error("Failed to open the partition file " + filename);
  std::stringstream ss;
  ss << f.rdbuf();
  return parse_partition_text(ss.str());
}
