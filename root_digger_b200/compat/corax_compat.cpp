// corax_compat.cpp -- host-only half of corax/corax.h: the unrooted-tree module and the
// alignment readers RootDigger's sources call (src/tree.cpp, src/msa.cpp).  The likelihood half of
// that header is the engine's C ABI itself (name mappings, no code here).
//
// The tree module is the engine host's own utree code (host/tree.cpp, namespace rd: newick
// parsing + unrooting, index numbering -- SURVEY Appendix A-7) re-expressed on coraxlib's plain C
// node structure, because RootDigger allocates, relinks and frees those nodes itself
// (src/tree.cpp:213-236, 273-358).  Built with -Drooted_tree_t=... renames of host/tree.hpp's
// classes so that it can be linked next to RootDigger's own rooted_tree_t.
#include "corax/corax.h"

#include "../host/tree.hpp"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

// ---------------------------------------------------------------------------------------------
// trees
// ---------------------------------------------------------------------------------------------
namespace {

char *dup_cstr(const std::string &s) {
  char *p = (char *)malloc(s.size() + 1);
  memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

// every unode reachable from `root` over next / back
std::vector<corax_unode_t *> reachable(corax_unode_t *root) {
  std::vector<corax_unode_t *>             out, stack{root};
  std::unordered_map<corax_unode_t *, int> seen;
  while (!stack.empty()) {
    corax_unode_t *n = stack.back();
    stack.pop_back();
    if (!n || seen.count(n)) continue;
    seen[n] = 1;
    out.push_back(n);
    stack.push_back(n->next);
    stack.push_back(n->back);
  }
  return out;
}

void traverse_rec(corax_unode_t *node, int (*cb)(corax_unode_t *), corax_unode_t **out, unsigned int *n) {
  if (!cb(node)) return;  // tips are emitted, inner nodes descended, iff the callback accepts them
  if (node->next) {
    corax_unode_t *s = node->next;
    do {
      traverse_rec(s->back, cb, out, n);
      s = s->next;
    } while (s && s != node);
  }
  out[(*n)++] = node;
}

std::string newick_rec(const corax_unode_t *n, char *(*ser)(const corax_unode_t *)) {
  char       *own = ser(n);
  std::string self = own ? own : "";
  free(own);
  if (!n->next) return self;
  std::string s = "(";
  bool        first = true;
  for (const corax_unode_t *k = n->next; k != n; k = k->next) {
    if (!first) s += ",";
    s += newick_rec(k->back, ser);
    first = false;
  }
  return s + ")" + self;
}

}  // namespace

extern "C" corax_utree_t *corax_utree_parse_newick_unroot(const char *filename) {
  std::unique_ptr<rd::utree_t> src = rd::utree_parse_newick_unroot(filename);
  if (!src) {
    rdk_errno = 1; /* RDK_ERROR_PARAM */
    snprintf(rdk_errmsg, 200, "Unable to parse tree file %s", filename);
    return nullptr;
  }
  std::unordered_map<const rd::unode_t *, corax_unode_t *> m;
  for (const rd::unode_t &u : src->arena) {
    corax_unode_t *c = (corax_unode_t *)calloc(1, sizeof(corax_unode_t));
    c->label = u.has_label ? dup_cstr(u.label) : nullptr;
    c->length = u.length;
    c->node_index = u.node_index;
    c->clv_index = u.clv_index;
    c->scaler_index = u.scaler_index;
    c->pmatrix_index = u.pmatrix_index;
    m[&u] = c;
  }
  for (const rd::unode_t &u : src->arena) {
    m[&u]->next = u.next ? m.at(u.next) : nullptr;
    m[&u]->back = u.back ? m.at(u.back) : nullptr;
  }
  corax_utree_t *t = (corax_utree_t *)calloc(1, sizeof(corax_utree_t));
  t->tip_count = src->tip_count;
  t->inner_count = src->inner_count;
  t->edge_count = src->edge_count;
  t->binary = 1;
  t->nodes = (corax_unode_t **)malloc(sizeof(corax_unode_t *) * src->nodes.size());
  for (size_t i = 0; i < src->nodes.size(); ++i) t->nodes[i] = m.at(src->nodes[i]);
  t->vroot = m.at(src->vroot);
  return t;
}

extern "C" corax_utree_t *corax_utree_clone(const corax_utree_t *tree) {
  std::vector<corax_unode_t *>                        all = reachable(tree->vroot);
  std::unordered_map<corax_unode_t *, corax_unode_t *> m;
  for (corax_unode_t *u : all) {
    corax_unode_t *c = (corax_unode_t *)malloc(sizeof(corax_unode_t));
    *c = *u;  // indices, length and the data pointer as they are (coraxlib's clone is shallow there)
    c->label = u->label ? dup_cstr(u->label) : nullptr;
    m[u] = c;
  }
  for (corax_unode_t *u : all) {
    m[u]->next = u->next ? m.at(u->next) : nullptr;
    m[u]->back = u->back ? m.at(u->back) : nullptr;
  }
  corax_utree_t *t = (corax_utree_t *)calloc(1, sizeof(corax_utree_t));
  *t = *tree;
  const unsigned int count = tree->tip_count + tree->inner_count;
  t->nodes = (corax_unode_t **)malloc(sizeof(corax_unode_t *) * count);
  for (unsigned int i = 0; i < count; ++i) t->nodes[i] = m.at(tree->nodes[i]);
  t->vroot = m.at(tree->vroot);
  return t;
}

extern "C" void corax_utree_destroy(corax_utree_t *tree, void (*cb_destroy)(void *)) {
  if (!tree) return;
  for (corax_unode_t *u : reachable(tree->vroot)) {
    if (cb_destroy && u->data) cb_destroy(u->data);
    // the unodes of an inner node share one label string in coraxlib; here every unode owns a copy
    free(u->label);
    free(u);
  }
  free(tree->nodes);
  free(tree);
}

extern "C" int corax_utree_traverse(corax_unode_t *root, int traversal, int (*cbtrav)(corax_unode_t *),
                                    corax_unode_t **outbuffer, unsigned int *trav_size) {
  *trav_size = 0;
  if (traversal != CORAX_TREE_TRAVERSE_POSTORDER || !root->next) return CORAX_FAILURE;
  // the subtree behind root->back first, then root's other children, then root itself
  traverse_rec(root->back, cbtrav, outbuffer, trav_size);
  traverse_rec(root, cbtrav, outbuffer, trav_size);
  return CORAX_SUCCESS;
}

extern "C" void corax_utree_create_operations(corax_unode_t *const *trav, unsigned int count, double *branches,
                                              unsigned int *pmatrix_indices, corax_operation_t *ops,
                                              unsigned int *matrix_count, unsigned int *ops_count) {
  *matrix_count = 0;
  *ops_count = 0;
  for (unsigned int i = 0; i < count; ++i) {
    const corax_unode_t *n = trav[i];
    // the far end of the last node's edge would list that edge twice
    if (n != trav[count - 1]->back) {
      branches[*matrix_count] = n->length;
      pmatrix_indices[*matrix_count] = n->pmatrix_index;
      ++*matrix_count;
    }
    if (n->next) {
      const corax_unode_t *a = n->next->back, *b = n->next->next->back;
      corax_operation_t   &op = ops[*ops_count];
      op.parent_clv_index = n->clv_index;
      op.parent_scaler_index = n->scaler_index;
      op.child1_clv_index = a->clv_index;
      op.child1_scaler_index = a->scaler_index;
      op.child1_matrix_index = a->pmatrix_index;
      op.child2_clv_index = b->clv_index;
      op.child2_scaler_index = b->scaler_index;
      op.child2_matrix_index = b->pmatrix_index;
      ++*ops_count;
    }
  }
}

extern "C" char *corax_utree_export_newick(const corax_unode_t *root, char *(*ser)(const corax_unode_t *)) {
  if (!root->next) root = root->back;
  std::string          s = "(";
  const corax_unode_t *k = root;
  bool                 first = true;
  do {
    if (!first) s += ",";
    s += newick_rec(k->back, ser);
    first = false;
    k = k->next;
  } while (k != root);
  s += ")";
  if (root->label) s += root->label;
  s += ";";
  return dup_cstr(s);
}

namespace {
void show_rec(const corax_unode_t *n, int options, int depth) {
  printf("%*s", 2 * depth, "");
  if (options & CORAX_UTREE_SHOW_LABEL) printf("%s", n->label ? n->label : "*");
  if (options & CORAX_UTREE_SHOW_BRANCH_LENGTH) printf(" %f", n->length);
  printf("\n");
  if (n->next)
    for (const corax_unode_t *k = n->next; k != n; k = k->next) show_rec(k->back, options, depth + 1);
}
}  // namespace

extern "C" void corax_utree_show_ascii(const corax_unode_t *tree, int options) {
  if (!tree->next) tree = tree->back;
  show_rec(tree->back, options, 1);
  show_rec(tree, options, 0);
}

// ---------------------------------------------------------------------------------------------
// alignments
// ---------------------------------------------------------------------------------------------
struct corax_phylip_s {
  std::vector<std::string> lines;  // non-blank lines after the header
  size_t                   count = 0, length = 0;
};
struct corax_fasta_s {
  std::vector<std::string> labels, seqs;
  size_t                   next = 0;
};

// accepted characters: coraxlib's maps say which bytes are legal in a file; the readers here
// accept every printable character and leave validation to msa_t::valid_data (src/msa.cpp:669-687)
extern "C" const unsigned int corax_map_generic[256] = {0};
extern "C" const unsigned int corax_map_fasta[256] = {0};

namespace {

corax_msa_t *make_msa(const std::vector<std::string> &labels, const std::vector<std::string> &seqs) {
  corax_msa_t *m = (corax_msa_t *)malloc(sizeof(corax_msa_t));
  m->count = (int)seqs.size();
  m->length = seqs.empty() ? 0 : (int)seqs[0].size();
  m->sequence = (char **)malloc(sizeof(char *) * seqs.size());
  m->label = (char **)malloc(sizeof(char *) * seqs.size());
  for (size_t i = 0; i < seqs.size(); ++i) {
    m->sequence[i] = dup_cstr(seqs[i]);
    m->label[i] = dup_cstr(labels[i]);
  }
  return m;
}

}  // namespace

extern "C" corax_phylip_t *corax_phylip_open(const char *filename, const unsigned int *) {
  std::ifstream in(filename);
  if (!in) return nullptr;
  corax_phylip_t *fd = new corax_phylip_t();
  if (!(in >> fd->count >> fd->length) || fd->count == 0 || fd->length == 0) {
    delete fd;
    return nullptr;
  }
  std::string line;
  std::getline(in, line);
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.find_first_not_of(" \t") == std::string::npos) continue;
    fd->lines.push_back(line);
  }
  return fd;
}
extern "C" int  corax_phylip_rewind(corax_phylip_t *) { return CORAX_SUCCESS; }
extern "C" void corax_phylip_close(corax_phylip_t *fd) { delete fd; }

extern "C" corax_msa_t *corax_phylip_parse_interleaved(corax_phylip_t *fd) {
  const size_t n = fd->count;
  if (fd->lines.size() < n) return nullptr;
  std::vector<std::string> labels(n), seqs(n);
  for (size_t i = 0; i < fd->lines.size(); ++i) {
    std::istringstream ls(fd->lines[i]);
    std::string        tok;
    const size_t       row = i % n;
    if (i < n) {
      if (!(ls >> tok)) return nullptr;
      labels[row] = tok;
    }
    while (ls >> tok) seqs[row] += tok;
  }
  for (auto &s : seqs)
    if (s.size() != fd->length) return nullptr;
  return make_msa(labels, seqs);
}

extern "C" corax_msa_t *corax_phylip_parse_sequential(corax_phylip_t *fd) {
  const size_t             n = fd->count;
  std::vector<std::string> labels(n), seqs(n), toks;
  for (auto &l : fd->lines) {
    std::istringstream ls(l);
    std::string        t;
    while (ls >> t) toks.push_back(t);
  }
  size_t q = 0;
  for (size_t r = 0; r < n; ++r) {
    if (q >= toks.size()) return nullptr;
    labels[r] = toks[q++];
    while (seqs[r].size() < fd->length && q < toks.size()) seqs[r] += toks[q++];
    if (seqs[r].size() != fd->length) return nullptr;
  }
  if (q != toks.size()) return nullptr;
  return make_msa(labels, seqs);
}

extern "C" corax_fasta_t *corax_fasta_open(const char *filename, const unsigned int *) {
  std::ifstream in(filename);
  if (!in) return nullptr;
  corax_fasta_t *fd = new corax_fasta_t();
  std::string    line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '>') {
      std::string l = line.substr(1);
      while (!l.empty() && std::isspace((unsigned char)l.back())) l.pop_back();
      fd->labels.push_back(l);
      fd->seqs.emplace_back();
    } else {
      if (fd->labels.empty()) {
        delete fd;
        return nullptr;
      }
      for (char c : line)
        if (!std::isspace((unsigned char)c)) fd->seqs.back().push_back(c);
    }
  }
  if (fd->labels.empty()) {
    delete fd;
    return nullptr;
  }
  return fd;
}

extern "C" int corax_fasta_getnext(corax_fasta_t *fd, char **head, long *head_len, char **seq, long *seq_len,
                                   long *seqno) {
  if (fd->next >= fd->labels.size()) return CORAX_FAILURE;
  const size_t i = fd->next++;
  *head = dup_cstr(fd->labels[i]);
  *head_len = (long)fd->labels[i].size();
  *seq = dup_cstr(fd->seqs[i]);
  *seq_len = (long)fd->seqs[i].size();
  *seqno = (long)i;
  return CORAX_SUCCESS;
}
extern "C" void corax_fasta_close(corax_fasta_t *fd) { delete fd; }

extern "C" void corax_msa_destroy(corax_msa_t *msa) {
  if (!msa) return;
  for (int i = 0; i < msa->count; ++i) {
    free(msa->sequence[i]);
    if (msa->label) free(msa->label[i]);
  }
  free(msa->sequence);
  free(msa->label);
  free(msa);
}

extern "C" unsigned int *corax_compress_site_patterns(char **sequence, const corax_state_t *map, int count,
                                                      int *length) {
  const size_t n = (size_t)count, L = (size_t)*length;
  if (n == 0 || L == 0) {
    rdk_errno = 1; /* RDK_ERROR_PARAM */
    snprintf(rdk_errmsg, 200, "empty alignment");
    return nullptr;
  }
  std::vector<size_t> order(L);
  std::iota(order.begin(), order.end(), 0);
  auto code = [&](size_t r, size_t c) { return (unsigned)map[(unsigned char)sequence[r][c]]; };
  auto less = [&](size_t a, size_t b) {
    for (size_t r = 0; r < n; ++r) {
      const unsigned x = code(r, a), y = code(r, b);
      if (x != y) return x < y;
    }
    return false;
  };
  std::stable_sort(order.begin(), order.end(), less);
  std::vector<size_t>   keep;
  std::vector<unsigned> w;
  for (size_t i = 0; i < L; ++i) {
    if (!keep.empty() && !less(keep.back(), order[i]) && !less(order[i], keep.back()))
      w.back() += 1;
    else {
      keep.push_back(order[i]);
      w.push_back(1);
    }
  }
  for (size_t r = 0; r < n; ++r) {
    std::string out;
    out.reserve(keep.size());
    for (size_t c : keep) out.push_back(sequence[r][c]);
    memcpy(sequence[r], out.data(), out.size());
    sequence[r][out.size()] = '\0';
  }
  *length = (int)keep.size();
  unsigned int *weights = (unsigned int *)malloc(sizeof(unsigned int) * keep.size());
  std::copy(w.begin(), w.end(), weights);
  return weights;
}
