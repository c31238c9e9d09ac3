// NEXT ROW f4: infinite-depth Context-Tree-Weighting entropy-rate estimator.
// Reference: chaos/cppctw.cpp (ctw.estimate_entropy(seq, alphabet_size), chaos/ctw.pyx:2-3) -- the reference's only
// native code.  It is an irregular, pointer-chasing suffix-tree build: it stays on the HOST (a GPU version is not
// justified); what changes here is the data structure and the batching:
//   * nodes live in one arena (one contiguous int32 record per node: counts[A] | child[A] | tail position | tail symbol)
//     instead of a heap object with two std::vectors per node -> no per-node allocation, indices instead of pointers;
//   * a child is always created after its parent, so the code-length pass is ONE reverse sweep over the arena
//     (children before parents) -- no recursion (the reference recurses as deep as the tree: ~N for a constant sequence);
//   * lgamma(c + beta) is memoised per count (same argument -> bit-identical value);
//   * dib_ctw_estimate_entropy_batch runs independent sequences on a thread pool (nb-chaos cell 3 evaluates 75
//     independent sub-sequences per partition one after another).
// Semantics kept exactly (results are bit-identical to the reference, tests/test_ctw.py): lazily extended "tail" leaves
// (cppctw.cpp:121-129), creation depth limit 512 checked only when a new child would be made (:133-137), KT estimator
// with beta = 1/A and the weighting rule (:57-81), result root_code_length / N rounded through float (:100-104).
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dib_b200.h"

namespace {

constexpr int kMaxCreateDepth = 512;     // cppctw.cpp:13

int fail_ctw(const std::string& m);

struct ContextTree {
  int A, S;                               // alphabet size; int32 words per node
  // one record per node, contiguous: counts[A] | child[A] | tail position | tail symbol -- a visit touches one or two
  // cache lines instead of three arrays.  tail position > 0: this leaf stands for the context continuing before
  // sequence[tail position]; tail symbol: the one symbol counted so far along that continuation.
  std::vector<int32_t> pool;
  int32_t nodes = 0;

  explicit ContextTree(int alphabet) : A(alphabet), S(2 * alphabet + 2) { add_node(-1, -1); }

  int32_t* rec(int32_t node) { return pool.data() + (size_t)node * S; }
  const int32_t* rec(int32_t node) const { return pool.data() + (size_t)node * S; }

  int32_t add_node(int32_t pos, int32_t sym) {
    const size_t base = pool.size();
    pool.resize(base + S);
    int32_t* r = pool.data() + base;
    for (int i = 0; i < A; ++i) { r[i] = 0; r[A + i] = -1; }
    r[2 * A] = pos; r[2 * A + 1] = sym;
    return nodes++;
  }

  void reserve(size_t n) { pool.reserve(n * S); }

  // cppctw.cpp:106-154
  void insert_all(const int8_t* s, int64_t n) {
    const int TP = 2 * A, TS = 2 * A + 1;
    for (int64_t t = 0; t < n; ++t) {
      const int cur = s[t];
      int32_t node = 0;
      rec(0)[cur] += 1;
      for (int64_t c = t - 1; c >= 0; --c) {
        int32_t* r = rec(node);
        if (r[TP] > 0) {                                 // push the pending continuation one symbol deeper
          const int32_t p = r[TP] - 1, ts = r[TS];
          r[TP] = -1; r[TS] = -1;
          const int32_t nn = add_node(p, ts);            // may move the pool: re-derive the record pointers
          rec(node)[A + s[p]] = nn;
          rec(nn)[ts] += 1;
          r = rec(node);
        }
        const int ctx = s[c];
        const int32_t nxt = r[A + ctx];
        if (nxt < 0) {
          if (t - c > kMaxCreateDepth) break;
          const int32_t nn = c > 0 ? add_node((int32_t)c, cur) : add_node(-1, -1);
          rec(node)[A + ctx] = nn;
          rec(nn)[cur] += 1;
          break;
        }
        node = nxt;
        rec(node)[cur] += 1;
      }
    }
  }

  // cppctw.cpp:57-81 for every node, children first (child id > parent id)
  double root_code_length() {
    const double beta = 1. / A;                                    // cppctw.cpp:166
    const double lg_ab = std::lgamma(A * beta), lg_b = std::lgamma(beta), ln2 = std::log(2);
    std::vector<double> lg_cache, lg_total_cache;                  // lgamma(c + beta), lgamma(total + A beta)
    int sg = 0;
    auto memo = [&](std::vector<double>& cache, int64_t c, double offset) {
      if ((size_t)c >= cache.size()) cache.resize((size_t)c + 64, std::numeric_limits<double>::quiet_NaN());
      double& v = cache[c];
      if (std::isnan(v)) v = lgamma_r((double)c + offset, &sg);
      return v;
    };
    auto lg_count = [&](int32_t c) { return memo(lg_cache, c, beta); };
    const int32_t nn = nodes;
    std::vector<double> weighted(nn);
    for (int32_t node = nn - 1; node >= 0; --node) {
      const int32_t* cnt = rec(node);
      const int32_t* ch = cnt + A;
      int64_t itotal = 0;
      for (int i = 0; i < A; ++i) itotal += cnt[i];
      const double total = (double)itotal;
      double le = memo(lg_total_cache, itotal, A * beta) - lg_ab;
      for (int i = 0; i < A; ++i) le -= lg_count(cnt[i]) - lg_b;
      le /= ln2;
      double lc = 0.;
      bool any = false;
      for (int i = 0; i < A; ++i)
        if (ch[i] >= 0) { any = true; lc += weighted[ch[i]]; }
      weighted[node] = (any && total > 1) ? 1 + std::fmin(lc, le) - std::log2(1 + std::pow(2, -std::fabs(le - lc))) : le;
    }
    return weighted[0];
  }
};

int estimate_one(const int8_t* s, int64_t n, int A, double* out) {
  for (int64_t i = 0; i < n; ++i)
    if (s[i] < 0 || s[i] >= A) return 1;
  ContextTree tree(A);
  tree.reserve((size_t)n * 2 + 16);
  tree.insert_all(s, n);
  const float per_symbol = (float)(tree.root_code_length() / (double)(int)n);   // float return type, cppctw.cpp:100-104
  *out = per_symbol;
  return 0;
}

thread_local std::string g_ctw_error;
int fail_ctw(const std::string& m) { g_ctw_error = m; return 1; }

}  // namespace

extern "C" {

const char* dib_ctw_last_error(void) { return g_ctw_error.c_str(); }

int dib_ctw_estimate_entropy(const int8_t* sequence, int64_t length, int32_t alphabet_size, double* out_bits_per_symbol) {
  if (!out_bits_per_symbol || (!sequence && length > 0) || length < 0 || length > 0x7fffffffll || alphabet_size < 1 ||
      alphabet_size > 127)
    return fail_ctw("dib_ctw_estimate_entropy: bad arguments (0 <= length < 2^31, 1 <= alphabet_size <= 127)");
  if (estimate_one(sequence, length, alphabet_size, out_bits_per_symbol))
    return fail_ctw("dib_ctw_estimate_entropy: symbol outside [0, alphabet_size)");
  return 0;
}

int dib_ctw_estimate_entropy_batch(const int8_t* sequences, const int64_t* offsets, int32_t count, int32_t alphabet_size,
                                   int32_t num_threads, double* out_bits_per_symbol) {
  if (count < 0 || !offsets || (!sequences && count > 0 && offsets[count] > 0) || !out_bits_per_symbol ||
      alphabet_size < 1 || alphabet_size > 127)
    return fail_ctw("dib_ctw_estimate_entropy_batch: bad arguments");
  for (int32_t i = 0; i < count; ++i)
    if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 0x7fffffffll)
      return fail_ctw("dib_ctw_estimate_entropy_batch: offsets must be non-decreasing, each sequence < 2^31 symbols");
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > count) nt = count;
  std::vector<int> status((size_t)(count > 0 ? count : 1), 0);
  auto work = [&](int tid) {
    for (int32_t i = tid; i < count; i += nt)
      status[i] = estimate_one(sequences + offsets[i], offsets[i + 1] - offsets[i], alphabet_size, out_bits_per_symbol + i);
  };
  if (nt <= 1) { if (count > 0) work(0); }
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  for (int32_t i = 0; i < count; ++i)
    if (status[i]) return fail_ctw("dib_ctw_estimate_entropy_batch: sequence " + std::to_string(i) + " has a symbol outside [0, alphabet_size)");
  return 0;
}

}  // extern "C"
