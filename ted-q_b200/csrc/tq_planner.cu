// tq_planner.cu — host-only part of the contraction planner that is worth compiling: the dynamic programme of the
// subtree reconfiguration (planner.py: reconfigure).  For a subtree with L <= 12 leaves it finds, over all subsets
// of the leaves, the cheapest order of contracting them pairwise.  planner._subtree_dp_py is the Python mirror
// (tests/test_planner_cpu.py checks bit-equality of costs and splits); the reference's counterpart is the external
// cotengra / jdtensorpath search called at compiled_circuit.py:340-393.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "tq_common.h"

namespace {

struct Bits {
  int words;
  std::vector<uint64_t> w;  // [n_sets][words]
  Bits(int n_sets, int n_idx) : words((n_idx + 63) / 64), w((size_t)n_sets * ((n_idx + 63) / 64), 0) {}
  uint64_t* at(int s) { return w.data() + (size_t)s * words; }
  int count(int s) {
    int c = 0;
    for (int i = 0; i < words; ++i) c += __builtin_popcountll(at(s)[i]);
    return c;
  }
  int count_union(int a, int b) {
    int c = 0;
    for (int i = 0; i < words; ++i) c += __builtin_popcountll(at(a)[i] | at(b)[i]);
    return c;
  }
};

// planner.step_time_model, same operations in the same order (doubles)
inline double pair_cost(int n_a, int n_b, int n_out, int n_union, const double* model) {
  if (!model) return std::ldexp(1.0, n_union);
  if (n_union <= 14) return 1e-7;
  const double t_fl = 8.0 * std::ldexp(1.0, n_union) / model[0];
  const double t_by = 8.0 * (std::ldexp(1.0, n_a) + std::ldexp(1.0, n_b) + std::ldexp(1.0, n_out)) / model[1];
  return (t_fl >= t_by ? t_fl : t_by) + model[2];
}

}  // namespace

extern "C" {

int32_t tq_tn_subtree_order(int32_t n_leaves, int32_t n_idx, const int32_t* leaf_open, const int32_t* inside,
                            const int32_t* count, const double* time_model, double* best_full, int32_t* split) {
  TQ_REQUIRE(n_leaves >= 2 && n_leaves <= 12 && n_idx >= 0 && leaf_open && inside && count && best_full && split,
             TQ_E_INVALID, "tq_tn_subtree_order: invalid argument");
  const int L = n_leaves, n_sets = 1 << L, full = n_sets - 1;
  std::vector<int32_t> cnt((size_t)n_sets * (size_t)(n_idx > 0 ? n_idx : 1), 0);
  Bits open(n_sets, n_idx > 0 ? n_idx : 1);
  std::vector<int> size(n_sets, 0);
  std::vector<double> best(n_sets, 0.0);
  std::memset(split, 0, sizeof(int32_t) * (size_t)n_sets);
  for (int i = 0; i < L; ++i) {
    const int s = 1 << i;
    for (int x = 0; x < n_idx; ++x) {
      cnt[(size_t)s * n_idx + x] = inside[(size_t)i * n_idx + x];
      if (leaf_open[(size_t)i * n_idx + x]) open.at(s)[x >> 6] |= 1ull << (x & 63);
    }
    size[s] = open.count(s);
  }
  for (int S = 1; S <= full; ++S) {
    if ((S & (S - 1)) == 0) continue;
    const int low = S & -S, rest = S ^ low;
    for (int x = 0; x < n_idx; ++x) {
      const int32_t c = cnt[(size_t)rest * n_idx + x] + cnt[(size_t)low * n_idx + x];
      cnt[(size_t)S * n_idx + x] = c;
      if (c > 0 && c < count[x]) open.at(S)[x >> 6] |= 1ull << (x & 63);
    }
    size[S] = open.count(S);
    bool have = false;
    double b = 0.0;
    int bs = 0;
    for (int sub = (S - 1) & S; sub; sub = (sub - 1) & S) {
      if (!(sub & low)) continue;  // canonical split: the lowest member stays in the first part
      const int o = S ^ sub;
      const double c = best[sub] + best[o] + pair_cost(size[sub], size[o], size[S], open.count_union(sub, o), time_model);
      if (!have || c < b) {
        have = true;
        b = c;
        bs = sub;
      }
    }
    best[S] = b;
    split[S] = bs;
  }
  *best_full = best[full];
  return TQ_OK;
}

}  // extern "C"
