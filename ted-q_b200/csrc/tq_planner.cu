// tq_planner.cu — host-only parts of the contraction planner that are worth compiling: the randomised greedy pass
// (planner.py: _greedy_once; tq_tn_greedy_path below is its bit-identical mirror) and the dynamic programme of the
// subtree reconfiguration (planner.py: reconfigure).  For a subtree with L <= 12 leaves it finds, over all subsets
// of the leaves, the cheapest order of contracting them pairwise.  planner._subtree_dp_py is the Python mirror
// (tests/test_planner_cpu.py checks bit-equality of costs and splits); the reference's counterpart is the external
// cotengra / jdtensorpath search called at compiled_circuit.py:340-393.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <set>
#include <tuple>
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

// planner.step_time_model, same operations in the same order (doubles); model = {tensor-core flop/s, bytes/s, seconds
// per step, FP32-GEMM flop/s or 0, per-element-kernel flop/s}: FIVE values
inline double pair_cost(int n_a, int n_b, int n_out, int n_union, const double* model) {
  if (!model) return std::ldexp(1.0, n_union);
  if (n_union <= 14) return 1e-7;
  double f = model[0];
  if (model[3] > 0) {
    const int k = n_union - n_out, b = n_a + n_b - n_union - k, m = n_a - k - b, n = n_b - k - b;
    const int hi = m > n ? m : n, lo = m < n ? m : n;
    if (n_union >= 20 && hi >= 7 && lo >= 4) {
    } else if (n_out <= 6 && k >= 12) {
    } else if (m >= 6 && n >= 6 && k >= 4) {
      f = model[3];
    } else if (lo <= 4 && k <= 4 && n_out >= 10 && b <= 8) {
    } else {
      f = model[4];
    }
  }
  const double t_fl = 8.0 * std::ldexp(1.0, n_union) / f;
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

// One randomised greedy pass over an all-extent-2 network (planner._greedy_once, same operations in the same order
// and the same libm calls, so the path is identical to the Python mirror's).
//   idx_off[n_inputs + 1], idx[]: index lists of the inputs, indices renumbered 0 .. n_idx-1
//   keep[n_idx]: 1 for the network's output indices
//   init_pairs[2 * n_init]: the pairs of tensors that share an index, in the order the mirror pushes them
//   u[n_u]: uniform random numbers in [0, 1), one per push when temperature > 0
// Writes the ssa path (2 * (n_inputs - 1) ids) and the number of random numbers consumed.  TQ_E_WORKSPACE: u was too
// short (call again with more).
int32_t tq_tn_greedy_path(int32_t n_inputs, int32_t n_idx, const int32_t* idx_off, const int32_t* idx, const int32_t* keep,
                          int32_t n_init, const int32_t* init_pairs, const double* u, int64_t n_u, double alpha,
                          double temperature, int32_t* path, int64_t* n_u_used) {
  TQ_REQUIRE(n_inputs >= 1 && n_idx >= 0 && idx_off && keep && path && n_u_used && (n_init == 0 || init_pairs), TQ_E_INVALID,
             "tq_tn_greedy_path: invalid argument");
  const int words = (std::max(n_idx, 1) + 63) / 64;
  const int max_ids = 2 * n_inputs;
  std::vector<uint64_t> bits((size_t)max_ids * words, 0);
  std::vector<int> size(max_ids, 0);
  std::vector<char> alive(max_ids, 0);
  std::vector<std::set<int>> where((size_t)std::max(n_idx, 1));
  auto at = [&](int t) { return bits.data() + (size_t)t * words; };
  for (int i = 0; i < n_inputs; ++i) {
    alive[i] = 1;
    for (int k = idx_off[i]; k < idx_off[i + 1]; ++k) {
      const int x = idx[k];
      TQ_REQUIRE(x >= 0 && x < n_idx, TQ_E_INVALID, "tq_tn_greedy_path: index out of range");
      if (!((at(i)[x >> 6] >> (x & 63)) & 1ull)) {
        at(i)[x >> 6] |= 1ull << (x & 63);
        ++size[i];
      }
      where[x].insert(i);
    }
  }
  // the indices of a x b that survive: output indices and indices some third tensor still holds
  std::vector<uint64_t> tmp(words);
  auto result_of = [&](int a, int b) {
    int cnt = 0;
    for (int w = 0; w < words; ++w) {
      uint64_t un = at(a)[w] | at(b)[w], out = 0;
      while (un) {
        const int bit = __builtin_ctzll(un);
        un &= un - 1;
        const int x = w * 64 + bit;
        bool stays = keep[x] != 0;
        if (!stays)
          for (int t : where[x])
            if (t != a && t != b) {
              stays = true;
              break;
            }
        if (stays) {
          out |= 1ull << bit;
          ++cnt;
        }
      }
      tmp[w] = out;
    }
    return cnt;
  };
  typedef std::tuple<double, int, int> Item;
  std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
  int64_t used = 0;
  bool starved = false;
  auto push = [&](int a, int b) {
    const int so = result_of(a, b);
    const double cost = std::ldexp(1.0, so) - alpha * (std::ldexp(1.0, size[a]) + std::ldexp(1.0, size[b]));
    double score = std::copysign(std::log2(std::fabs(cost) + 1.0), cost);
    if (temperature > 0) {
      if (used >= n_u) {
        starved = true;
        return;
      }
      const double r = u[used++];
      score -= temperature * (-std::log(-std::log(r + 1e-300) + 1e-300));
    }
    heap.push(Item(score, a, b));
  };
  for (int k = 0; k < n_init && !starved; ++k) push(init_pairs[2 * k], init_pairs[2 * k + 1]);
  int next_id = n_inputs, n_path = 0;
  while (!heap.empty() && !starved) {
    const Item top = heap.top();
    heap.pop();
    const int a = std::get<1>(top), b = std::get<2>(top);
    if (!alive[a] || !alive[b]) continue;
    const int so = result_of(a, b);
    const int c = next_id++;
    std::memcpy(at(c), tmp.data(), sizeof(uint64_t) * words);
    size[c] = so;
    for (int side = 0; side < 2; ++side) {
      const int t = side ? b : a;
      for (int w = 0; w < words; ++w) {
        uint64_t v = at(t)[w];
        while (v) {
          const int bit = __builtin_ctzll(v);
          v &= v - 1;
          where[w * 64 + bit].erase(t);
        }
      }
      alive[t] = 0;
    }
    alive[c] = 1;
    path[2 * n_path] = a;
    path[2 * n_path + 1] = b;
    ++n_path;
    std::set<int> nbrs;
    for (int w = 0; w < words; ++w) {
      uint64_t v = at(c)[w];
      while (v) {
        const int bit = __builtin_ctzll(v);
        v &= v - 1;
        std::set<int>& h = where[w * 64 + bit];
        nbrs.insert(h.begin(), h.end());
        h.insert(c);
      }
    }
    for (int t : nbrs) {
      push(t, c);
      if (starved) break;
    }
  }
  if (starved) {
    ::tq::set_error("tq_tn_greedy_path: %lld random numbers were not enough", (long long)n_u);
    return TQ_E_WORKSPACE;
  }
  // disconnected leftovers: outer products, smallest first
  std::vector<int> rest;
  for (int t = 0; t < next_id; ++t)
    if (alive[t]) rest.push_back(t);
  auto by_size = [&](int x, int y) { return size[x] != size[y] ? size[x] < size[y] : x < y; };
  std::sort(rest.begin(), rest.end(), by_size);
  while (rest.size() > 1) {
    const int a = rest[0], b = rest[1], c = next_id++;
    int cnt = 0;
    for (int w = 0; w < words; ++w) {
      at(c)[w] = at(a)[w] | at(b)[w];
      cnt += __builtin_popcountll(at(c)[w]);
    }
    size[c] = cnt;
    path[2 * n_path] = a;
    path[2 * n_path + 1] = b;
    ++n_path;
    rest.erase(rest.begin(), rest.begin() + 2);
    rest.push_back(c);
    std::sort(rest.begin(), rest.end(), by_size);
  }
  TQ_REQUIRE(n_path == n_inputs - 1, TQ_E_INVALID, "tq_tn_greedy_path: %d steps for %d inputs", n_path, n_inputs);
  *n_u_used = used;
  return TQ_OK;
}

}  // extern "C"
