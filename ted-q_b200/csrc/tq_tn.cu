// tq_tn.cu — tensor-network side of tedq_b200: index maps, plan lowering, contraction executor.
//
// Replaces (a8) gen_tensor_networks' index assignment, tedq/tensor_network/tensor_network.py:850-1099,
// and (a10) the third-party tree.contract(arrays, backend='torch') call sites,
// tedq/backends/pytorch_backend.py:276,:339 (cotengra / jdtensorpath / opt_einsum: not vendored).
// All extents are 2: a tensor of rank r is addressed by r bits; "permute into GEMM" is a bit
// permutation folded into the address computation of the contraction kernel itself (no transposed
// copy of either operand is ever materialised).
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "tq_common.h"
#include "tq_tn_tc.cuh"

namespace tq {

// ---------------------------------------------------------------------------
// host: index maps
// ---------------------------------------------------------------------------
static void thread_gate(std::vector<int>& wire, int& cur, const int32_t* qubits, int k, std::vector<int>& idx) {
  idx.clear();
  for (int j = 0; j < k; ++j) idx.push_back(cur + 1 + j);
  for (int j = 0; j < k; ++j) idx.push_back(wire[qubits[j]]);
  for (int j = 0; j < k; ++j) wire[qubits[j]] = cur + 1 + j;
  cur += k;
}

// ---------------------------------------------------------------------------
// host: lowering
// ---------------------------------------------------------------------------
struct LTensor {
  std::vector<int> idx, bit;  // fast -> slow
};

struct LowerResult {
  std::vector<tq_tn_step> steps;
  std::vector<int> slice_tensor, slice_ord, slice_bit;
  std::vector<int> final_perm;
};

static int lower_impl(const int32_t* toff, const int32_t* tidx, int n_in, const int32_t* out_idx, int n_out,
                      const int32_t* path, int n_steps, const int32_t* sliced, int n_sliced, LowerResult& R) {
  std::map<int, int> sl_ord;
  for (int i = 0; i < n_sliced; ++i) sl_ord[sliced[i]] = i;
  std::map<int, int> count;
  std::map<int, LTensor> live;
  for (int t = 0; t < n_in; ++t) {
    const int r = toff[t + 1] - toff[t];
    TQ_REQUIRE(r >= 0 && r <= TQ_TN_MAX_RANK, TQ_E_UNSUPPORTED, "tq_tn_lower: input %d has rank %d > %d", t, r,
               TQ_TN_MAX_RANK);
    LTensor L;
    for (int pos = r - 1; pos >= 0; --pos) {  // fast -> slow
      const int ix = tidx[toff[t] + pos];
      const int bit = r - 1 - pos;
      auto it = sl_ord.find(ix);
      if (it != sl_ord.end()) {
        R.slice_tensor.push_back(t);
        R.slice_ord.push_back(it->second);
        R.slice_bit.push_back(bit);
      } else {
        L.idx.push_back(ix);
        L.bit.push_back(bit);
        count[ix] += 1;
      }
    }
    live[t] = L;
  }
  // the Python mirror lists slice entries per tensor in slow -> fast order
  {
    std::vector<int> order(R.slice_tensor.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (R.slice_tensor[a] != R.slice_tensor[b]) return R.slice_tensor[a] < R.slice_tensor[b];
      return R.slice_bit[a] > R.slice_bit[b];
    });
    std::vector<int> a, b, c;
    for (int i : order) {
      a.push_back(R.slice_tensor[i]);
      b.push_back(R.slice_ord[i]);
      c.push_back(R.slice_bit[i]);
    }
    R.slice_tensor = a;
    R.slice_ord = b;
    R.slice_bit = c;
  }
  for (int i = 0; i < n_out; ++i) {
    TQ_REQUIRE(!sl_ord.count(out_idx[i]), TQ_E_INVALID, "tq_tn_lower: an open output index cannot be sliced");
    count[out_idx[i]] += 1;
  }
  int nxt = n_in;
  for (int s = 0; s < n_steps; ++s) {
    const int a = path[2 * s], b = path[2 * s + 1];
    TQ_REQUIRE(live.count(a) && live.count(b) && a != b, TQ_E_INVALID, "tq_tn_lower: step %d uses a dead tensor", s);
    LTensor A = live[a], B = live[b];
    live.erase(a);
    live.erase(b);
    std::map<int, int> pos_b;
    for (size_t i = 0; i < B.idx.size(); ++i) pos_b[B.idx[i]] = B.bit[i];
    std::set<int> in_a(A.idx.begin(), A.idx.end());
    tq_tn_step st;
    memset(&st, 0, sizeof(st));
    st.lhs = a;
    st.rhs = b;
    std::vector<int> K, M, Bt, N;  // positions into A (K, M, Bt) / B (N)
    for (size_t i = 0; i < A.idx.size(); ++i) {
      const int ix = A.idx[i];
      if (pos_b.count(ix)) {
        (count[ix] == 2 ? K : Bt).push_back((int)i);
      } else {
        M.push_back((int)i);
      }
    }
    for (size_t i = 0; i < B.idx.size(); ++i)
      if (!in_a.count(B.idx[i])) N.push_back((int)i);
    for (int i : M) TQ_REQUIRE(count[A.idx[i]] >= 2, TQ_E_INVALID, "tq_tn_lower: dangling index %d", A.idx[i]);
    for (int i : N) TQ_REQUIRE(count[B.idx[i]] >= 2, TQ_E_INVALID, "tq_tn_lower: dangling index %d", B.idx[i]);
    st.n_k = (int)K.size();
    st.n_m = (int)M.size();
    st.n_n = (int)N.size();
    st.n_b = (int)Bt.size();
    TQ_REQUIRE(st.n_k + st.n_m + st.n_b <= TQ_TN_MAX_RANK && st.n_k + st.n_n + st.n_b <= TQ_TN_MAX_RANK &&
                   st.n_m + st.n_n + st.n_b <= TQ_TN_MAX_RANK,
               TQ_E_UNSUPPORTED, "tq_tn_lower: step %d exceeds rank %d", s, TQ_TN_MAX_RANK);
    int w = 0;
    for (int i : K) st.lhs_bits[w++] = (int8_t)A.bit[i];
    for (int i : M) st.lhs_bits[w++] = (int8_t)A.bit[i];
    for (int i : Bt) st.lhs_bits[w++] = (int8_t)A.bit[i];
    w = 0;
    for (int i : K) st.rhs_bits[w++] = (int8_t)pos_b[A.idx[i]];
    for (int i : N) st.rhs_bits[w++] = (int8_t)B.bit[i];
    for (int i : Bt) st.rhs_bits[w++] = (int8_t)pos_b[A.idx[i]];
    LTensor C;
    for (int i : N) C.idx.push_back(B.idx[i]);
    for (int i : M) C.idx.push_back(A.idx[i]);
    for (int i : Bt) C.idx.push_back(A.idx[i]);
    for (size_t j = 0; j < C.idx.size(); ++j) {
      C.bit.push_back((int)j);
      st.out_idx[j] = C.idx[j];
    }
    for (int i : K) count[A.idx[i]] = 0;
    for (int i : Bt) count[A.idx[i]] -= 1;
    R.steps.push_back(st);
    live[nxt++] = C;
  }
  TQ_REQUIRE(live.size() == 1, TQ_E_INVALID, "tq_tn_lower: path leaves %zu tensors", live.size());
  const LTensor& last = live.begin()->second;
  TQ_REQUIRE((int)last.idx.size() == n_out, TQ_E_INVALID, "tq_tn_lower: final rank %zu != %d open indices",
             last.idx.size(), n_out);
  R.final_perm.assign(n_out, -1);
  for (int j = 0; j < n_out; ++j) {
    const int want = out_idx[n_out - 1 - j];
    for (size_t i = 0; i < last.idx.size(); ++i)
      if (last.idx[i] == want) R.final_perm[j] = last.bit[i];
    TQ_REQUIRE(R.final_perm[j] >= 0, TQ_E_INVALID, "tq_tn_lower: open index %d missing from the result", want);
  }
  return TQ_OK;
}

// ---------------------------------------------------------------------------
// device: contraction kernels
// ---------------------------------------------------------------------------
struct StepDev {
  int32_t n_k, n_m, n_n, n_b;
  int8_t a_m[TQ_TN_MAX_RANK], a_b[TQ_TN_MAX_RANK];  // scatter of m / kept-shared bits into A
  int8_t b_n[TQ_TN_MAX_RANK], b_b[TQ_TN_MAX_RANK];  // scatter of n / kept-shared bits into B
  const int32_t* ka;                                // K offset tables, 2^n_k entries each (2^10 for reductions)
  const int32_t* kb;
  int32_t a_k_fast, b_k_fast;  // the fastest physical bit of A / B is a K bit
  int8_t a_k[TQ_TN_MAX_RANK], b_k[TQ_TN_MAX_RANK];  // scatter of k bits into A / B
};

__device__ __forceinline__ uint32_t scat(uint32_t v, const int8_t* pos, int n) {
  uint32_t r = 0;
  for (int j = 0; j < n; ++j) r |= ((v >> j) & 1u) << pos[j];
  return r;
}

// One thread per output element; K loop through offset tables.  Used for every small step
// (the bandwidth/latency-bound early part of a network) and as the general fallback.
template <typename R>
__global__ void __launch_bounds__(256)
k_tn_step(const cx<R>* __restrict__ A, int64_t sA, const cx<R>* __restrict__ B, int64_t sB, cx<R>* __restrict__ C,
          int64_t sC, const __grid_constant__ StepDev d, int64_t n_out_elems) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out_elems) return;
  const int64_t set = blockIdx.y;
  const uint32_t n = (uint32_t)(o & ((1ll << d.n_n) - 1));
  const uint32_t m = (uint32_t)((o >> d.n_n) & ((1ll << d.n_m) - 1));
  const uint32_t bb = (uint32_t)(o >> (d.n_n + d.n_m));
  const cx<R>* a = A + set * sA + (scat(m, d.a_m, d.n_m) | scat(bb, d.a_b, d.n_b));
  const cx<R>* b = B + set * sB + (scat(n, d.b_n, d.n_n) | scat(bb, d.b_b, d.n_b));
  cx<R> acc = mk<R>(0, 0);
  const int K = 1 << d.n_k;
  for (int k = 0; k < K; ++k) acc = cfma(a[d.ka[k]], b[d.kb[k]], acc);
  C[set * sC + o] = acc;
}

// "Apply" kernel: one operand is gate-sized (<= 4 free indices, <= 4 contracted), the other is large — the shape
// of every step of a state-vector-like plan and of the mid-size steps between the fused runs and the GEMMs.  One
// thread per (large free index value f, kept-shared value bb): it loads the 2^k entries of the large operand it
// needs ONCE, keeps them in registers, and produces all 2^s outputs; the small operand sits in shared memory in
// [bb][s][k] order.  Traffic: the large operand read once, the result written once.
constexpr int APPLY_MAX = 4;  // log2 of the largest small-operand extents handled (k and s each)
struct ApplyDev {
  int32_t n_k, n_s, n_f, n_b;
  int32_t s_shift, f_shift;  // result element = bb << (n_s + n_f) | s << s_shift | f << f_shift
  int8_t small_k[APPLY_MAX], small_s[APPLY_MAX], small_b[8];
  int8_t big_k[APPLY_MAX], big_f[TQ_TN_MAX_RANK], big_b[8];
};
template <typename R, int KMAX, int SMAX>
__global__ void __launch_bounds__(256)
k_tn_apply(const cx<R>* __restrict__ S, int64_t sS, const cx<R>* __restrict__ T, int64_t sT, cx<R>* __restrict__ C,
           int64_t sC, const __grid_constant__ ApplyDev d, int64_t n_threads) {
  __shared__ cx<R> small[1 << (2 * APPLY_MAX + 2)];  // [bb' ][s][k], bb' = low 2 kept-shared bits at most staged
  const int64_t set = blockIdx.y;
  const int K = 1 << d.n_k, Sn = 1 << d.n_s;
  const int nb_st = min(d.n_b, 2);  // kept-shared bits resolved through the staged copy
  const cx<R>* sp = S + set * sS;
  for (int i = threadIdx.x; i < (1 << (d.n_k + d.n_s + nb_st)); i += 256) {
    const uint32_t k = i & (K - 1), sv = (i >> d.n_k) & (Sn - 1), bl = i >> (d.n_k + d.n_s);
    small[i] = sp[scat(k, d.small_k, d.n_k) | scat(sv, d.small_s, d.n_s) | scat(bl, d.small_b, nb_st)];
  }
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= n_threads) return;
  const uint32_t f = (uint32_t)(t & (((int64_t)1 << d.n_f) - 1)), bb = (uint32_t)(t >> d.n_f);
  const cx<R>* tp = T + set * sT + (scat(f, d.big_f, d.n_f) | scat(bb, d.big_b, d.n_b));
  cx<R> v[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) v[k] = tp[scat((uint32_t)k, d.big_k, d.n_k)];
  // kept-shared bits beyond the staged two select a different small tensor: read it from global memory
  const uint32_t b_lo = bb & ((1u << nb_st) - 1u), b_hi = bb >> nb_st;
  const cx<R>* sm = small + ((int64_t)b_lo << (d.n_k + d.n_s));
  const cx<R>* sg = sp + scat(b_hi, d.small_b + nb_st, d.n_b - nb_st);
  cx<R>* cp = C + set * sC + ((int64_t)bb << (d.n_s + d.n_f)) + ((int64_t)f << d.f_shift);
#pragma unroll
  for (int sv = 0; sv < SMAX; ++sv) {
    if (sv < Sn) {
      cx<R> acc = mk<R>(0, 0);
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
          const cx<R> a = d.n_b > nb_st
                              ? sg[scat((uint32_t)k, d.small_k, d.n_k) | scat((uint32_t)sv, d.small_s, d.n_s) |
                                   scat(b_lo, d.small_b, nb_st)]
                              : sm[(sv << d.n_k) | k];
          acc = cfma(a, v[k], acc);
        }
      cp[(int64_t)sv << d.s_shift] = acc;
    }
  }
}

// Split-K reduction for steps with a handful of output elements and a long contracted extent (the closing
// steps of an amplitude network: a 2^21-term dot product).  Block (x, o, set) reduces K range x of output
// element o into partial[set][o][x]; k_tn_dot_sum adds the partials in a fixed order (deterministic).
constexpr int DOT_LO = 10;  // k bits resolved through the offset tables; the rest by bit scatter per 1024-chunk
template <typename R>
__global__ void __launch_bounds__(256)
k_tn_dot(const cx<R>* __restrict__ A, int64_t sA, const cx<R>* __restrict__ B, int64_t sB,
         cx<R>* __restrict__ partial, const __grid_constant__ StepDev d) {
  const uint32_t o = blockIdx.y, set = blockIdx.z;
  const uint32_t n = o & ((1u << d.n_n) - 1u);
  const uint32_t m = (o >> d.n_n) & ((1u << d.n_m) - 1u);
  const uint32_t bb = o >> (d.n_n + d.n_m);
  const cx<R>* a = A + (int64_t)set * sA + (scat(m, d.a_m, d.n_m) | scat(bb, d.a_b, d.n_b));
  const cx<R>* b = B + (int64_t)set * sB + (scat(n, d.b_n, d.n_n) | scat(bb, d.b_b, d.n_b));
  const int chunks = 1 << (d.n_k - DOT_LO);
  cx<R> acc = mk<R>(0, 0);
  for (int ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    const uint32_t ah = scat((uint32_t)ch, d.a_k + DOT_LO, d.n_k - DOT_LO);
    const uint32_t bh = scat((uint32_t)ch, d.b_k + DOT_LO, d.n_k - DOT_LO);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kl = threadIdx.x + 256 * i;
      acc = cfma(a[ah | (uint32_t)d.ka[kl]], b[bh | (uint32_t)d.kb[kl]], acc);
    }
  }
  acc.x = warp_sum(acc.x);
  acc.y = warp_sum(acc.y);
  __shared__ cx<R> part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    cx<R> t = part[0];
    for (int w = 1; w < 8; ++w) {
      t.x += part[w].x;
      t.y += part[w].y;
    }
    partial[((int64_t)set * gridDim.y + o) * gridDim.x + blockIdx.x] = t;
  }
}
template <typename R>
__global__ void k_tn_dot_sum(const cx<R>* __restrict__ partial, int n_part, cx<R>* __restrict__ C, int64_t sC,
                             int n_out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int64_t set = blockIdx.y;
  const cx<R>* src = partial + (set * n_out + o) * n_part;
  cx<R> t = mk<R>(0, 0);
  for (int i = 0; i < n_part; ++i) {
    t.x += src[i].x;
    t.y += src[i].y;
  }
  C[set * sC + o] = t;
}

// ---------------------------------------------------------------------------
// Fused executor for the small steps of a plan (the bandwidth/latency-bound early part of every network:
// 1700 of the 1743 steps of the 12-qubit MBL network have <= 16 output elements).  One launch runs a whole
// dependency-closed set of small steps: one CTA per parameter set walks the run level by level (steps of a
// level are independent); within a level the larger steps are done by the whole CTA one after the other and
// the tiny ones are dealt out one per warp; one __syncthreads per level.  Output elements with few MACs are
// split over up to 32 adjacent lanes along K and combined with warp shuffles.  Intermediates stay in the
// (L1/L2-resident) arenas, so the larger kernels read them unchanged.
// ---------------------------------------------------------------------------
constexpr int FUSE_MAX_RANK = 16, FUSE_MAX_SLICE_BITS = 4;
struct alignas(16) FusedStep {
  int64_t a_off, b_off, c_off;  // complex-entry offsets inside their space
  int32_t a_in, b_in;           // >= 0: input tensor id, -1: shared arena, -2: per-set arena
  int32_t c_space;              // -1 / -2
  int8_t n_k, n_m, n_n, n_b;
  int8_t is_cta;
  int8_t micro_iters;           // > 0: a "micro" step — <= 32 (output, K-part) items, <= 8 K iterations per lane;
  int8_t micro_lpo, pad;        //      every lane's operand offsets are precomputed on the host (micro_row)
  int32_t micro_row;            //      first row of its [micro_iters][32] table of (a_off | b_off << 16)
  int8_t a_bits[FUSE_MAX_RANK];  // bit positions inside A of [k..., m..., b...]
  int8_t b_bits[FUSE_MAX_RANK];  // bit positions inside B of [k..., n..., b...]
  int8_t a_sl_ord[FUSE_MAX_SLICE_BITS], a_sl_bit[FUSE_MAX_SLICE_BITS];  // sliced bits of an input operand (-1: none)
  int8_t b_sl_ord[FUSE_MAX_SLICE_BITS], b_sl_bit[FUSE_MAX_SLICE_BITS];
};
struct InputRef {
  const void* ptr;
  int64_t stride;  // complex entries between parameter sets (0: shared)
};

template <typename R>
__device__ __forceinline__ const cx<R>* fused_operand(const InputRef* __restrict__ inputs, cx<R>* shared, cx<R>* perset,
                                                      int64_t set, int64_t slice, int32_t in, int64_t off,
                                                      const int8_t* ord, const int8_t* bit) {
  if (in == -1) return shared + off;
  if (in == -2) return perset + off;
  int64_t so = 0;
  if (*reinterpret_cast<const int32_t*>(ord) != -1) {  // any sliced bit at all? (four int8 entries, -1 = none)
#pragma unroll
    for (int j = 0; j < FUSE_MAX_SLICE_BITS; ++j)
      if (ord[j] >= 0) so |= (int64_t)((slice >> ord[j]) & 1) << bit[j];
  }
  return reinterpret_cast<const cx<R>*>(inputs[in].ptr) + set * inputs[in].stride + so + off;
}

// One small step by a group of G threads (a warp, or the whole CTA when KTAB is set).  KTAB: the K offsets of
// both operands come from shared-memory tables built once per step (low FUSE_KTAB_LOG2 bits; higher bits by
// scatter per outer iteration) — the inner loop is 2 LDS + 2 LDG + 4 FMA instead of two bit-scatter loops.
constexpr int FUSE_KTAB_LOG2 = 10;
constexpr int FUSE_TABLE_WORK = 12;  // CTA steps with at least 2^12 MACs build offset tables in shared memory
template <typename R, bool KTAB>
__device__ __forceinline__ void fused_exec(const FusedStep& f, const InputRef* __restrict__ inputs, cx<R>* shared,
                                           cx<R>* perset, int64_t set, int64_t slice, int t, int G, int log2G,
                                           const uint32_t* ktab_a, const uint32_t* ktab_b, const uint32_t* mtab,
                                           const uint32_t* ntab) {
  const cx<R>* a = fused_operand<R>(inputs, shared, perset, set, slice, f.a_in, f.a_off, f.a_sl_ord, f.a_sl_bit);
  const cx<R>* b = fused_operand<R>(inputs, shared, perset, set, slice, f.b_in, f.b_off, f.b_sl_ord, f.b_sl_bit);
  cx<R>* c = (f.c_space == -1 ? shared : perset) + f.c_off;
  const int n_k = f.n_k, n_m = f.n_m, n_n = f.n_n, n_b = f.n_b;
  // lanes per output element: split K over adjacent lanes when the step has fewer outputs than the group
  const int lpo_log2 = max(0, min(min(n_k, 5), log2G - (n_m + n_n + n_b)));
  const int lpo = 1 << lpo_log2;
  const int items = 1 << (n_m + n_n + n_b + lpo_log2);
  const int K = 1 << n_k;
  const int n_lo = min(n_k, FUSE_KTAB_LOG2), K_lo = 1 << n_lo;
  for (int base = 0; base < items; base += G) {
    const int it = base + t;
    const bool valid = it < items;
    const uint32_t o = (uint32_t)(it >> lpo_log2), kp = (uint32_t)(it & (lpo - 1));
    const uint32_t n = o & ((1u << n_n) - 1u);
    const uint32_t m = (o >> n_n) & ((1u << n_m) - 1u);
    const uint32_t bb = o >> (n_n + n_m);
    cx<R> acc = mk<R>(0, 0);
    if (valid) {
      uint32_t ao, bo;
      if (KTAB) {  // m / n offsets from the per-step tables (low FUSE_KTAB_LOG2 bits), the rest by scatter
        const int m_lo = min(n_m, FUSE_KTAB_LOG2), n_lo2 = min(n_n, FUSE_KTAB_LOG2);
        ao = mtab[m & ((1u << m_lo) - 1u)] | scat(m >> m_lo, f.a_bits + n_k + m_lo, n_m - m_lo) |
             scat(bb, f.a_bits + n_k + n_m, n_b);
        bo = ntab[n & ((1u << n_lo2) - 1u)] | scat(n >> n_lo2, f.b_bits + n_k + n_lo2, n_n - n_lo2) |
             scat(bb, f.b_bits + n_k + n_n, n_b);
      } else {
        ao = scat(m, f.a_bits + n_k, n_m) | scat(bb, f.a_bits + n_k + n_m, n_b);
        bo = scat(n, f.b_bits + n_k, n_n) | scat(bb, f.b_bits + n_k + n_n, n_b);
      }
      if (KTAB) {
        for (int kh = 0; kh < (K >> n_lo); ++kh) {
          const uint32_t ah = ao | scat((uint32_t)kh, f.a_bits + n_lo, n_k - n_lo);
          const uint32_t bh = bo | scat((uint32_t)kh, f.b_bits + n_lo, n_k - n_lo);
          int k = (int)kp;
          for (; k + 3 * lpo < K_lo; k += 4 * lpo) {  // four independent operand pairs in flight
            const cx<R> a0 = a[ah | ktab_a[k]], b0 = b[bh | ktab_b[k]];
            const cx<R> a1 = a[ah | ktab_a[k + lpo]], b1 = b[bh | ktab_b[k + lpo]];
            const cx<R> a2 = a[ah | ktab_a[k + 2 * lpo]], b2 = b[bh | ktab_b[k + 2 * lpo]];
            const cx<R> a3 = a[ah | ktab_a[k + 3 * lpo]], b3 = b[bh | ktab_b[k + 3 * lpo]];
            acc = cfma(a0, b0, acc);
            acc = cfma(a1, b1, acc);
            acc = cfma(a2, b2, acc);
            acc = cfma(a3, b3, acc);
          }
          for (; k < K_lo; k += lpo) acc = cfma(a[ah | ktab_a[k]], b[bh | ktab_b[k]], acc);
        }
      } else {
        int k = (int)kp;
        for (; k + lpo < K; k += 2 * lpo) {  // two independent operand pairs in flight
          const cx<R> a0 = a[ao | scat((uint32_t)k, f.a_bits, n_k)], b0 = b[bo | scat((uint32_t)k, f.b_bits, n_k)];
          const cx<R> a1 = a[ao | scat((uint32_t)(k + lpo), f.a_bits, n_k)];
          const cx<R> b1 = b[bo | scat((uint32_t)(k + lpo), f.b_bits, n_k)];
          acc = cfma(a0, b0, acc);
          acc = cfma(a1, b1, acc);
        }
        for (; k < K; k += lpo)
          acc = cfma(a[ao | scat((uint32_t)k, f.a_bits, n_k)], b[bo | scat((uint32_t)k, f.b_bits, n_k)], acc);
      }
    }
    for (int d = lpo >> 1; d > 0; d >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, d);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, d);
    }
    if (valid && kp == 0) c[o] = acc;
  }
}

// Micro step: the 1700-of-1743 kind (a handful of outputs, K <= 32).  Nothing is decoded on the device: lane l
// reads its operand offsets for K iteration j from table[j][l] (one coalesced 128-byte load per iteration).
template <typename R>
__device__ __forceinline__ void fused_micro(const FusedStep& f, const uint32_t* __restrict__ table,
                                            const InputRef* __restrict__ inputs, cx<R>* shared, cx<R>* perset,
                                            int64_t set, int64_t slice, int lane) {
  const cx<R>* a = fused_operand<R>(inputs, shared, perset, set, slice, f.a_in, f.a_off, f.a_sl_ord, f.a_sl_bit);
  const cx<R>* b = fused_operand<R>(inputs, shared, perset, set, slice, f.b_in, f.b_off, f.b_sl_ord, f.b_sl_bit);
  const uint32_t* row = table + (int64_t)f.micro_row * 32 + lane;
  uint32_t off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) off[j] = j < f.micro_iters ? row[j * 32] : 0xffffffffu;
  cx<R> acc = mk<R>(0, 0);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (off[j] != 0xffffffffu) acc = cfma(a[off[j] & 0xffffu], b[off[j] >> 16], acc);
  const int lpo = 1 << f.micro_lpo;
  for (int d = lpo >> 1; d > 0; d >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, d);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, d);
  }
  if (off[0] != 0xffffffffu && (lane & (lpo - 1)) == 0) {
    cx<R>* c = (f.c_space == -1 ? shared : perset) + f.c_off;
    c[lane >> f.micro_lpo] = acc;
  }
}

constexpr int FUSE_STAGE = 128;  // step descriptors staged in shared memory per chunk
constexpr int FUSE_CLUSTER = 8;  // CTAs that share a run common to every parameter set
// CLUSTER > 1 (runs shared by every parameter set: ONE set of steps for the whole GPU): a thread-block cluster
// of CLUSTER CTAs works on the run together — CTA steps are dealt round-robin to the CTAs, warp steps to all
// CLUSTER x NW warps — and the level boundary is a hardware cluster barrier (release / acquire) instead of
// __syncthreads.  Intermediates are written once and read in a later level, in 128-byte-aligned blocks.
template <typename R, int THREADS, int CLUSTER>
__global__ void __launch_bounds__(THREADS)
k_tn_fused(const FusedStep* __restrict__ steps, const int32_t* __restrict__ level_off,
           const int32_t* __restrict__ level_ncta, int n_levels, const uint32_t* __restrict__ micro,
           const InputRef* __restrict__ inputs, cx<R>* shared, cx<R>* perset_base, int64_t set_stride,
           int64_t slice) {
  static_assert(sizeof(FusedStep) % 16 == 0, "descriptors are staged with 16-byte copies");
  __shared__ __align__(16) FusedStep sdesc[FUSE_STAGE];
  __shared__ uint32_t ktab_a[1 << FUSE_KTAB_LOG2], ktab_b[1 << FUSE_KTAB_LOG2];
  __shared__ uint32_t mtab[1 << FUSE_KTAB_LOG2], ntab[1 << FUSE_KTAB_LOG2];
  uint32_t crank = 0;
  if (CLUSTER > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int64_t set = blockIdx.x / CLUSTER;
  cx<R>* perset = perset_base + set * set_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = THREADS / 32, V = sizeof(FusedStep) / 16;
  constexpr int LOG2T = THREADS == 1024 ? 10 : THREADS == 512 ? 9 : 8;
  for (int L = 0; L < n_levels; ++L) {
    const int beg = level_off[L], end = level_off[L + 1], cta_end = beg + level_ncta[L];
    for (int c0 = beg; c0 < end; c0 += FUSE_STAGE) {
      const int c1 = min(end, c0 + FUSE_STAGE);
      const uint4* src = reinterpret_cast<const uint4*>(steps + c0);
      for (int i = threadIdx.x; i < (c1 - c0) * V; i += THREADS) reinterpret_cast<uint4*>(sdesc)[i] = src[i];
      __syncthreads();
      const int ncta = max(0, min(cta_end, c1) - c0);  // cta steps come first inside a level
      for (int i = (int)crank; i < ncta; i += CLUSTER) {
        const FusedStep& f = sdesc[i];
        if (f.n_k + f.n_m + f.n_n + f.n_b < FUSE_TABLE_WORK) {
          // medium step: the whole CTA, offsets by bit scatter — no tables, hence no barriers: independent steps
          // of a level stream through back to back
          fused_exec<R, false>(f, inputs, shared, perset, set, slice, (int)threadIdx.x, THREADS, LOG2T, nullptr,
                               nullptr, nullptr, nullptr);
          continue;
        }
        const int n_lo = min((int)f.n_k, FUSE_KTAB_LOG2);
        for (int k = threadIdx.x; k < (1 << n_lo); k += THREADS) {
          ktab_a[k] = scat((uint32_t)k, f.a_bits, n_lo);
          ktab_b[k] = scat((uint32_t)k, f.b_bits, n_lo);
        }
        const int m_lo = min((int)f.n_m, FUSE_KTAB_LOG2), n_lo2 = min((int)f.n_n, FUSE_KTAB_LOG2);
        for (int v = threadIdx.x; v < (1 << m_lo); v += THREADS) mtab[v] = scat((uint32_t)v, f.a_bits + f.n_k, m_lo);
        for (int v = threadIdx.x; v < (1 << n_lo2); v += THREADS) ntab[v] = scat((uint32_t)v, f.b_bits + f.n_k, n_lo2);
        __syncthreads();
        fused_exec<R, true>(f, inputs, shared, perset, set, slice, (int)threadIdx.x, THREADS, LOG2T, ktab_a, ktab_b,
                            mtab, ntab);
        __syncthreads();
      }
      for (int i = ncta + (int)crank * NW + warp; i < c1 - c0; i += NW * CLUSTER) {
        if (sdesc[i].micro_iters > 0)
          fused_micro<R>(sdesc[i], micro, inputs, shared, perset, set, slice, lane);
        else
          fused_exec<R, false>(sdesc[i], inputs, shared, perset, set, slice, lane, 32, 5, nullptr, nullptr, nullptr,
                               nullptr);
      }
      __syncthreads();  // descriptor buffer reuse (and the level boundary when one CTA owns the run)
    }
    if (CLUSTER > 1) {  // level boundary across the cluster
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  }
}

// ---------------------------------------------------------------------------
// Apply-chain sweep: a RUN of consecutive gate-like apply steps on the same large tensor (every step contracts k <= 4
// indices of it with a small operand and puts k new ones in their place — the shape of every step of a
// state-vector-like plan) in ONE launch.  A CTA owns a tile of 2^L elements of one parameter set: the L "local"
// indices are those the run's steps touch (plus the lowest bits of the layout, for coalescing); the other indices
// of the tensor pass through untouched and only number the tiles.  The tile is read once from the first step's
// operand, every step is applied in shared memory (a new index takes the bit position of the one it replaces: no
// data movement), and the tile is written once in the LAST step's result layout.  The tensors between the steps of
// the run are never materialised: 2 passes over the tensor per run instead of 2 per step.
// ---------------------------------------------------------------------------
constexpr int CHAIN_MAX_K = 4, CHAIN_MAX_B = 2, CHAIN_MAX_LOCAL = 13, CHAIN_GATE_ENTRIES = 4096;
struct alignas(16) ChainStep {
  int64_t off;                 // element offset of the small operand inside its space
  int32_t in;                  // >= 0: input tensor id, -1: shared arena, -2: per-set arena
  int32_t g_begin;             // first entry of its staged gate G[bv][s][k] in shared memory
  int8_t n_k, n_b, pad0, pad1;
  int8_t tpos[CHAIN_MAX_K];    // tile bit of contracted index j = tile bit of the new index j (ascending in j? no: any)
  int8_t tsort[CHAIN_MAX_K];   // tpos sorted ascending (zero-bit insertion order)
  int8_t bpos[CHAIN_MAX_B];    // kept-shared index j: tile bit (>= 0) or -1 - (tile-number bit) when it is not local
  int8_t gk[CHAIN_MAX_K], gs[CHAIN_MAX_K], gb[CHAIN_MAX_B];  // bit of k_j / new_j / b_j inside the small operand
  int8_t sl_ord[4], sl_bit[4]; // sliced bits of an input operand (-1: none)
};
struct ChainHdr {
  int32_t L, n_glob, n_steps, gate_entries;
  int8_t in_loc[CHAIN_MAX_LOCAL], out_loc[CHAIN_MAX_LOCAL];  // physical bit of tile bit j in the first operand / last result
  int8_t in_glob[TQ_TN_MAX_RANK], out_glob[TQ_TN_MAX_RANK];  // physical bit of tile-number bit j
};

template <typename R, int K>
__device__ __forceinline__ void chain_apply(cx<R>* tile, const cx<R>* G, const ChainStep& c, int L, uint32_t tile_id) {
  constexpr int D = 1 << K;
  const uint32_t groups = 1u << (L - K);
  uint32_t toff[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    uint32_t o = 0;
#pragma unroll
    for (int t = 0; t < K; ++t) o |= ((j >> t) & 1u) << c.tpos[t];
    toff[j] = o;
  }
  for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) {
    uint32_t base = g;
#pragma unroll
    for (int t = 0; t < K; ++t) base = insert_zero_bit(base, c.tsort[t]);
    uint32_t bv = 0;
    for (int j = 0; j < c.n_b; ++j) {
      const int bp = c.bpos[j];
      bv |= (bp >= 0 ? (base >> bp) & 1u : (tile_id >> (-1 - bp)) & 1u) << j;
    }
    const cx<R>* Gb = G + ((size_t)bv << (2 * K));
    cx<R> v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = tile[base | toff[j]];
#pragma unroll
    for (int sv = 0; sv < D; ++sv) {
      cx<R> acc = cmul(Gb[sv * D], v[0]);
#pragma unroll
      for (int j = 1; j < D; ++j) acc = cfma(Gb[sv * D + j], v[j], acc);
      tile[base | toff[sv]] = acc;
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(256)
k_tn_chain(const __grid_constant__ ChainHdr h, const ChainStep* __restrict__ steps, const InputRef* __restrict__ inputs,
           cx<R>* shared, cx<R>* perset_base, int64_t set_stride, int64_t slice, const cx<R>* __restrict__ src,
           int64_t src_stride, cx<R>* __restrict__ dst, int64_t dst_stride) {
  extern __shared__ __align__(16) unsigned char chain_smem[];
  cx<R>* tile = reinterpret_cast<cx<R>*>(chain_smem);
  cx<R>* gates = tile + ((size_t)1 << h.L);
  const int64_t set = blockIdx.y;
  const uint32_t tile_id = blockIdx.x;
  cx<R>* perset = perset_base + set * set_stride;
  const uint32_t n_tile = 1u << h.L;
  uint32_t in_base = 0, out_base = 0;
  for (int j = 0; j < h.n_glob; ++j) {
    in_base |= ((tile_id >> j) & 1u) << h.in_glob[j];
    out_base |= ((tile_id >> j) & 1u) << h.out_glob[j];
  }
  const cx<R>* sp = src + set * src_stride + in_base;
  for (uint32_t l = threadIdx.x; l < n_tile; l += blockDim.x) tile[l] = sp[scat(l, h.in_loc, h.L)];
  // stage every gate of the run: G[bv][new][old], entry e of step i at gates[g_begin + e]
  for (int i = 0; i < h.n_steps; ++i) {
    const ChainStep& c = steps[i];
    const cx<R>* g = fused_operand<R>(inputs, shared, perset, set, slice, c.in, c.off, c.sl_ord, c.sl_bit);
    const int n = 1 << (2 * c.n_k + c.n_b);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const uint32_t kk = e & ((1u << c.n_k) - 1u), sv = (e >> c.n_k) & ((1u << c.n_k) - 1u), bv = e >> (2 * c.n_k);
      gates[c.g_begin + e] = g[scat(kk, c.gk, c.n_k) | scat(sv, c.gs, c.n_k) | scat(bv, c.gb, c.n_b)];
    }
  }
  __syncthreads();
  for (int i = 0; i < h.n_steps; ++i) {
    const ChainStep c = steps[i];
    const cx<R>* G = gates + c.g_begin;
    switch (c.n_k) {
      case 1: chain_apply<R, 1>(tile, G, c, h.L, tile_id); break;
      case 2: chain_apply<R, 2>(tile, G, c, h.L, tile_id); break;
      case 3: chain_apply<R, 3>(tile, G, c, h.L, tile_id); break;
      default: chain_apply<R, 4>(tile, G, c, h.L, tile_id); break;
    }
    __syncthreads();
  }
  cx<R>* dp = dst + set * dst_stride + out_base;
  for (uint32_t l = threadIdx.x; l < n_tile; l += blockDim.x) dp[scat(l, h.out_loc, h.L)] = tile[l];
}

// Tiled complex GEMM with the permutation folded into the gathers: a CTA owns a 64 x 64 tile of
// C[m, n] for one kept-shared index value and one parameter set; A and B tiles are gathered through
// per-CTA offset tables into shared memory (K tile = 16), each thread accumulates a 4 x 4 block.
constexpr int TM = 64, TN = 64, TK = 16;

template <typename R>
__global__ void __launch_bounds__(256)
k_tn_gemm(const cx<R>* __restrict__ A, int64_t sA, const cx<R>* __restrict__ B, int64_t sB, cx<R>* __restrict__ C,
          int64_t sC, const __grid_constant__ StepDev d) {
  __shared__ cx<R> As[TK][TM + 2];
  __shared__ cx<R> Bs[TK][TN + 2];
  __shared__ uint32_t offm[TM], offn[TN];
  const int tiles_n = 1 << (d.n_n - 6);
  const int tiles_m = 1 << (d.n_m - 6);
  uint32_t bid = blockIdx.x;
  const uint32_t tn = bid % tiles_n;
  bid /= tiles_n;
  const uint32_t tm = bid % tiles_m;
  const uint32_t bb = bid / tiles_m;
  const int64_t set = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < TM) offm[tid] = scat(tm * TM + tid, d.a_m, d.n_m) | scat(bb, d.a_b, d.n_b);
  if (tid >= 64 && tid < 64 + TN) offn[tid - 64] = scat(tn * TN + (tid - 64), d.b_n, d.n_n) | scat(bb, d.b_b, d.n_b);
  __syncthreads();
  const cx<R>* a = A + set * sA;
  const cx<R>* b = B + set * sB;
  const int tx = tid & 15, ty = tid >> 4;
  cx<R> acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = mk<R>(0, 0);
  const int K = 1 << d.n_k;
  for (int k0 = 0; k0 < K; k0 += TK) {
    // gather: 64 x 16 elements per operand, 4 per thread; the index that is contiguous in memory varies fastest
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int l = tid + e * 256;
      int mi, ki;
      if (d.a_k_fast) {
        ki = l & (TK - 1);
        mi = l >> 4;
      } else {
        mi = l & (TM - 1);
        ki = l >> 6;
      }
      As[ki][mi] = a[offm[mi] + d.ka[k0 + ki]];
      int ni, kj;
      if (d.b_k_fast) {
        kj = l & (TK - 1);
        ni = l >> 4;
      } else {
        ni = l & (TN - 1);
        kj = l >> 6;
      }
      Bs[kj][ni] = b[offn[ni] + d.kb[k0 + kj]];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      cx<R> av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = cfma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  cx<R>* c = C + set * sC + ((int64_t)bb << (d.n_m + d.n_n));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = (int64_t)tm * TM + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) c[(m << d.n_n) + tn * TN + tx * 4 + j] = acc[i][j];
  }
}

// complex128 GEMM-shaped steps on the FP64 tensor cores (DMMA): same tiling and gathers as k_tn_gemm (a CTA owns
// a 64 x 64 tile of C[m, n]; permutation folded into the gathers; K tile = 16), operands staged PLANAR (re / im)
// in shared memory, each warp owns 16 x 32 outputs = 2 x 4 mma.m8n8k4 blocks and issues the 4M real products
//   Cr += Ar Br + (-Ai) Bi,   Ci += Ar Bi + Ai Br
// as four mma.sync.m8n8k4.f64 per block and k4 step; fp64 accumulation in registers (round-to-nearest).
constexpr int DLD = 68;  // leading dimension of the staged tiles: 2*DLD mod 32 == 8 -> conflict-free fragment loads
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256, 2)
k_tn_gemm_dmma(const cx<double>* __restrict__ A, int64_t sA, const cx<double>* __restrict__ B, int64_t sB,
               cx<double>* __restrict__ C, int64_t sC, const __grid_constant__ StepDev d) {
  __shared__ double As_re[TK][DLD], As_im[TK][DLD], Bs_re[TK][DLD], Bs_im[TK][DLD];
  __shared__ uint32_t offm[TM], offn[TN];
  const int tiles_n = 1 << (d.n_n - 6);
  const int tiles_m = 1 << (d.n_m - 6);
  uint32_t bid = blockIdx.x;
  const uint32_t tn = bid % tiles_n;
  bid /= tiles_n;
  const uint32_t tm = bid % tiles_m;
  const uint32_t bb = bid / tiles_m;
  const int64_t set = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < TM) offm[tid] = scat(tm * TM + tid, d.a_m, d.n_m) | scat(bb, d.a_b, d.n_b);
  if (tid >= 64 && tid < 64 + TN) offn[tid - 64] = scat(tn * TN + (tid - 64), d.b_n, d.n_n) | scat(bb, d.b_b, d.n_b);
  __syncthreads();
  const cx<double>* a = A + set * sA;
  const cx<double>* b = B + set * sB;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;  // warp tile origin inside the CTA tile
  const int fr = lane >> 2, fk = lane & 3;                // fragment row (or column) / k index of this lane
  double cr[2][4][2], ci[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
  const int K = 1 << d.n_k;
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int l = tid + e * 256;
      int mi, ki;
      if (d.a_k_fast) {
        ki = l & (TK - 1);
        mi = l >> 4;
      } else {
        mi = l & (TM - 1);
        ki = l >> 6;
      }
      const cx<double> av = a[offm[mi] + d.ka[k0 + ki]];
      As_re[ki][mi] = av.x;
      As_im[ki][mi] = av.y;
      int ni, kj;
      if (d.b_k_fast) {
        kj = l & (TK - 1);
        ni = l >> 4;
      } else {
        ni = l & (TN - 1);
        kj = l >> 6;
      }
      const cx<double> bv = b[offn[ni] + d.kb[k0 + kj]];
      Bs_re[kj][ni] = bv.x;
      Bs_im[kj][ni] = bv.y;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < TK; k4 += 4) {
      double ar[2], ai[2], nai[2], br[4], bi[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ar[i] = As_re[k4 + fk][wm + i * 8 + fr];
        ai[i] = As_im[k4 + fk][wm + i * 8 + fr];
        nai[i] = -ai[i];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        br[j] = Bs_re[k4 + fk][wn + j * 8 + fr];
        bi[j] = Bs_im[k4 + fk][wn + j * 8 + fr];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma884(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
          dmma884(cr[i][j][0], cr[i][j][1], nai[i], bi[j]);
          dmma884(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
          dmma884(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
        }
    }
    __syncthreads();
  }
  cx<double>* c = C + set * sC + ((int64_t)bb << (d.n_m + d.n_n));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int64_t m = (int64_t)tm * TM + wm + i * 8 + fr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      cx<double>* dst = c + (m << d.n_n) + tn * TN + wn + j * 8 + 2 * fk;
      dst[0] = mk<double>(cr[i][j][0], ci[i][j][0]);
      dst[1] = mk<double>(cr[i][j][1], ci[i][j][1]);
    }
  }
}

// operands of a simplified network: dst[set][i] = (idx[i] >= 0 ? gates[set][idx[i]] : adjoints[set][-idx[i]-1])
template <typename R>
__global__ void k_tn_gather(const cx<R>* __restrict__ g, const cx<R>* __restrict__ a, int64_t src_stride,
                            const int32_t* __restrict__ idx, int64_t n, cx<R>* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t set = blockIdx.y;
  const int32_t j = idx[i];
  dst[set * n + i] = j >= 0 ? g[set * src_stride + j] : a[set * src_stride + (-j - 1)];
}

// out[set][perm(o)] += last[set][o]
struct FinalDev {
  int32_t rank;
  int8_t pos[TQ_TN_MAX_RANK];  // bit j of the result goes to output bit pos[j]
};
template <typename R>
__global__ void k_tn_final(const cx<R>* __restrict__ last, int64_t sL, cx<R>* __restrict__ out, int64_t sO,
                           const __grid_constant__ FinalDev f, int64_t n) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const int64_t set = blockIdx.y;
  cx<R> v = last[set * sL + o];
  cx<R>* dst = out + set * sO + scat((uint32_t)o, f.pos, f.rank);
  dst->x += v.x;
  dst->y += v.y;
}

// gradient seed of the reverse pass: g[set][o] = conj(grad_out[set][perm(o)]) in the layout of the last forward tensor
template <typename R>
__global__ void k_tn_seed(const cx<R>* __restrict__ gout, int64_t sO, cx<R>* __restrict__ g, int64_t sG,
                          const __grid_constant__ FinalDev f, int64_t n) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const int64_t set = blockIdx.y;
  const cx<R> v = gout[set * sO + scat((uint32_t)o, f.pos, f.rank)];
  g[set * sG + o] = mk<R>(v.x, -v.y);
}


constexpr int DOT_MAX_OUT_LOG2 = 6, DOT_MIN_K_LOG2 = 12, DOT_BLOCKS = 128;
// a step joins a fused run when k+m+n+b <= this (one CTA per parameter set does the whole step)
constexpr int FUSE_MAX_WORK_BATCHED = 18, FUSE_MAX_WORK_SHARED = 14, FUSE_CTA_WORK = 8, FUSE_MAX_OUT_LOG2 = 12;

struct SchedItem {
  int step = -1;                       // >= 0: one step on its own kernel(s)
  int fs_begin = 0, n_fsteps = 0;      // else: a fused run = d_fsteps[fs_begin .. +n_fsteps)
  int lv_begin = 0, n_levels = 0;      // its level table: d_levels[lv_begin .. +n_levels+1) offsets (relative to
                                       // fs_begin), then n_levels cta-step counts
  bool batched = false;
  std::vector<int> members;            // steps of the run, in execution order
  int chain = -1;                      // >= 0: an apply-chain run (k_tn_chain) = chains[chain]; members = its steps
};

struct ChainRun {
  ChainHdr hdr;
  int step_begin = 0;   // first entry of its ChainStep table inside d_chain_steps
  int big_in = -1;      // tensor id of the first step's large operand
  int last = -1;        // last step of the run (its result is what the launch writes)
  int rank = 0;         // rank of the large tensor
};

// Tensor-core lowering of one step: which operand provides accumulator rows, image geometry, pack tables.
struct TcStep {
  bool shape_ok = false;  // rows >= 128, columns >= 16
  bool swap = false;      // accumulator rows come from the rhs (its free indices outnumber the lhs's)
  int c_t = 0, kblocks = 0, tiles_a = 0, tiles_b = 0, stages = 0;
  int64_t img_a_z = 0, img_b_z = 0;  // image bytes per z (= parameter set x kept-shared index value)
  // a slice-invariant operand of a per-slice step is packed ONCE per call into a pinned image
  bool pin_a = false, pin_b = false;
  // the row operand is streamed (every tile read once or twice): when its image is not pinned the GEMM kernel
  // gathers it itself (k_tc_gemm<C_T, true>), no image is written
  bool gather_a = false;
  tc::PackParams pa, pb;
  // fused pack: this step's epilogue writes operand image `fuse_is_b ? B : A` of step fuse_to (its only consumer);
  // src_a / src_b: the step that writes this step's A / B image that way (-1: packed from the plain tensor)
  int fuse_to = -1, src_a = -1, src_b = -1;
  int fuse_mode = 0;  // 1: the image's low k bits are accumulator ROW bits (8-byte stores, 8 lanes per 64-byte piece),
                      // 2: lowest k bit a column bit, the next two row bits (16-byte stores), 3: all three column bits
  bool fuse_is_b = false;
  // index orders of this step's GEMM: position t of the accumulator rows / columns / contracted extent is default
  // position ord_*[t] (default = the lowering's order).  A step that stores a plain result keeps its row / column
  // order (the result layout is [b | m | n]); a fused-pack producer's rows and columns are ordered so that its
  // epilogue's stores fall into whole 64-byte row pieces of the consumer's image.
  std::vector<int> ord_row, ord_col, ord_k;
  int64_t out_entries = 0;  // size of the image in complex entries per parameter set
  tc::ImgOut out;
};

static void tc_pack_tables(tc::PackParams& P, const int8_t* row_bits, int n_row, const int8_t* k_bits, int n_k,
                           const int8_t* b_bits, int n_b, bool is_b, int rows_t_log2) {
  memset(&P, 0, sizeof(P));
  P.n_row = n_row;
  P.n_k = n_k;
  P.n_b = n_b;
  P.is_b = is_b ? 1 : 0;
  P.rows_t_log2 = rows_t_log2;
  P.kblocks = n_k > tc::KB_LOG ? 1 << (n_k - tc::KB_LOG) : 1;
  for (int j = 0; j < n_row; ++j) P.row_bits[j] = row_bits[j];
  for (int j = 0; j < n_k; ++j) P.k_bits[j] = k_bits[j];
  for (int j = 0; j < n_b; ++j) P.b_bits[j] = b_bits[j];
  std::vector<std::pair<int, int>> loc;  // (source bit, value in r << 4 | kk)
  for (int j = 0; j < rows_t_log2; ++j) loc.push_back({row_bits[j], 1 << (4 + j)});
  for (int j = 0; j < std::min(n_k, tc::KB_LOG); ++j) loc.push_back({k_bits[j], 1 << j});
  std::sort(loc.begin(), loc.end());
  P.n_local = (int)loc.size();
  for (size_t j = 0; j < loc.size(); ++j) {
    P.local_src[j] = (int8_t)loc[j].first;
    P.local_dst[j] = (int16_t)loc[j].second;
  }
}

}  // namespace tq

using namespace tq;

struct tq_tn_plan;
static void tc_build_tables(tq_tn_plan* p, int s);

struct tq_tn_plan {
  int dtype = TQ_C64, n_in = 0, n_out = 0, n_sliced = 0;
  int device = -1;  // CUDA device that owns the plan's tables
  std::vector<tq_tn_step> steps;
  std::vector<int> slice_tensor, slice_ord, slice_bit, final_perm;
  std::vector<int> in_rank;
  std::vector<char> in_batched;
  // per step
  std::vector<char> dep_batch, dep_slice;
  std::vector<int64_t> arena_off;  // element offset of every step's output inside its arena
  std::vector<char> arena_const;   // 1: shared arena, 0: per-set arena
  int64_t arena_set = 0, arena_shared = 0;
  std::vector<int32_t*> d_ka, d_kb;
  std::vector<StepDev> dev;
  double flops = 0;
  int width = 0;
  // schedule: what runs, in which order (rebuilt when an option changes)
  std::vector<int> kind;            // per step: 0 per-element, 1 FMA GEMM, 2 tcgen05 GEMM, 3 split-K, 4 fused run,
                                    //           5 apply, 6 gradient seed
  std::vector<SchedItem> items[3];  // phase 0: once per call; phase 1: every slice; phase 2: backward pass
  std::vector<char> phase;          // per step
  std::vector<char> t_slice;        // per tensor id: depends on a sliced index
  std::vector<char> t_batch;        // per tensor id: differs per parameter set
  std::vector<int> t_rank;
  // reverse mode (tq_tn_plan_enable_backward): forward steps [0, n_fwd), then the seed step, then two
  // contractions per forward step (conjugated gradients, see tq_tn_backward)
  int n_fwd = 0;                    // number of forward steps (== steps.size() until backward is enabled)
  int seed_step = -1;
  std::vector<std::vector<int>> layout;  // per tensor id: index id of every bit (fast -> slow)
  std::vector<int> grad_of;              // per tensor id: tensor id of its (conjugated) gradient, -1 if none
  std::vector<ApplyDev> apply;      // per step (kind 5 only)
  std::vector<char> apply_small_rhs;
  FusedStep* d_fsteps = nullptr;
  uint32_t* d_micro = nullptr;      // offset tables of the micro steps
  int32_t* d_levels = nullptr;      // [level_off (n+1 per run) | level_ncta] blocks, see SchedItem
  int fuse_enabled = 1;             // TQ_TN_OPT_FUSE_SMALL
  // tensor-core path (complex64 only): per-step operand-image descriptions
  std::vector<TcStep> tc;
  int tc_enabled = 1;      // TQ_TN_OPT_TENSOR_CORE
  int tc_min_log2 = 20;    // TQ_TN_OPT_TC_MIN_LOG2: a step runs on tensor cores when k+m+n+b >= this
  int tc_chunk = 32;       // TQ_TN_OPT_TC_CHUNK: complex k accumulated in TMEM between round-to-nearest drains
  int tc_splitk = 1;       // TQ_TN_OPT_TC_SPLITK
  int tc_gather = 0;       // TQ_TN_OPT_TC_GATHER
  int tc_fuse_pack = 1;    // TQ_TN_OPT_TC_FUSE_PACK
  int chain_enabled = 1;   // TQ_TN_OPT_CHAIN
  std::vector<ChainRun> chains;
  ChainStep* d_chain_steps = nullptr;
  int num_sms = 148;
};

// Default (lowering-order) physical bit lists of a step's two GEMM operands: A provides the accumulator rows.
struct TcSides {
  const int8_t *a_free, *a_k, *a_b, *b_free, *b_k, *b_b;
  int n_row, n_col;
};
static TcSides tc_sides(const tq_tn_step& st, bool swap) {
  const int8_t* lhs_k = st.lhs_bits;
  const int8_t* lhs_m = st.lhs_bits + st.n_k;
  const int8_t* lhs_b = st.lhs_bits + st.n_k + st.n_m;
  const int8_t* rhs_k = st.rhs_bits;
  const int8_t* rhs_n = st.rhs_bits + st.n_k;
  const int8_t* rhs_b = st.rhs_bits + st.n_k + st.n_n;
  if (!swap) return TcSides{lhs_m, lhs_k, lhs_b, rhs_n, rhs_k, rhs_b, st.n_m, st.n_n};
  return TcSides{rhs_n, rhs_k, rhs_b, lhs_m, lhs_k, lhs_b, st.n_n, st.n_m};
}

// pack tables of both operand images of step s for its current row / column / k orders
static void tc_build_tables(tq_tn_plan* p, int s) {
  const tq_tn_step& st = p->steps[s];
  TcStep& T = p->tc[s];
  const TcSides S = tc_sides(st, T.swap);
  int8_t rowb[TQ_TN_MAX_RANK], colb[TQ_TN_MAX_RANK], ka[TQ_TN_MAX_RANK], kb[TQ_TN_MAX_RANK];
  for (int i = 0; i < S.n_row; ++i) rowb[i] = S.a_free[T.ord_row[i]];
  for (int i = 0; i < S.n_col; ++i) colb[i] = S.b_free[T.ord_col[i]];
  for (int i = 0; i < st.n_k; ++i) {
    ka[i] = S.a_k[T.ord_k[i]];
    kb[i] = S.b_k[T.ord_k[i]];
  }
  tc_pack_tables(T.pa, rowb, S.n_row, ka, st.n_k, S.a_b, st.n_b, false, 7);
  tc_pack_tables(T.pb, colb, S.n_col, kb, st.n_k, S.b_b, st.n_b, true, std::min(S.n_col, 7));
}

// Per-step derived data for steps [first, end): dependency flags, ranks, layouts, K tables, tensor-core lowering.
static int setup_steps(tq_tn_plan* p, int first) {
  const int n_in = p->n_in, n_steps = (int)p->steps.size();
  p->dep_batch.resize(n_steps);
  p->dep_slice.resize(n_steps);
  p->arena_const.resize(n_steps);
  p->t_batch.resize(n_in + n_steps);
  p->t_slice.resize(n_in + n_steps);
  p->t_rank.resize(n_in + n_steps);
  p->layout.resize(n_in + n_steps);
  for (int s = first; s < n_steps; ++s) {
    const tq_tn_step& st = p->steps[s];
    const int o = n_in + s;
    p->t_batch[o] = p->t_batch[st.lhs] || p->t_batch[st.rhs];
    p->t_slice[o] = p->t_slice[st.lhs] || p->t_slice[st.rhs];
    p->t_rank[o] = s == p->seed_step ? p->n_out : st.n_m + st.n_n + st.n_b;
    p->dep_batch[s] = p->t_batch[o];
    p->dep_slice[s] = p->t_slice[o];
    p->arena_const[s] = !p->t_batch[o];
    p->width = std::max(p->width, p->t_rank[o]);
    p->layout[o].assign(st.out_idx, st.out_idx + p->t_rank[o]);
  }
  // device step tables
  p->dev.resize(n_steps);
  p->d_ka.resize(n_steps, nullptr);
  p->d_kb.resize(n_steps, nullptr);
  for (int s = first; s < n_steps; ++s) {
    if (s == p->seed_step) continue;
    const tq_tn_step& st = p->steps[s];
    StepDev& d = p->dev[s];
    memset(&d, 0, sizeof(d));
    d.n_k = st.n_k;
    d.n_m = st.n_m;
    d.n_n = st.n_n;
    d.n_b = st.n_b;
    TQ_REQUIRE(st.n_k <= 24 || st.n_m + st.n_n + st.n_b <= DOT_MAX_OUT_LOG2, TQ_E_UNSUPPORTED,
               "tq_tn_plan_create: step %d contracts 2^%d terms", s, st.n_k);
    for (int j = 0; j < st.n_m; ++j) d.a_m[j] = st.lhs_bits[st.n_k + j];
    for (int j = 0; j < st.n_b; ++j) d.a_b[j] = st.lhs_bits[st.n_k + st.n_m + j];
    for (int j = 0; j < st.n_n; ++j) d.b_n[j] = st.rhs_bits[st.n_k + j];
    for (int j = 0; j < st.n_b; ++j) d.b_b[j] = st.rhs_bits[st.n_k + st.n_n + j];
    for (int j = 0; j < st.n_k; ++j) {
      d.a_k[j] = st.lhs_bits[j];
      d.b_k[j] = st.rhs_bits[j];
    }
    const bool is_dot = st.n_m + st.n_n + st.n_b <= DOT_MAX_OUT_LOG2 && st.n_k >= DOT_MIN_K_LOG2;
    const int K = 1 << (is_dot ? DOT_LO : st.n_k);
    std::vector<int32_t> ka(K), kb(K);
    for (int k = 0; k < K; ++k) {
      uint32_t a = 0, b = 0;
      for (int j = 0; j < st.n_k; ++j)
        if ((k >> j) & 1) {
          a |= 1u << st.lhs_bits[j];
          b |= 1u << st.rhs_bits[j];
        }
      ka[k] = (int32_t)a;
      kb[k] = (int32_t)b;
    }
    for (int j = 0; j < st.n_k; ++j) {
      if (st.lhs_bits[j] == 0) d.a_k_fast = 1;
      if (st.rhs_bits[j] == 0) d.b_k_fast = 1;
    }
    TQ_CUDA_OK(cudaMalloc((void**)&p->d_ka[s], K * sizeof(int32_t)));
    TQ_CUDA_OK(cudaMalloc((void**)&p->d_kb[s], K * sizeof(int32_t)));
    TQ_CUDA_OK(cudaMemcpy(p->d_ka[s], ka.data(), K * sizeof(int32_t), cudaMemcpyHostToDevice));
    TQ_CUDA_OK(cudaMemcpy(p->d_kb[s], kb.data(), K * sizeof(int32_t), cudaMemcpyHostToDevice));
    d.ka = p->d_ka[s];
    d.kb = p->d_kb[s];
  }
  // tensor-core lowering (complex64): the operand with more free indices provides the 128 accumulator rows
  p->tc.resize(n_steps);
  for (int s = first; s < n_steps && p->dtype == TQ_C64; ++s) {
    if (s == p->seed_step) continue;
    const tq_tn_step& st = p->steps[s];
    TcStep& T = p->tc[s];
    // Accumulator rows (2x image expansion) come from the operand with more free indices, unless exactly one
    // operand changes per slice and is tall enough: then that one takes the rows (its per-slice image is the
    // cheaper one to write) and the invariant operand's 4x image is packed once.
    const bool lhs_var = p->t_slice[st.lhs] != 0, rhs_var = p->t_slice[st.rhs] != 0;
    T.swap = st.n_n > st.n_m;
    if (lhs_var != rhs_var && std::min(st.n_m, st.n_n) >= 7) T.swap = rhs_var;
    const int n_row = T.swap ? st.n_n : st.n_m, n_col = T.swap ? st.n_m : st.n_n;
    {
      const bool a_var = T.swap ? rhs_var : lhs_var, b_var = T.swap ? lhs_var : rhs_var;
      T.pin_a = p->dep_slice[s] && !a_var;
      T.pin_b = p->dep_slice[s] && !b_var;
      if (getenv("TQ_TN_NO_PIN")) T.pin_a = T.pin_b = false;  // experiments: repack invariant operands every slice
    }
    T.shape_ok = n_row >= 7 && n_col >= 4 && n_row + n_col + st.n_b <= 31;
    if (!T.shape_ok) continue;
    const int col_t_log2 = std::min(n_col, 7);
    T.c_t = 1 << col_t_log2;
    T.kblocks = st.n_k > tc::KB_LOG ? 1 << (st.n_k - tc::KB_LOG) : 1;
    T.tiles_a = 1 << (n_row - 7);
    T.tiles_b = 1 << (n_col - col_t_log2);
    T.stages = tc::num_stages(T.c_t);
    T.gather_a = st.n_k >= tc::KB_LOG && T.tiles_b <= 2;
    T.img_a_z = (int64_t)T.tiles_a * T.kblocks * tc::A_CHUNK;
    T.img_b_z = (int64_t)T.tiles_b * T.kblocks * tc::b_chunk_bytes(T.c_t);
    T.ord_row.resize(n_row);
    T.ord_col.resize(n_col);
    T.ord_k.resize(st.n_k);
    for (int i = 0; i < n_row; ++i) T.ord_row[i] = i;
    for (int i = 0; i < n_col; ++i) T.ord_col[i] = i;
    for (int i = 0; i < st.n_k; ++i) T.ord_k[i] = i;
    tc_build_tables(p, s);
  }
  return TQ_OK;
}

// Decide which kernel runs every step, group the small ones into fused runs, fix the execution order of both
// phases (once per call / every slice) and lay the intermediates out in the two arenas for that order.
static int build_schedule(tq_tn_plan* p) {
  const int n_in = p->n_in, n_steps = (int)p->steps.size();
  std::vector<int> in_slice_bits(n_in, 0);
  for (int t : p->slice_tensor) in_slice_bits[t]++;
  p->kind.assign(n_steps, 0);
  p->apply.resize(n_steps);
  p->apply_small_rhs.assign(n_steps, 0);
  for (int s = 0; s < n_steps; ++s) {
    const tq_tn_step& st = p->steps[s];
    const int work = st.n_k + st.n_m + st.n_n + st.n_b, outl = st.n_m + st.n_n + st.n_b;
    const bool tc_ok = p->dtype == TQ_C64 && p->tc_enabled && p->tc[s].shape_ok && work >= p->tc_min_log2;
    auto opnd_ok = [&](int t, int rank) {
      return rank <= FUSE_MAX_RANK && (t >= n_in || in_slice_bits[t] <= FUSE_MAX_SLICE_BITS);
    };
    // fused: little work AND a small result (a 2^16-element state update is a whole-GPU job, not one CTA's)
    const bool small = p->fuse_enabled && work <= (p->dep_batch[s] ? FUSE_MAX_WORK_BATCHED : FUSE_MAX_WORK_SHARED) &&
                       outl <= FUSE_MAX_OUT_LOG2 &&
                       opnd_ok(st.lhs, st.n_k + st.n_m + st.n_b) && opnd_ok(st.rhs, st.n_k + st.n_n + st.n_b);
    const bool lhs_small = st.n_m <= APPLY_MAX && st.n_k <= APPLY_MAX, rhs_small = st.n_n <= APPLY_MAX && st.n_k <= APPLY_MAX;
    const bool apply_ok = (lhs_small || rhs_small) && outl >= 10 && st.n_b <= 8;
    p->kind[s] = tc_ok ? 2 : small ? 4 : (outl <= DOT_MAX_OUT_LOG2 && st.n_k >= DOT_MIN_K_LOG2) ? 3
                 : (st.n_m >= 6 && st.n_n >= 6 && st.n_k >= 4) ? 1 : apply_ok ? 5 : 0;
    if (s == p->seed_step) p->kind[s] = 6;
    if (p->kind[s] == 5) {
      const bool small_rhs = !lhs_small || (rhs_small && st.n_n < st.n_m);
      ApplyDev& a = p->apply[s];
      memset(&a, 0, sizeof(a));
      p->apply_small_rhs[s] = small_rhs;
      const int8_t* sb = small_rhs ? st.rhs_bits : st.lhs_bits;
      const int8_t* bbits = small_rhs ? st.lhs_bits : st.rhs_bits;
      a.n_k = st.n_k;
      a.n_b = st.n_b;
      a.n_s = small_rhs ? st.n_n : st.n_m;
      a.n_f = small_rhs ? st.n_m : st.n_n;
      a.s_shift = small_rhs ? 0 : st.n_n;
      a.f_shift = small_rhs ? st.n_n : 0;
      for (int j = 0; j < st.n_k; ++j) {
        a.small_k[j] = sb[j];
        a.big_k[j] = bbits[j];
      }
      for (int j = 0; j < a.n_s; ++j) a.small_s[j] = sb[st.n_k + j];
      for (int j = 0; j < a.n_f; ++j) a.big_f[j] = bbits[st.n_k + j];
      for (int j = 0; j < st.n_b; ++j) {
        a.small_b[j] = sb[st.n_k + a.n_s + j];
        a.big_b[j] = bbits[st.n_k + a.n_f + j];
      }
    }
  }
  // ---- fused pack: a tensor-core step whose result is read by exactly one step, itself a tensor-core step of the
  // same phase with the same batch / slice dependence, writes that step's operand image from its epilogue.
  // Steps are visited consumers first: a consumer d picks the three contracted indices that become the low k bits
  // of its image (one 64-byte row piece = 8 consecutive k) among those its producer c holds as accumulator ROW bits
  // (lanes of a warp: 8 lanes x 8 bytes fill a piece) or, failing that, as COLUMN bits (consecutive registers of one
  // thread: 16-byte stores), and fixes c's row / column order accordingly; c's own k order is chosen when c is
  // visited as a consumer.
  {
    static_assert(tc::KB_LOG == 3, "the image address maps below assume 64-byte rows");
    std::vector<int> n_cons(n_in + n_steps, 0);
    for (int s = 0; s < n_steps; ++s) {
      n_cons[p->steps[s].lhs] += 1;
      if (s != p->seed_step) n_cons[p->steps[s].rhs] += 1;
    }
    for (int s = 0; s < n_steps; ++s) {
      TcStep& T = p->tc[s];
      T.fuse_to = T.src_a = T.src_b = -1;
      T.out_entries = 0;
      for (size_t i = 0; i < T.ord_row.size(); ++i) T.ord_row[i] = (int)i;
      for (size_t i = 0; i < T.ord_col.size(); ++i) T.ord_col[i] = (int)i;
      for (size_t i = 0; i < T.ord_k.size(); ++i) T.ord_k[i] = (int)i;
    }
    auto front = [](std::vector<int>& ord, const std::vector<int>& first) {  // `first`, then the rest in old order
      std::vector<int> out(first);
      for (int v : ord)
        if (std::find(first.begin(), first.end(), v) == first.end()) out.push_back(v);
      ord.swap(out);
    };
    for (int d = n_steps - 1; d >= 0 && p->tc_fuse_pack; --d) {
      if (p->kind[d] != 2 || !p->tc[d].shape_ok) continue;
      const tq_tn_step& sd = p->steps[d];
      TcStep& Td = p->tc[d];
      if (sd.n_k < tc::KB_LOG) continue;  // a padded k-block has zero columns only the pack kernel writes
      bool k_fixed = false;
      for (int pass = 0; pass < 2 && !k_fixed; ++pass) {  // the row operand (A image) first
        const int side = (pass == 0) == !Td.swap ? 0 : 1;  // 0: lhs, 1: rhs
        const bool is_b = pass == 1;
        const int t = side == 0 ? sd.lhs : sd.rhs;
        if (t < n_in || n_cons[t] != 1) continue;
        const int c = t - n_in;
        if (p->kind[c] != 2 || !p->tc[c].shape_ok || p->phase[c] != p->phase[d] ||
            p->dep_batch[c] != p->dep_batch[d] || p->dep_slice[c] != p->dep_slice[d] || p->tc[c].fuse_to >= 0)
          continue;
        if (is_b ? Td.pin_b : Td.pin_a) continue;
        const int64_t img_z = is_b ? Td.img_b_z : Td.img_a_z;
        if ((img_z << sd.n_b) >= ((int64_t)1 << 32)) continue;
        const tq_tn_step& sc = p->steps[c];
        TcStep& Tc = p->tc[c];
        if (sc.n_b > 8) continue;
        // role of every bit of the tensor inside the producer's GEMM: default row / column position
        int prow[TQ_TN_MAX_RANK], pcol[TQ_TN_MAX_RANK];
        for (int b = 0; b < TQ_TN_MAX_RANK; ++b) prow[b] = pcol[b] = -1;
        for (int i = 0; i < sc.n_n; ++i) (Tc.swap ? prow : pcol)[i] = i;
        for (int i = 0; i < sc.n_m; ++i) (Tc.swap ? pcol : prow)[sc.n_n + i] = i;
        const int8_t* bits_t = side == 0 ? sd.lhs_bits : sd.rhs_bits;  // [k..., free..., b...] positions in the tensor
        std::vector<int> kr, kc;  // contracted indices of d (default k positions) held as rows / columns by c
        for (int j = 0; j < sd.n_k; ++j) {  // (an index the producer kept as a batch index is neither)
          if (prow[bits_t[j]] >= 0) kr.push_back(j);
          else if (pcol[bits_t[j]] >= 0) kc.push_back(j);
        }
        std::vector<int> kk;       // the three low k bits of d's images
        std::vector<int> c_rows_first, c_cols_first;
        int mode = kr.size() >= 3 ? 1 : (kr.size() == 2 && !kc.empty()) ? 2 : 3;
        if (kr.size() >= 3) {
          kk = {kr[0], kr[1], kr[2]};
          for (int j : kk) c_rows_first.push_back(prow[bits_t[j]]);
        } else if (kr.size() == 2 && !kc.empty()) {
          kk = {kc[0], kr[0], kr[1]};
          c_cols_first.push_back(pcol[bits_t[kc[0]]]);
          c_rows_first = {prow[bits_t[kr[0]]], prow[bits_t[kr[1]]]};
        } else if (kc.size() >= 3) {
          kk = {kc[0], kc[1], kc[2]};
          for (int j : kk) c_cols_first.push_back(pcol[bits_t[j]]);
        } else {
          continue;
        }
        // after the k bits: the image's row order, so that neighbouring lanes / registers hit neighbouring rows
        const std::vector<int>& d_rows = is_b ? Td.ord_col : Td.ord_row;
        for (int pos : d_rows) {
          const int bit = bits_t[sd.n_k + pos];
          if (prow[bit] >= 0) c_rows_first.push_back(prow[bit]);
          else if (pcol[bit] >= 0) c_cols_first.push_back(pcol[bit]);
        }
        front(Td.ord_k, kk);
        front(Tc.ord_row, c_rows_first);
        front(Tc.ord_col, c_cols_first);
        Tc.fuse_to = d;
        Tc.fuse_mode = mode;
        Tc.fuse_is_b = is_b;
        Tc.out_entries = (img_z << sd.n_b) / 8;
        (is_b ? Td.src_b : Td.src_a) = c;
        k_fixed = true;
      }
    }
    for (int s = 0; s < n_steps; ++s)
      if (p->dtype == TQ_C64 && s != p->seed_step && p->tc[s].shape_ok) tc_build_tables(p, s);
    // byte-offset maps of every fused producer (the consumers' tables are final now)
    for (int c = 0; c < n_steps; ++c) {
      TcStep& Tc = p->tc[c];
      if (Tc.fuse_to < 0) continue;
      const int d = Tc.fuse_to;
      const tq_tn_step& sd = p->steps[d];
      const tq_tn_step& sc = p->steps[c];
      const TcStep& Td = p->tc[d];
      const bool is_b = Tc.fuse_is_b;
      const tc::PackParams& P = is_b ? Td.pb : Td.pa;
      const int64_t img_z = is_b ? Td.img_b_z : Td.img_a_z;
      const uint32_t chunk = is_b ? (uint32_t)tc::b_chunk_bytes(Td.c_t) : (uint32_t)tc::A_CHUNK;
      const uint32_t tile_stride = (uint32_t)Td.kblocks * chunk;
      uint32_t contrib[TQ_TN_MAX_RANK];
      for (int j = 0; j < TQ_TN_MAX_RANK; ++j) contrib[j] = 0;
      for (int j = 0; j < P.n_k; ++j)
        contrib[P.k_bits[j]] = j == 0 ? 8u : j == 1 ? 16u : j == 2 ? 32u : (1u << (j - 3)) * chunk;
      for (int j = 0; j < P.n_row; ++j)
        contrib[P.row_bits[j]] = j < P.rows_t_log2 ? (((1u << j) * 64u) ^ (j == 1 ? 16u : j == 2 ? 32u : 0u))
                                                   : (1u << (j - P.rows_t_log2)) * tile_stride;
      for (int j = 0; j < P.n_b; ++j) contrib[P.b_bits[j]] = (1u << j) * (uint32_t)img_z;
      const int n_row_c = Tc.swap ? sc.n_n : sc.n_m, n_col_c = Tc.swap ? sc.n_m : sc.n_n;
      tc::ImgOut& O = Tc.out;
      memset(&O, 0, sizeof(O));
      // accumulator row bit i = default row position ord_row[i] = result bit (n bits first, then m bits)
      for (int i = 0; i < n_row_c; ++i) O.rmap[i] = contrib[Tc.swap ? Tc.ord_row[i] : sc.n_n + Tc.ord_row[i]];
      for (int i = 0; i < n_col_c; ++i) O.cmap[i] = contrib[Tc.swap ? sc.n_n + Tc.ord_col[i] : Tc.ord_col[i]];
      for (int i = 0; i < sc.n_b; ++i) O.bmap[i] = contrib[sc.n_m + sc.n_n + i];
      O.plane = is_b ? (uint32_t)tc::b_plane_bytes(Td.c_t) : (uint32_t)tc::A_PLANE;
      O.im_off = is_b ? (uint32_t)(Td.c_t * tc::ROW_BYTES) : 0u;
      O.n_row = n_row_c;
      O.n_col = n_col_c;
      O.n_b = sc.n_b;
    }
  }
  // ---- execution order
  std::vector<char> done(n_in + n_steps, 0);
  for (int t = 0; t < n_in; ++t) done[t] = 1;
  std::vector<FusedStep> fsteps;
  std::vector<int32_t> levels;
  std::vector<uint32_t> microtab;
  // ---- apply-chain runs (k_tn_chain): consecutive gate-like apply steps on the same large tensor
  p->chains.clear();
  std::vector<ChainStep> chain_steps;
  std::vector<int> chain_small;  // small operand (tensor id) of every chain step: arena offsets are filled in later
  std::vector<int> consumer_of(n_in + n_steps, -1), n_consumers(n_in + n_steps, 0);
  for (int s = 0; s < n_steps; ++s) {
    for (int t : {p->steps[s].lhs, p->steps[s].rhs}) {
      consumer_of[t] = s;
      n_consumers[t] += 1;
    }
  }
  auto gate_like = [&](int s) {
    if (p->kind[s] != 5 || s == p->seed_step) return false;
    const ApplyDev& a = p->apply[s];
    return a.n_k >= 1 && a.n_k <= CHAIN_MAX_K && a.n_s == a.n_k && a.n_b <= CHAIN_MAX_B;
  };
  auto big_of = [&](int s) { return p->apply_small_rhs[s] ? p->steps[s].lhs : p->steps[s].rhs; };
  auto small_of = [&](int s) { return p->apply_small_rhs[s] ? p->steps[s].rhs : p->steps[s].lhs; };
  auto build_chain = [&](int s0, SchedItem& it) -> bool {
    if (!gate_like(s0)) return false;
    const int big0 = big_of(s0);
    const std::vector<int>& lay0 = p->layout[big0];
    const int r = (int)lay0.size();
    for (int x : lay0)
      if (x < 0) return false;  // an input with sliced bits: leave it to the apply kernel
    // candidate run: follow the single consumer while it is a gate-like apply on the result
    std::vector<int> cand{s0};
    for (int cur = s0; (int)cand.size() < 2048;) {
      const int t = n_in + cur;
      if (n_consumers[t] != 1) break;
      const int c = consumer_of[t];
      if (c < 0 || !gate_like(c) || big_of(c) != t || p->phase[c] != p->phase[s0] ||
          p->dep_batch[c] != p->dep_batch[s0] || p->dep_slice[c] != p->dep_slice[s0])
        break;
      const int sm = small_of(c);
      if (!done[sm] || (sm < n_in && in_slice_bits[sm] > 4)) break;
      cand.push_back(c);
      cur = c;
    }
    {
      const int sm0 = small_of(s0);
      if (sm0 < n_in && in_slice_bits[sm0] > 4) return false;
    }
    if (cand.size() < 2) return false;
    const int Lmax = p->dtype == TQ_C64 ? CHAIN_MAX_LOCAL : CHAIN_MAX_LOCAL - 1;
    const int L = std::min(Lmax, r);
    const int gate_cap = (32 * 1024) / (p->dtype == TQ_C64 ? 8 : 16);
    // local set: the indices the run contracts (as long as they fit), then the lowest bits of the layout
    std::set<int> original(lay0.begin(), lay0.end()), local;
    for (int b = 0; b < std::min(3, r) && (int)local.size() < L; ++b) local.insert(lay0[b]);
    int n_take = 0, gate_total = 0;
    {
      int big = big0;
      std::set<int> created;
      for (int s : cand) {
        const tq_tn_step& st = p->steps[s];
        const ApplyDev& a = p->apply[s];
        const int8_t* bb = p->apply_small_rhs[s] ? st.lhs_bits : st.rhs_bits;
        std::vector<int> need;
        for (int j = 0; j < a.n_k; ++j) {
          const int x = p->layout[big][bb[j]];
          if (original.count(x) && !created.count(x) && !local.count(x) &&
              std::find(need.begin(), need.end(), x) == need.end())
            need.push_back(x);
        }
        const int entries = 1 << (2 * a.n_k + a.n_b);
        if ((int)(local.size() + need.size()) > L || gate_total + entries > gate_cap) break;
        for (int x : need) local.insert(x);
        const int8_t* sb = p->apply_small_rhs[s] ? st.rhs_bits : st.lhs_bits;
        const int sm = small_of(s);
        for (int j = 0; j < a.n_k; ++j) created.insert(p->layout[sm][sb[a.n_k + j]]);
        gate_total += entries;
        big = n_in + s;
        ++n_take;
      }
    }
    if (n_take < 2) return false;
    for (int b = 0; b < r && (int)local.size() < L; ++b) local.insert(lay0[b]);
    ChainRun run;
    memset(&run.hdr, 0, sizeof(run.hdr));
    run.hdr.L = L;
    run.hdr.n_glob = r - L;
    run.hdr.n_steps = n_take;
    run.hdr.gate_entries = gate_total;
    run.step_begin = (int)chain_steps.size();
    run.big_in = big0;
    run.rank = r;
    std::map<int, int> pos, glob;        // index id -> tile bit / tile-number bit
    std::vector<int> at(L, -1);          // tile bit -> index id it holds now
    for (int b = 0, jl = 0, jg = 0; b < r; ++b) {
      const int x = lay0[b];
      if (local.count(x)) {
        pos[x] = jl;
        at[jl] = x;
        run.hdr.in_loc[jl++] = (int8_t)b;
      } else {
        glob[x] = jg;
        run.hdr.in_glob[jg++] = (int8_t)b;
      }
    }
    int big = big0, g_begin = 0;
    for (int i = 0; i < n_take; ++i) {
      const int s = cand[i];
      const tq_tn_step& st = p->steps[s];
      const ApplyDev& a = p->apply[s];
      const bool srhs = p->apply_small_rhs[s] != 0;
      const int8_t* bb = srhs ? st.lhs_bits : st.rhs_bits;   // large operand: [k..., free..., b...]
      const int8_t* sb = srhs ? st.rhs_bits : st.lhs_bits;   // small operand: [k..., new..., b...]
      const int sm = small_of(s);
      ChainStep c;
      memset(&c, 0, sizeof(c));
      c.n_k = (int8_t)a.n_k;
      c.n_b = (int8_t)a.n_b;
      c.g_begin = g_begin;
      g_begin += 1 << (2 * a.n_k + a.n_b);
      c.in = sm < n_in ? sm : (p->arena_const[sm - n_in] ? -1 : -2);
      c.off = -1;  // arena offsets are filled in once the layout is known (below: small_tensor)
      for (int j = 0; j < a.n_k; ++j) {
        const int x = p->layout[big][bb[j]];
        const int tp = pos.at(x);
        c.tpos[j] = (int8_t)tp;
        c.gk[j] = sb[j];
        c.gs[j] = sb[a.n_k + j];
        const int nx = p->layout[sm][sb[a.n_k + j]];
        pos.erase(x);
        pos[nx] = tp;
        at[tp] = nx;
      }
      for (int j = 0; j < a.n_k; ++j) c.tsort[j] = c.tpos[j];
      std::sort(c.tsort, c.tsort + a.n_k);
      const int n_f = a.n_f;
      for (int j = 0; j < a.n_b; ++j) {
        const int x = p->layout[big][bb[a.n_k + n_f + j]];
        c.bpos[j] = pos.count(x) ? (int8_t)pos[x] : (int8_t)(-1 - glob.at(x));
        c.gb[j] = sb[2 * a.n_k + j];
      }
      for (int j = 0; j < 4; ++j) c.sl_ord[j] = c.sl_bit[j] = -1;
      int ns = 0;
      for (size_t e = 0; e < p->slice_tensor.size(); ++e)
        if (p->slice_tensor[e] == sm && ns < 4) {
          c.sl_ord[ns] = (int8_t)p->slice_ord[e];
          c.sl_bit[ns++] = (int8_t)p->slice_bit[e];
        }
      chain_steps.push_back(c);
      chain_small.push_back(sm);
      it.members.push_back(s);
      big = n_in + s;
    }
    run.last = cand[n_take - 1];
    const std::vector<int>& layf = p->layout[n_in + run.last];
    std::map<int, int> fpos;
    for (size_t b = 0; b < layf.size(); ++b) fpos[layf[b]] = (int)b;
    for (int j = 0; j < L; ++j) run.hdr.out_loc[j] = (int8_t)fpos.at(at[j]);
    for (auto& kv : glob) run.hdr.out_glob[kv.second] = (int8_t)fpos.at(kv.first);
    it.chain = (int)p->chains.size();
    p->chains.push_back(run);
    return true;
  };
  for (int phase = 0; phase < 3; ++phase) {
    p->items[phase].clear();
    std::vector<int> remaining;
    for (int s = 0; s < n_steps; ++s)
      if (p->phase[s] == phase) remaining.push_back(s);
    while (!remaining.empty()) {
      bool progressed = false;
      for (int batched = 0; batched < 2; ++batched) {
        std::vector<int> run;
        std::vector<int> level(n_in + n_steps, 0);
        std::vector<char> in_run(n_in + n_steps, 0);
        for (int s : remaining) {
          const tq_tn_step& st = p->steps[s];
          if (p->kind[s] != 4 || (p->dep_batch[s] != 0) != (batched == 1)) continue;
          if (!(done[st.lhs] || in_run[st.lhs]) || !(done[st.rhs] || in_run[st.rhs])) continue;
          run.push_back(s);
          in_run[n_in + s] = 1;
          level[n_in + s] = 1 + std::max(in_run[st.lhs] ? level[st.lhs] : 0, in_run[st.rhs] ? level[st.rhs] : 0);
        }
        if (run.empty()) continue;
        progressed = true;
        auto is_cta = [&](int s) {
          const tq_tn_step& st = p->steps[s];
          return st.n_k + st.n_m + st.n_n + st.n_b >= FUSE_CTA_WORK;
        };
        std::stable_sort(run.begin(), run.end(), [&](int x, int y) {
          if (level[n_in + x] != level[n_in + y]) return level[n_in + x] < level[n_in + y];
          return is_cta(x) > is_cta(y);
        });
        SchedItem it;
        it.batched = batched == 1;
        it.members = run;
        it.fs_begin = (int)fsteps.size();
        it.n_fsteps = (int)run.size();
        it.n_levels = level[n_in + run.back()];
        it.lv_begin = (int)levels.size();
        std::vector<int32_t> off(it.n_levels + 1, 0), ncta(it.n_levels, 0);
        for (int s : run) {
          off[level[n_in + s]] += 1;  // counts, shifted by one level
          if (is_cta(s)) ncta[level[n_in + s] - 1] += 1;
        }
        for (int L = 0; L < it.n_levels; ++L) off[L + 1] += off[L];
        levels.insert(levels.end(), off.begin(), off.end());
        levels.insert(levels.end(), ncta.begin(), ncta.end());
        fsteps.resize(fsteps.size() + run.size());  // filled once the arena offsets are known
        for (int s : run) done[n_in + s] = 1;
        std::vector<int> rest;
        for (int s : remaining)
          if (!in_run[n_in + s]) rest.push_back(s);
        remaining.swap(rest);
        p->items[phase].push_back(std::move(it));
      }
      std::vector<int> rest;
      for (int s : remaining) {
        const tq_tn_step& st = p->steps[s];
        if (done[n_in + s]) continue;  // taken by an apply-chain run earlier in this pass
        if (p->kind[s] != 4 && done[st.lhs] && done[st.rhs]) {
          SchedItem it;
          it.step = s;
          it.batched = p->dep_batch[s] != 0;
          if (p->kind[s] == 5 && p->chain_enabled && build_chain(s, it)) {
            for (int m : it.members) {
              done[n_in + m] = 1;
              p->kind[m] = 7;
            }
            it.step = -1;
          } else {
            done[n_in + s] = 1;
          }
          p->items[phase].push_back(it);
          progressed = true;
        } else {
          rest.push_back(s);
        }
      }
      remaining.swap(rest);
      TQ_REQUIRE(progressed, TQ_E_INVALID, "tq_tn_plan: the contraction path is not a valid ssa order");
    }
  }
  // ---- arena layout: best-fit over a free list, simulated in execution order.  Outputs of once-per-call steps
  // that feed per-slice steps stay pinned for the whole slice loop; nothing is recycled inside a fused run.
  std::vector<int> item_of(n_steps, 0);
  std::vector<const SchedItem*> order;
  for (int phase = 0; phase < 3; ++phase)
    for (const SchedItem& it : p->items[phase]) {
      const int pos = (int)order.size();
      order.push_back(&it);
      if (it.step >= 0) item_of[it.step] = pos;
      for (int s : it.members) item_of[s] = pos;
    }
  std::vector<int> last_pos(n_in + n_steps, -1);
  for (int s = 0; s < n_steps; ++s) {
    last_pos[p->steps[s].lhs] = std::max(last_pos[p->steps[s].lhs], item_of[s]);
    last_pos[p->steps[s].rhs] = std::max(last_pos[p->steps[s].rhs], item_of[s]);
  }
  p->arena_off.assign(n_steps, 0);
  for (int arena = 0; arena < 2; ++arena) {  // 0: per-set, 1: shared
    struct Blk {
      int64_t off, size;
    };
    std::vector<Blk> freel;
    int64_t top = 0;
    std::vector<std::pair<int, Blk>> live;  // (tensor id, block)
    for (size_t pos = 0; pos < order.size(); ++pos) {
      const SchedItem& it = *order[pos];
      std::vector<int> members = it.members;
      if (it.step >= 0) members.push_back(it.step);
      for (int s : members) {
        if ((int)p->arena_const[s] != arena) continue;
        if (it.chain >= 0 && s != p->chains[it.chain].last) continue;  // never materialised: lives in shared memory
        int64_t size = ((int64_t)1 << p->t_rank[n_in + s]);
        size = std::max(size, p->tc[s].out_entries);  // a fused-pack result lives as its consumer's operand image
        size = (size + 15) & ~(int64_t)15;
        int best = -1;
        for (size_t j = 0; j < freel.size(); ++j)
          if (freel[j].size >= size && (best < 0 || freel[j].size < freel[best].size)) best = (int)j;
        Blk b;
        if (best >= 0) {
          b.off = freel[best].off;
          b.size = size;
          if (freel[best].size > size) {
            freel[best].off += size;
            freel[best].size -= size;
          } else {
            freel.erase(freel.begin() + best);
          }
        } else {
          b.off = top;
          b.size = size;
          top += size;
        }
        p->arena_off[s] = b.off;
        live.push_back({n_in + s, b});
      }
      for (int s : members) {
        const tq_tn_step& st = p->steps[s];
        for (int opnd : {st.lhs, st.rhs}) {
          if (opnd < n_in || last_pos[opnd] != (int)pos) continue;
          const int ps = opnd - n_in;
          if ((int)p->arena_const[ps] != arena) continue;
          if (!p->dep_slice[ps] && p->dep_slice[s]) continue;  // pinned across the slice loop
          for (size_t j = 0; j < live.size(); ++j)
            if (live[j].first == opnd) {
              freel.push_back(live[j].second);
              live.erase(live.begin() + j);
              break;
            }
        }
      }
    }
    (arena ? p->arena_shared : p->arena_set) = top;
  }
  // ---- device tables of the fused runs
  for (int phase = 0; phase < 3; ++phase)
    for (const SchedItem& it : p->items[phase]) {
      if (it.chain >= 0) continue;  // an apply-chain run has its own step table
      for (size_t i = 0; i < it.members.size(); ++i) {
        const int s = it.members[i];
        const tq_tn_step& st = p->steps[s];
        FusedStep& f = fsteps[it.fs_begin + i];
        memset(&f, 0, sizeof(f));
        auto space = [&](int t) { return t < n_in ? t : (p->arena_const[t - n_in] ? -1 : -2); };
        f.a_in = space(st.lhs);
        f.b_in = space(st.rhs);
        f.a_off = st.lhs < n_in ? 0 : p->arena_off[st.lhs - n_in];
        f.b_off = st.rhs < n_in ? 0 : p->arena_off[st.rhs - n_in];
        f.c_space = p->arena_const[s] ? -1 : -2;
        f.c_off = p->arena_off[s];
        f.n_k = (int8_t)st.n_k;
        f.n_m = (int8_t)st.n_m;
        f.n_n = (int8_t)st.n_n;
        f.n_b = (int8_t)st.n_b;
        const int outl = st.n_m + st.n_n + st.n_b;
        f.is_cta = st.n_k + outl >= FUSE_CTA_WORK;
        {  // micro step: everything a lane needs is tabulated here
          const int lpo_log2 = std::max(0, std::min({(int)st.n_k, 5, 5 - outl}));
          const int items = 1 << (outl + lpo_log2), iters = (1 << st.n_k) >> lpo_log2;
          if (!f.is_cta && outl + lpo_log2 <= 5 && iters <= 8) {
            f.micro_iters = (int8_t)iters;
            f.micro_lpo = (int8_t)lpo_log2;
            f.micro_row = (int32_t)(microtab.size() / 32);
            auto sc = [](uint32_t v, const int8_t* pos, int n) {
              uint32_t r = 0;
              for (int j = 0; j < n; ++j) r |= ((v >> j) & 1u) << pos[j];
              return r;
            };
            for (int j = 0; j < iters; ++j)
              for (int l = 0; l < 32; ++l) {
                if (l >= items) {
                  microtab.push_back(0xffffffffu);
                  continue;
                }
                const uint32_t o = (uint32_t)l >> lpo_log2, k = ((uint32_t)l & ((1u << lpo_log2) - 1u)) + ((uint32_t)j << lpo_log2);
                const uint32_t n = o & ((1u << st.n_n) - 1u), m = (o >> st.n_n) & ((1u << st.n_m) - 1u);
                const uint32_t bb = o >> (st.n_n + st.n_m);
                const uint32_t ao = sc(k, st.lhs_bits, st.n_k) | sc(m, st.lhs_bits + st.n_k, st.n_m) |
                                    sc(bb, st.lhs_bits + st.n_k + st.n_m, st.n_b);
                const uint32_t bo = sc(k, st.rhs_bits, st.n_k) | sc(n, st.rhs_bits + st.n_k, st.n_n) |
                                    sc(bb, st.rhs_bits + st.n_k + st.n_n, st.n_b);
                microtab.push_back(ao | (bo << 16));
              }
          }
        }
        for (int j = 0; j < st.n_k + st.n_m + st.n_b; ++j) f.a_bits[j] = st.lhs_bits[j];
        for (int j = 0; j < st.n_k + st.n_n + st.n_b; ++j) f.b_bits[j] = st.rhs_bits[j];
        for (int j = 0; j < FUSE_MAX_SLICE_BITS; ++j) f.a_sl_ord[j] = f.b_sl_ord[j] = -1;
        int na = 0, nb = 0;
        for (size_t e = 0; e < p->slice_tensor.size(); ++e) {
          if (p->slice_tensor[e] == st.lhs) {
            f.a_sl_ord[na] = (int8_t)p->slice_ord[e];
            f.a_sl_bit[na++] = (int8_t)p->slice_bit[e];
          }
          if (p->slice_tensor[e] == st.rhs) {
            f.b_sl_ord[nb] = (int8_t)p->slice_ord[e];
            f.b_sl_bit[nb++] = (int8_t)p->slice_bit[e];
          }
        }
      }
    }
  for (size_t i = 0; i < chain_steps.size(); ++i) {
    const int sm = chain_small[i];
    chain_steps[i].off = sm < n_in ? 0 : p->arena_off[sm - n_in];
  }
  cudaFree(p->d_chain_steps);
  p->d_chain_steps = nullptr;
  if (!chain_steps.empty()) {
    TQ_CUDA_OK(cudaMalloc((void**)&p->d_chain_steps, chain_steps.size() * sizeof(ChainStep)));
    TQ_CUDA_OK(cudaMemcpy(p->d_chain_steps, chain_steps.data(), chain_steps.size() * sizeof(ChainStep),
                          cudaMemcpyHostToDevice));
  }
  cudaFree(p->d_fsteps);
  cudaFree(p->d_levels);
  cudaFree(p->d_micro);
  p->d_fsteps = nullptr;
  p->d_levels = nullptr;
  p->d_micro = nullptr;
  if (!microtab.empty()) {
    TQ_CUDA_OK(cudaMalloc((void**)&p->d_micro, microtab.size() * sizeof(uint32_t)));
    TQ_CUDA_OK(cudaMemcpy(p->d_micro, microtab.data(), microtab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  if (!fsteps.empty()) {
    TQ_CUDA_OK(cudaMalloc((void**)&p->d_fsteps, fsteps.size() * sizeof(FusedStep)));
    TQ_CUDA_OK(cudaMemcpy(p->d_fsteps, fsteps.data(), fsteps.size() * sizeof(FusedStep), cudaMemcpyHostToDevice));
    TQ_CUDA_OK(cudaMalloc((void**)&p->d_levels, levels.size() * sizeof(int32_t)));
    TQ_CUDA_OK(cudaMemcpy(p->d_levels, levels.data(), levels.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  return TQ_OK;
}

extern "C" {

int32_t tq_tn_symbol(int32_t i) {
  static const char base[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ";
  return i < 52 ? (int32_t)base[i] : i + 140;
}

int32_t tq_tn_index_map(int32_t n_qubits, const int32_t* gate_nq, const int32_t* gate_qubits, int32_t n_gates,
                        int32_t meas_kind, const int32_t* obs_nq, const int32_t* obs_qubits, int32_t n_obs,
                        const int32_t* kept_qubits, int32_t n_kept, int32_t* tensor_off, int32_t* tensor_idx,
                        int32_t tensor_cap, int32_t idx_cap, int32_t* out_idx, int32_t* n_out) {
  TQ_REQUIRE(n_qubits > 0 && n_gates >= 0 && tensor_off && tensor_idx && out_idx && n_out, TQ_E_INVALID,
             "tq_tn_index_map: bad arguments");
  std::vector<std::vector<int>> T;
  std::vector<int> wire(n_qubits);
  int cur = n_qubits - 1;
  for (int q = 0; q < n_qubits; ++q) {
    wire[q] = q;
    T.push_back({q});
  }
  std::vector<int> idx;
  for (int g = 0; g < n_gates; ++g) {
    thread_gate(wire, cur, gate_qubits + 4 * g, gate_nq[g], idx);
    T.push_back(idx);
  }
  std::vector<int> outv;
  if (meas_kind == TQ_M_STATE) {
    for (int q = 0; q < n_qubits; ++q) outv.push_back(wire[q]);
  } else {
    if (meas_kind == TQ_M_EXPVAL) {
      for (int j = 0; j < n_obs; ++j) {
        thread_gate(wire, cur, obs_qubits + 4 * j, obs_nq[j], idx);
        T.push_back(idx);
      }
    } else if (meas_kind == TQ_M_PROBS) {
      for (int j = 0; j < n_kept; ++j) outv.push_back(wire[kept_qubits[j]]);
    } else {
      TQ_REQUIRE(false, TQ_E_INVALID, "tq_tn_index_map: unknown measurement kind %d", meas_kind);
    }
    for (int g = n_gates - 1; g >= 0; --g) {
      TQ_REQUIRE(gate_nq[g] <= 3, TQ_E_UNSUPPORTED,
                 "Error!! unknown operator with len of applied qubits larger than 3!");
      thread_gate(wire, cur, gate_qubits + 4 * g, gate_nq[g], idx);
      T.push_back(idx);
    }
    for (int q = 0; q < n_qubits; ++q) T.push_back({wire[q]});
  }
  TQ_REQUIRE((int)T.size() <= tensor_cap, TQ_E_INVALID, "tq_tn_index_map: %zu tensors > capacity", T.size());
  int off = 0;
  for (size_t t = 0; t < T.size(); ++t) {
    tensor_off[t] = off;
    TQ_REQUIRE(off + (int)T[t].size() <= idx_cap, TQ_E_INVALID, "tq_tn_index_map: index capacity exceeded");
    for (int ix : T[t]) tensor_idx[off++] = ix;
  }
  tensor_off[T.size()] = off;
  *n_out = (int)outv.size();
  for (size_t i = 0; i < outv.size(); ++i) out_idx[i] = outv[i];
  return (int32_t)T.size();
}

int32_t tq_tn_lower(const int32_t* tensor_off, const int32_t* tensor_idx, int32_t n_in, const int32_t* out_idx,
                    int32_t n_out, const int32_t* ssa_path, int32_t n_steps, const int32_t* sliced,
                    int32_t n_sliced, tq_tn_step* steps, int32_t* slice_tensor, int32_t* slice_ord,
                    int32_t* slice_bit, int32_t slice_cap, int32_t* n_slice_entries, int32_t* final_perm) {
  TQ_REQUIRE(tensor_off && tensor_idx && ssa_path && steps && n_in > 0 && n_steps == n_in - 1, TQ_E_INVALID,
             "tq_tn_lower: bad arguments (a path over n inputs has n-1 steps)");
  LowerResult R;
  int rc = lower_impl(tensor_off, tensor_idx, n_in, out_idx, n_out, ssa_path, n_steps, sliced, n_sliced, R);
  if (rc) return rc;
  for (int s = 0; s < n_steps; ++s) steps[s] = R.steps[s];
  TQ_REQUIRE((int)R.slice_tensor.size() <= slice_cap, TQ_E_INVALID, "tq_tn_lower: slice capacity exceeded");
  for (size_t i = 0; i < R.slice_tensor.size(); ++i) {
    slice_tensor[i] = R.slice_tensor[i];
    slice_ord[i] = R.slice_ord[i];
    slice_bit[i] = R.slice_bit[i];
  }
  if (n_slice_entries) *n_slice_entries = (int)R.slice_tensor.size();
  if (final_perm)
    for (int j = 0; j < n_out; ++j) final_perm[j] = R.final_perm[j];
  return n_steps;
}

void tq_tn_plan_destroy(tq_tn_plan* p) {
  if (!p) return;
  for (int32_t* q : p->d_ka) cudaFree(q);
  for (int32_t* q : p->d_kb) cudaFree(q);
  cudaFree(p->d_fsteps);
  cudaFree(p->d_levels);
  cudaFree(p->d_micro);
  cudaFree(p->d_chain_steps);
  delete p;
}

int tq_tn_plan_create(const int32_t* tensor_off, const int32_t* tensor_idx, int32_t n_in, const int32_t* out_idx,
                      int32_t n_out, const int32_t* ssa_path, int32_t n_steps, const int32_t* sliced,
                      int32_t n_sliced, const int32_t* input_batched, int32_t dtype, tq_tn_plan** out) {
  TQ_REQUIRE(out, TQ_E_INVALID, "tq_tn_plan_create: out is null");
  *out = nullptr;
  TQ_REQUIRE(dtype == TQ_C64 || dtype == TQ_C128, TQ_E_INVALID, "tq_tn_plan_create: bad dtype");
  TQ_REQUIRE(n_in > 0 && n_steps == n_in - 1 && n_out <= TQ_TN_MAX_RANK && n_sliced >= 0 && n_sliced < 40,
             TQ_E_INVALID, "tq_tn_plan_create: bad sizes");
  std::unique_ptr<tq_tn_plan, void (*)(tq_tn_plan*)> P(new tq_tn_plan(), tq_tn_plan_destroy);
  tq_tn_plan* p = P.get();
  p->dtype = dtype;
  p->n_in = n_in;
  p->n_out = n_out;
  p->n_sliced = n_sliced;
  p->device = tq::current_device();
  LowerResult R;
  int rc = lower_impl(tensor_off, tensor_idx, n_in, out_idx, n_out, ssa_path, n_steps, sliced, n_sliced, R);
  if (rc) return rc;
  p->steps = R.steps;
  p->slice_tensor = R.slice_tensor;
  p->slice_ord = R.slice_ord;
  p->slice_bit = R.slice_bit;
  p->final_perm = R.final_perm;
  p->in_rank.resize(n_in);
  p->in_batched.resize(n_in);
  p->t_batch.assign(n_in, 0);
  p->t_slice.assign(n_in, 0);
  p->t_rank.assign(n_in, 0);
  p->layout.assign(n_in, {});
  for (int t = 0; t < n_in; ++t) {
    p->in_rank[t] = tensor_off[t + 1] - tensor_off[t];
    p->in_batched[t] = input_batched && input_batched[t];
    p->t_batch[t] = p->in_batched[t];
    p->t_rank[t] = p->in_rank[t];
    p->width = std::max(p->width, p->in_rank[t]);
    for (int pos = p->in_rank[t] - 1; pos >= 0; --pos) {  // a sliced index keeps its physical bit: placeholder -1
      const int ix = tensor_idx[tensor_off[t] + pos];
      bool is_sliced = false;
      for (int j = 0; j < n_sliced; ++j) is_sliced |= sliced[j] == ix;
      p->layout[t].push_back(is_sliced ? -1 : ix);
    }
  }
  for (int t : p->slice_tensor) p->t_slice[t] = 1;
  p->n_fwd = n_steps;
  p->phase.assign(n_steps, 0);
  {
    int dev_id = 0, sms = 0;
    if (cudaGetDevice(&dev_id) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id) == cudaSuccess && sms > 0)
      p->num_sms = sms;
    (void)cudaGetLastError();
  }
  rc = setup_steps(p, 0);
  if (rc) return rc;
  for (int s = 0; s < n_steps; ++s) {
    p->phase[s] = p->dep_slice[s] ? 1 : 0;
    const tq_tn_step& st = p->steps[s];
    p->flops += 8.0 * (double)((int64_t)1 << (st.n_k + st.n_m + st.n_n + st.n_b));
  }
  rc = build_schedule(p);
  if (rc) return rc;
  *out = P.release();
  return TQ_OK;
}

// one pairwise step lhs x rhs -> tensor with index set `target` (every shared index outside target is summed)
static int make_grad_step(tq_tn_plan* p, int lhs, int rhs, const std::vector<int>& target, tq_tn_step& st,
                          std::vector<int>& out_layout) {
  const std::vector<int>&L = p->layout[lhs], &Rr = p->layout[rhs];
  std::map<int, int> posR;
  for (size_t i = 0; i < Rr.size(); ++i)
    if (Rr[i] >= 0) posR[Rr[i]] = (int)i;
  std::set<int> inL(L.begin(), L.end()), tgt(target.begin(), target.end());
  tgt.erase(-1);  // sliced indices of an input are fixed per slice: not part of its gradient tensor
  memset(&st, 0, sizeof(st));
  st.lhs = lhs;
  st.rhs = rhs;
  std::vector<int> K, M, Bt, N;  // bit positions in L (K, M, Bt) / in R (N)
  for (size_t i = 0; i < L.size(); ++i) {
    if (L[i] < 0) continue;
    if (posR.count(L[i])) (tgt.count(L[i]) ? Bt : K).push_back((int)i);
    else M.push_back((int)i);
  }
  for (size_t i = 0; i < Rr.size(); ++i)
    if (Rr[i] >= 0 && !inL.count(Rr[i])) N.push_back((int)i);
  st.n_k = (int)K.size();
  st.n_m = (int)M.size();
  st.n_n = (int)N.size();
  st.n_b = (int)Bt.size();
  TQ_REQUIRE(st.n_k + st.n_m + st.n_b <= TQ_TN_MAX_RANK && st.n_k + st.n_n + st.n_b <= TQ_TN_MAX_RANK &&
                 st.n_m + st.n_n + st.n_b <= TQ_TN_MAX_RANK,
             TQ_E_UNSUPPORTED, "tq_tn_plan_enable_backward: a gradient step exceeds rank %d", TQ_TN_MAX_RANK);
  TQ_REQUIRE((size_t)(st.n_m + st.n_n + st.n_b) == tgt.size(), TQ_E_INVALID,
             "tq_tn_plan_enable_backward: gradient step does not produce the operand's index set");
  int w = 0;
  for (int i : K) st.lhs_bits[w++] = (int8_t)i;
  for (int i : M) st.lhs_bits[w++] = (int8_t)i;
  for (int i : Bt) st.lhs_bits[w++] = (int8_t)i;
  w = 0;
  for (int i : K) st.rhs_bits[w++] = (int8_t)posR[L[i]];
  for (int i : N) st.rhs_bits[w++] = (int8_t)i;
  for (int i : Bt) st.rhs_bits[w++] = (int8_t)posR[L[i]];
  out_layout.clear();
  for (int i : N) out_layout.push_back(Rr[i]);
  for (int i : M) out_layout.push_back(L[i]);
  for (int i : Bt) out_layout.push_back(L[i]);
  for (size_t j = 0; j < out_layout.size(); ++j) {
    TQ_REQUIRE(tgt.count(out_layout[j]), TQ_E_INVALID, "tq_tn_plan_enable_backward: stray index in a gradient step");
    st.out_idx[j] = out_layout[j];
  }
  return TQ_OK;
}

int tq_tn_plan_enable_backward(tq_tn_plan* p, const int32_t* input_needs_grad) {
  TQ_REQUIRE(p && input_needs_grad, TQ_E_INVALID, "tq_tn_plan_enable_backward: null argument");
  TQ_REQUIRE(p->seed_step < 0, TQ_E_INVALID, "tq_tn_plan_enable_backward: already enabled");
  const int n_in = p->n_in, nf = p->n_fwd;
  std::vector<char> needs(n_in + nf, 0);
  for (int t = 0; t < n_in; ++t) needs[t] = input_needs_grad[t] != 0;
  for (int s = 0; s < nf; ++s) needs[n_in + s] = needs[p->steps[s].lhs] || needs[p->steps[s].rhs];
  const int final_t = n_in + nf - 1;
  TQ_REQUIRE(needs[final_t], TQ_E_INVALID, "tq_tn_plan_enable_backward: no input needs a gradient");
  p->grad_of.assign(n_in + nf, -1);
  // seed: conj(grad_out) in the layout of the last forward tensor
  tq_tn_step seed;
  memset(&seed, 0, sizeof(seed));
  seed.lhs = seed.rhs = final_t;
  for (size_t j = 0; j < p->layout[final_t].size(); ++j) seed.out_idx[j] = p->layout[final_t][j];
  p->seed_step = (int)p->steps.size();
  p->steps.push_back(seed);
  p->layout.push_back(p->layout[final_t]);
  p->grad_of[final_t] = n_in + p->seed_step;
  int rc;
  for (int s = nf - 1; s >= 0; --s) {
    const int C = n_in + s;
    if (!needs[C]) continue;
    const int gC = p->grad_of[C];
    const int A = p->steps[s].lhs, B = p->steps[s].rhs;
    for (int side = 0; side < 2; ++side) {
      const int T = side == 0 ? A : B, other = side == 0 ? B : A;
      if (!needs[T]) continue;
      tq_tn_step st;
      std::vector<int> lay;
      // g_A = g_C x B summed over C's indices that A lacks;  g_B = A x g_C likewise (conjugated gradients:
      // no operand needs conjugating)
      if ((rc = side == 0 ? make_grad_step(p, gC, other, p->layout[T], st, lay)
                          : make_grad_step(p, other, gC, p->layout[T], st, lay)))
        return rc;
      p->grad_of[T] = n_in + (int)p->steps.size();
      p->steps.push_back(st);
      p->layout.push_back(lay);
    }
  }
  p->grad_of.resize(n_in + p->steps.size(), -1);
  p->phase.resize(p->steps.size(), 2);
  if ((rc = setup_steps(p, nf))) return rc;
  return build_schedule(p);
}

/* Where the (conjugated) gradient of input t lives after tq_tn_backward: element offset inside its arena,
 * space (-1 shared arena, -2 per-set arena) and, for the i-th index of the input's own index list
 * (slow -> fast, as given to tq_tn_plan_create), its bit position inside the gradient tensor. */
int tq_tn_grad_info(const tq_tn_plan* p, int32_t t, int64_t* offset, int32_t* space, int32_t* bits) {
  TQ_REQUIRE(p && offset && space && bits && t >= 0 && t < p->n_in, TQ_E_INVALID, "tq_tn_grad_info: bad argument");
  TQ_REQUIRE(p->seed_step >= 0 && p->grad_of[t] >= 0, TQ_E_INVALID, "tq_tn_grad_info: input %d has no gradient", t);
  const int g = p->grad_of[t], sidx = g - p->n_in;
  *offset = p->arena_off[sidx];
  *space = p->arena_const[sidx] ? -1 : -2;
  const std::vector<int>& lay = p->layout[g];
  const std::vector<int>& mine = p->layout[t];  // fast -> slow
  const int r = (int)mine.size();
  for (int i = 0; i < r; ++i) {
    const int ix = mine[r - 1 - i];
    bits[i] = -1;  // stays -1 for a sliced index (fixed per slice, absent from the gradient tensor)
    if (ix < 0) continue;
    for (size_t j = 0; j < lay.size(); ++j)
      if (lay[j] == ix) bits[i] = (int32_t)j;
    TQ_REQUIRE(bits[i] >= 0, TQ_E_INVALID, "tq_tn_grad_info: index missing from the gradient tensor");
  }
  return TQ_OK;
}

/* byte offsets of the two arenas inside an (aligned) workspace and the per-set stride in complex entries */
int tq_tn_workspace_layout(const tq_tn_plan* p, int64_t* shared_off, int64_t* perset_off, int64_t* set_stride) {
  TQ_REQUIRE(p && shared_off && perset_off && set_stride, TQ_E_INVALID, "tq_tn_workspace_layout: null argument");
  const int64_t cs = p->dtype == TQ_C64 ? 8 : 16;
  *shared_off = (int64_t)(((size_t)p->n_in * sizeof(InputRef) + 255) & ~(size_t)255);
  *perset_off = *shared_off + p->arena_shared * cs;
  *set_stride = p->arena_set;
  return TQ_OK;
}

int tq_tn_plan_set_option(tq_tn_plan* p, int32_t option, int32_t value) {
  TQ_REQUIRE(p, TQ_E_INVALID, "tq_tn_plan_set_option: null plan");
  switch (option) {
    case TQ_TN_OPT_TENSOR_CORE:
      p->tc_enabled = value != 0;
      return build_schedule(p);
    case TQ_TN_OPT_TC_MIN_LOG2:
      p->tc_min_log2 = value;
      return build_schedule(p);
    case TQ_TN_OPT_TC_SPLITK:
      p->tc_splitk = value != 0;
      return TQ_OK;
    case TQ_TN_OPT_TC_GATHER:
      p->tc_gather = value != 0;
      return TQ_OK;
    case TQ_TN_OPT_TC_FUSE_PACK:
      p->tc_fuse_pack = value != 0;
      return build_schedule(p);
    case TQ_TN_OPT_CHAIN:
      p->chain_enabled = value != 0;
      return build_schedule(p);
    case TQ_TN_OPT_FUSE_SMALL:
      p->fuse_enabled = value != 0;
      return build_schedule(p);
    case TQ_TN_OPT_TC_CHUNK:
      TQ_REQUIRE(value >= 8, TQ_E_INVALID, "tq_tn_plan_set_option: chunk must be >= 8 complex k");
      p->tc_chunk = value;
      return TQ_OK;
    default:
      TQ_REQUIRE(false, TQ_E_INVALID, "tq_tn_plan_set_option: unknown option %d", option);
  }
}

/* 0: one thread per output element, 1: tiled FMA GEMM, 2: tcgen05 split-TF32 GEMM, 3: split-K reduction,
 * 4: member of a fused run of small steps */
int32_t tq_tn_plan_step_kernel(const tq_tn_plan* p, int32_t s) {
  if (!p || s < 0 || s >= (int)p->steps.size()) return TQ_E_INVALID;
  return p->kind[s];
}

/* step whose operand image step s writes from its epilogue (fused pack), -1: s stores its result plain */
int32_t tq_tn_plan_step_fuse_to(const tq_tn_plan* p, int32_t s) {
  if (!p || s < 0 || s >= (int)p->steps.size() || !p->tc_fuse_pack || p->kind[s] != 2) return -1;
  return p->tc[s].fuse_to;
}

/* store mode of a fused-pack producer (0: none): 1 = the consumer's low k bits are accumulator row bits, 2 = lowest k
 * bit a column bit + two row bits, 3 = three column bits */
int32_t tq_tn_plan_step_fuse_mode(const tq_tn_plan* p, int32_t s) {
  if (tq_tn_plan_step_fuse_to(p, s) < 0) return 0;
  return p->tc[s].fuse_mode;
}

int32_t tq_tn_plan_num_steps(const tq_tn_plan* p) { return p ? (int32_t)p->steps.size() : -1; }
int64_t tq_tn_plan_num_slices(const tq_tn_plan* p) { return p ? ((int64_t)1 << p->n_sliced) : -1; }
double tq_tn_plan_flops(const tq_tn_plan* p) { return p ? p->flops : -1.0; }
int32_t tq_tn_plan_width(const tq_tn_plan* p) { return p ? p->width : -1; }
int32_t tq_tn_plan_get_step(const tq_tn_plan* p, int32_t s, tq_tn_step* out) {
  if (!p || !out || s < 0 || s >= (int)p->steps.size()) return TQ_E_INVALID;
  *out = p->steps[s];
  return TQ_OK;
}

/* bit 0: the step repeats for every slice, bit 1: it carries the parameter-set batch dimension */
int32_t tq_tn_plan_step_flags(const tq_tn_plan* p, int32_t s) {
  if (!p || s < 0 || s >= (int)p->steps.size()) return TQ_E_INVALID;
  return (p->dep_slice[s] ? 1 : 0) | (p->dep_batch[s] ? 2 : 0);
}

// split-K factor of tensor-core step s at a given batch: power of two, >= 16 k-blocks (128 complex k) per split,
// only when the step has tiles for at most half of the SMs
static int tc_splits(const tq_tn_plan* p, int s, int64_t sets) {
  const TcStep& T = p->tc[s];
  const int64_t nz = sets << p->steps[s].n_b;
  const int64_t tiles = (int64_t)T.tiles_a * T.tiles_b * nz;
  int splits = 1;
  while (p->tc_splitk && tiles * splits * 2 <= p->num_sms && T.kblocks / (splits * 2) >= 16) splits *= 2;
  return splits;
}

// pinned images (slice-invariant operands of per-slice tensor-core steps): total bytes; offsets[2*s], [2*s+1]
static size_t tn_pinned_bytes(const tq_tn_plan* p, int64_t batch, std::vector<int64_t>* offsets) {
  size_t top = 0;
  if (offsets) offsets->assign(2 * p->steps.size(), -1);
  for (size_t s = 0; s < p->steps.size(); ++s) {
    if (tq_tn_plan_step_kernel(p, (int32_t)s) != 2) continue;
    const int64_t nz = (p->dep_batch[s] ? batch : 1) << p->steps[s].n_b;
    const TcStep& T = p->tc[s];
    if (T.pin_a) {
      if (offsets) (*offsets)[2 * s] = (int64_t)top;
      top += (size_t)(T.img_a_z * nz);
    }
    if (T.pin_b) {
      if (offsets) (*offsets)[2 * s + 1] = (int64_t)top;
      top += (size_t)(T.img_b_z * nz);
    }
  }
  return top;
}

static size_t tn_image_bytes(const tq_tn_plan* p, int64_t batch) {
  size_t need = 0;
  for (size_t s = 0; s < p->steps.size(); ++s) {
    const int kernel = tq_tn_plan_step_kernel(p, (int32_t)s);
    if (kernel == 3) {  // partial sums of the split-K reduction
      const tq_tn_step& st = p->steps[s];
      const size_t cs = p->dtype == TQ_C64 ? 8 : 16;
      need = std::max(need, (size_t)(p->dep_batch[s] ? batch : 1) * ((size_t)DOT_BLOCKS << (st.n_m + st.n_n + st.n_b)) * cs);
    }
    if (kernel != 2) continue;
    const int64_t nz = (p->dep_batch[s] ? batch : 1) << p->steps[s].n_b;
    const TcStep& T = p->tc[s];
    const int splits = tc_splits(p, (int)s, p->dep_batch[s] ? batch : 1);
    const size_t partials = splits > 1 ? (size_t)splits * (((size_t)(p->dep_batch[s] ? batch : 1)) << p->t_rank[p->n_in + s]) * 8 : 0;
    need = std::max(need, (size_t)(((T.pin_a ? 0 : T.img_a_z) + (T.pin_b ? 0 : T.img_b_z)) * nz) + partials + 1024);
  }
  return need + tn_pinned_bytes(p, batch, nullptr);
}

static size_t tn_table_bytes(const tq_tn_plan* p) {  // input pointer table read by the fused runs
  return ((size_t)p->n_in * sizeof(InputRef) + 255) & ~(size_t)255;
}

size_t tq_tn_workspace_bytes(const tq_tn_plan* p, int64_t batch) {
  if (!p || batch <= 0) return 0;
  const size_t cs = p->dtype == TQ_C64 ? 8 : 16;
  const size_t arenas = ((size_t)(p->arena_shared + p->arena_set * batch) * cs + 1023) & ~(size_t)1023;
  return tn_table_bytes(p) + arenas + tn_image_bytes(p, batch) + 1024;
}

}  // extern "C"

namespace tq {

// kernel launches enqueued by the tensor-network entry points since the library was loaded (bench.py's gpu_launches)
static std::atomic<int64_t> g_tn_launches{0};

static int tc_setup_once() {  // once per device: the opt-in shared-memory size is a per-device function attribute
  static bool done_dev[64] = {};
  const int dev = current_device();
  bool& done = done_dev[dev >= 0 && dev < 64 ? dev : 0];
  if (done) return TQ_OK;
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_gemm<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GEMM_SMEM));
  TQ_CUDA_OK(cudaFuncSetAttribute(tc::k_tc_pack<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  done = true;
  return TQ_OK;
}

// Does step s write its consumer's operand image instead of a plain tensor?  (From the GEMM epilogue, or — a split-K
// launch — from the kernel that adds the partial sums.)
static bool fused_out_active(const tq_tn_plan* p, int s, int64_t sets) {
  (void)sets;
  return s >= 0 && p->tc_fuse_pack && p->tc[s].fuse_to >= 0;
}

// Pack operands into images (mode bit 0), run one persistent tcgen05 GEMM over all (z, row tile, column tile)
// (mode bit 1).  img_a / img_b: where each operand's image lives (scratch or pinned); skip_a / skip_b: that image
// is pinned and already packed.
static int run_step_tc(const tq_tn_plan* p, int s, const cx<float>* a, int64_t sa, const cx<float>* b, int64_t sb,
                       cx<float>* c, int64_t sc, int64_t sets, uint8_t* img_a, uint8_t* img_b, uint8_t* partials,
                       bool pack_a, bool pack_b, bool gemm, cudaStream_t st, cudaEvent_t ev_packed) {
  int rc = tc_setup_once();
  if (rc) return rc;
  const tq_tn_step& stp = p->steps[s];
  const TcStep& T = p->tc[s];
  const int64_t nz = sets << stp.n_b;
  TQ_REQUIRE(nz < 65536, TQ_E_UNSUPPORTED, "tq_tn_contract: step %d has %lld batched GEMMs", s, (long long)nz);
  // operand images written by the producing step's epilogue (fused pack): nothing to pack, the image is the tensor
  const bool premade_a = fused_out_active(p, T.src_a, sets), premade_b = fused_out_active(p, T.src_b, sets);
  int64_t img_a_set = T.img_a_z << stp.n_b, img_b_set = T.img_b_z << stp.n_b;
  if (premade_a) {
    pack_a = false;
    img_a = (uint8_t*)const_cast<cx<float>*>(T.swap ? b : a);
    img_a_set = (T.swap ? sb : sa) * (int64_t)sizeof(cx<float>);
  }
  if (premade_b) {
    pack_b = false;
    img_b = (uint8_t*)const_cast<cx<float>*>(T.swap ? a : b);
    img_b_set = (T.swap ? sa : sb) * (int64_t)sizeof(cx<float>);
  }
  const bool gather_a = gemm && pack_a && T.gather_a && p->tc_gather;
  if (gather_a) pack_a = false;
  if (pack_a || pack_b) {
    tc::PackPair pp;
    memset(&pp, 0, sizeof(pp));
    size_t smem = 0;
    if (pack_a) {
      pp.a = T.pa;
      pp.a.src = reinterpret_cast<const float2*>(T.swap ? b : a);
      pp.a.src_set_stride = T.swap ? sb : sa;
      pp.a.img = img_a;
      pp.a.img_z_stride = T.img_a_z;
      pp.nch_a = tc::pack_nch((int64_t)T.tiles_a * T.kblocks * nz, p->num_sms);
      pp.blocks_a = tc::pack_blocks(T.tiles_a, T.kblocks, pp.nch_a);
      smem = 2 * (size_t)tc::A_CHUNK;
    }
    if (pack_b) {
      pp.b = T.pb;
      pp.b.src = reinterpret_cast<const float2*>(T.swap ? a : b);
      pp.b.src_set_stride = T.swap ? sa : sb;
      pp.b.img = img_b;
      pp.b.img_z_stride = T.img_b_z;
      pp.nch_b = tc::pack_nch((int64_t)T.tiles_b * T.kblocks * nz, p->num_sms);
      pp.blocks_b = tc::pack_blocks(T.tiles_b, T.kblocks, pp.nch_b);
      smem = std::max(smem, 2 * (size_t)tc::b_chunk_bytes(T.c_t));
    }
    tc::k_tc_pack<256><<<dim3((unsigned)(pp.blocks_a + pp.blocks_b), (unsigned)nz), 256, smem, st>>>(pp);
    g_tn_launches += 1;
  }
  TQ_CUDA_OK(cudaGetLastError());
  if (ev_packed) TQ_CUDA_OK(cudaEventRecord(ev_packed, st));
  if (!gemm) return TQ_OK;
  tc::GemmParams g;
  memset(&g, 0, sizeof(g));
  g.img_a = img_a;
  g.img_b = img_b;
  g.c = reinterpret_cast<float2*>(c);
  g.img_a_z = T.img_a_z;
  g.img_b_z = T.img_b_z;
  g.img_a_set = img_a_set;
  g.img_b_set = img_b_set;
  g.c_set_stride = sc;
  g.c_rs = T.swap ? 1 : ((int64_t)1 << stp.n_n);
  g.c_cs = T.swap ? ((int64_t)1 << stp.n_n) : 1;
  g.tiles_a = T.tiles_a;
  g.tiles_b = T.tiles_b;
  g.tb_fast = T.img_a_z >= T.img_b_z ? 1 : 0;  // keep the larger image's tile shared through L2
  g.kblocks = T.kblocks;
  g.chunk = std::max(1, p->tc_chunk / tc::KB_CPLX);  // in k-blocks
  {
    const char* dbg = getenv("TQ_TC_DEBUG");
    g.debug = dbg ? atoi(dbg) : 0;
  }
  g.n_z = (int32_t)nz;
  g.n_b_log2 = stp.n_b;
  g.stages = T.stages;
  g.c_bb_stride = (int64_t)1 << (stp.n_m + stp.n_n);
  const int splits = tc_splits(p, s, sets);
  const int64_t c_elems = (int64_t)1 << (stp.n_m + stp.n_n + stp.n_b);
  g.splits = splits;
  g.kb_per_split = T.kblocks / splits;
  const bool img_out = fused_out_active(p, s, sets);
  tc::ImgOut out_img = T.out;
  out_img.img = (uint8_t*)c;
  out_img.set_stride = sc * (int64_t)sizeof(cx<float>);
  if (img_out && splits == 1) g.out = out_img;  // the result goes straight into the consumer's operand image
  if (splits > 1) {  // partial sums: [split][set][C], summed in order afterwards
    g.c = reinterpret_cast<float2*>(partials);
    g.c_set_stride = c_elems;
    g.c_split_stride = sets * c_elems;
    if (img_out) {  // accumulator order [bb][row][col]: the rows / columns of a fused producer are re-ordered
      const int n_col = T.swap ? stp.n_m : stp.n_n;
      g.c_rs = (int64_t)1 << n_col;
      g.c_cs = 1;
    }
  }
  const int64_t total = (int64_t)T.tiles_a * T.tiles_b * nz * splits;
  const unsigned grid = (unsigned)std::min<int64_t>(total, p->num_sms);
  const size_t smem = tc::GEMM_SMEM;
  if (gather_a) {
    g.ga = T.pa;
    g.ga.src = reinterpret_cast<const float2*>(T.swap ? b : a);
    g.ga.src_set_stride = T.swap ? sb : sa;
    switch (T.c_t) {
      case 16: tc::k_tc_gemm<16, true><<<grid, tc::GEMM_THREADS_GA, smem, st>>>(g); break;
      case 32: tc::k_tc_gemm<32, true><<<grid, tc::GEMM_THREADS_GA, smem, st>>>(g); break;
      case 64: tc::k_tc_gemm<64, true><<<grid, tc::GEMM_THREADS_GA, smem, st>>>(g); break;
      default: tc::k_tc_gemm<128, true><<<grid, tc::GEMM_THREADS_GA, smem, st>>>(g); break;
    }
  } else {
    switch (T.c_t) {
      case 16: tc::k_tc_gemm<16, false><<<grid, tc::GEMM_THREADS, smem, st>>>(g); break;
      case 32: tc::k_tc_gemm<32, false><<<grid, tc::GEMM_THREADS, smem, st>>>(g); break;
      case 64: tc::k_tc_gemm<64, false><<<grid, tc::GEMM_THREADS, smem, st>>>(g); break;
      default: tc::k_tc_gemm<128, false><<<grid, tc::GEMM_THREADS, smem, st>>>(g); break;
    }
  }
  g_tn_launches += splits > 1 ? 2 : 1;
  if (splits > 1 && img_out) {
    tc::k_tc_splitk_sum_img<<<dim3((unsigned)((c_elems + 255) / 256), (unsigned)sets), 256, 0, st>>>(
        reinterpret_cast<const float2*>(partials), splits, sets * c_elems, c_elems, out_img, c_elems);
  } else if (splits > 1) {
    const int64_t n4 = c_elems / 2;
    tc::k_tc_splitk_sum<<<dim3((unsigned)((n4 + 255) / 256), (unsigned)sets), 256, 0, st>>>(
        reinterpret_cast<const float4*>(partials), splits, sets * n4, n4, reinterpret_cast<float4*>(c), sc / 2, n4);
  }
  TQ_CUDA_OK(cudaGetLastError());
  return TQ_OK;
}

// step_ms (optional, profiling): 2 floats per step = [whole step, of which operand packing] in milliseconds,
// for the slice-invariant steps and the steps of slice s_begin; forces a synchronisation.
template <typename R>
static int contract_impl(const tq_tn_plan* p, const void* const* inputs, const int64_t* strides, int64_t B,
                         int64_t s_begin, int64_t s_end, void* out, void* workspace, size_t ws_bytes,
                         cudaStream_t st, float* step_ms, bool backward, int stage = 0) {
  // stage 0: the whole call; 1: only the once-per-call part (slice-invariant steps, pinned operand images);
  // 2: only the slice loop, on a workspace that stage 1 prepared with the same inputs
  TQ_REQUIRE(ws_bytes >= tq_tn_workspace_bytes(p, B), TQ_E_WORKSPACE, "tq_tn_contract: workspace too small");
  const int n_in = p->n_in;
  const int n_steps = (int)p->steps.size();
  InputRef* table = (InputRef*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  cx<R>* shared = (cx<R>*)((uint8_t*)table + tn_table_bytes(p));
  cx<R>* perset = shared + p->arena_shared;
  if (p->d_fsteps || p->d_chain_steps) {
    std::vector<InputRef> host(n_in);
    for (int t = 0; t < n_in; ++t) {
      host[t].ptr = inputs[t];
      host[t].stride = p->in_batched[t] ? strides[t] : 0;
    }
    TQ_CUDA_OK(cudaMemcpyAsync(table, host.data(), n_in * sizeof(InputRef), cudaMemcpyHostToDevice, st));
  }
  uint8_t* pinned = (uint8_t*)(((uintptr_t)(perset + p->arena_set * B) + 1023) & ~(uintptr_t)1023);
  std::vector<int64_t> pin_off;
  uint8_t* images = pinned + tn_pinned_bytes(p, B, &pin_off);
  std::vector<cudaEvent_t> ev;
  if (step_ms) {
    ev.resize(3 * (size_t)n_steps + 2);
    for (auto& e : ev) TQ_CUDA_OK(cudaEventCreate(&e));
  }

  auto tensor_ptr = [&](int t, int64_t slice, const cx<R>*& ptr, int64_t& stride) {
    if (t < n_in) {
      int64_t off = 0;
      for (size_t i = 0; i < p->slice_tensor.size(); ++i)
        if (p->slice_tensor[i] == t && ((slice >> p->slice_ord[i]) & 1)) off |= (int64_t)1 << p->slice_bit[i];
      ptr = (const cx<R>*)inputs[t] + off;
      stride = p->in_batched[t] ? strides[t] : 0;
    } else {
      const int s = t - n_in;
      if (p->arena_const[s]) {
        ptr = shared + p->arena_off[s];
        stride = 0;
      } else {
        ptr = perset + p->arena_off[s];
        stride = p->arena_set;
      }
    }
  };
  // pin_only: pack the pinned (slice-invariant) operand images of step s and return
  auto run_step = [&](int s, int64_t slice, bool pin_only) -> int {
    const tq_tn_step& stp = p->steps[s];
    const StepDev& d = p->dev[s];
    const cx<R>*a, *b, *c0;
    int64_t sa, sb, sc;
    tensor_ptr(stp.lhs, slice, a, sa);
    tensor_ptr(stp.rhs, slice, b, sb);
    tensor_ptr(n_in + s, slice, c0, sc);
    cx<R>* c = const_cast<cx<R>*>(c0);
    const int64_t sets = p->dep_batch[s] ? B : 1;
    const int64_t n_out_elems = (int64_t)1 << (stp.n_m + stp.n_n + stp.n_b);
    const int kernel = tq_tn_plan_step_kernel(p, s);
    if (pin_only) {
      if constexpr (sizeof(R) == 4) {
        const TcStep& T = p->tc[s];
        if (kernel == 2 && (T.pin_a || T.pin_b))
          return run_step_tc(p, s, a, sa, b, sb, c, sc, sets, pinned + std::max<int64_t>(0, pin_off[2 * s]),
                             pinned + std::max<int64_t>(0, pin_off[2 * s + 1]), nullptr, T.pin_a, T.pin_b, false, st,
                             nullptr);
      }
      return TQ_OK;
    }
    const bool timed = step_ms && (slice == s_begin || !p->dep_slice[s]);
    if (timed) TQ_CUDA_OK(cudaEventRecord(ev[3 * s], st));
    if (kernel == 6) {  // gradient seed: `out` is grad_out in this mode
      FinalDev fs;
      memset(&fs, 0, sizeof(fs));
      fs.rank = p->n_out;
      for (int j = 0; j < p->n_out; ++j) fs.pos[p->final_perm[j]] = (int8_t)j;
      const int64_t n_f = (int64_t)1 << p->n_out;
      k_tn_seed<R><<<dim3((unsigned)((n_f + 255) / 256), (unsigned)sets), 256, 0, st>>>((const cx<R>*)out, n_f, c, sc,
                                                                                      fs, n_f);
    } else if (kernel == 2) {
      if constexpr (sizeof(R) == 4) {
        const TcStep& T = p->tc[s];
        const int64_t nz = sets << stp.n_b;
        uint8_t* img_a = T.pin_a ? pinned + pin_off[2 * s] : images;
        uint8_t* img_b = T.pin_b ? pinned + pin_off[2 * s + 1] : images + (T.pin_a ? 0 : T.img_a_z * nz);
        uint8_t* partials = images + (T.pin_a ? 0 : T.img_a_z * nz) + (T.pin_b ? 0 : T.img_b_z * nz);
        partials = (uint8_t*)(((uintptr_t)partials + 255) & ~(uintptr_t)255);
        int rc = run_step_tc(p, s, a, sa, b, sb, c, sc, sets, img_a, img_b, partials, !T.pin_a, !T.pin_b, true, st,
                             timed ? ev[3 * s + 1] : nullptr);
        if (rc) return rc;
      }
    } else if (kernel == 5) {
      const ApplyDev& ad = p->apply[s];
      const bool srhs = p->apply_small_rhs[s] != 0;
      const cx<R>* sm = srhs ? b : a;
      const cx<R>* bg = srhs ? a : b;
      const int64_t ssm = srhs ? sb : sa, sbg = srhs ? sa : sb;
      const int64_t nthr = (int64_t)1 << (ad.n_f + ad.n_b);
      const dim3 grid((unsigned)((nthr + 255) / 256), (unsigned)sets);
      TQ_REQUIRE(sets < 65536, TQ_E_UNSUPPORTED, "tq_tn_contract: step too large");
      const int kc = ad.n_k <= 1 ? 0 : ad.n_k - 1, sc2 = ad.n_s <= 1 ? 0 : ad.n_s - 1;  // classes 2, 4, 8, 16
#define TQ_APPLY(KM, SM) k_tn_apply<R, KM, SM><<<grid, 256, 0, st>>>(sm, ssm, bg, sbg, c, sc, ad, nthr)
#define TQ_APPLY_ROW(KM)                 \
  switch (sc2) {                         \
    case 0: TQ_APPLY(KM, 2); break;      \
    case 1: TQ_APPLY(KM, 4); break;      \
    case 2: TQ_APPLY(KM, 8); break;      \
    default: TQ_APPLY(KM, 16); break;    \
  }
      switch (kc) {
        case 0: TQ_APPLY_ROW(2); break;
        case 1: TQ_APPLY_ROW(4); break;
        case 2: TQ_APPLY_ROW(8); break;
        default: TQ_APPLY_ROW(16); break;
      }
#undef TQ_APPLY_ROW
#undef TQ_APPLY
    } else if (kernel == 3) {
      const int nblk = std::min(DOT_BLOCKS, 1 << (stp.n_k - DOT_LO));
      TQ_REQUIRE(sets < 65536, TQ_E_UNSUPPORTED, "tq_tn_contract: step too large");
      cx<R>* partial = reinterpret_cast<cx<R>*>(images);
      k_tn_dot<R><<<dim3((unsigned)nblk, (unsigned)n_out_elems, (unsigned)sets), 256, 0, st>>>(a, sa, b, sb, partial, d);
      k_tn_dot_sum<R><<<dim3((unsigned)((n_out_elems + 63) / 64), (unsigned)sets), 64, 0, st>>>(
          partial, nblk, c, sc, (int)n_out_elems);
    } else if (kernel == 1) {
      const int64_t blocks = n_out_elems >> 12;
      TQ_REQUIRE(blocks < ((int64_t)1 << 31) && sets < 65536, TQ_E_UNSUPPORTED, "tq_tn_contract: step too large");
      static const bool no_dmma = getenv("TQ_TN_NO_DMMA") != nullptr;  // A/B experiments only
      if (sizeof(R) == 8 && !no_dmma) {  // complex128: FP64 tensor cores
        if constexpr (sizeof(R) == 8)
          k_tn_gemm_dmma<<<dim3((unsigned)blocks, (unsigned)sets), 256, 0, st>>>(a, sa, b, sb, c, sc, d);
      } else {
        k_tn_gemm<R><<<dim3((unsigned)blocks, (unsigned)sets), 256, 0, st>>>(a, sa, b, sb, c, sc, d);
      }
    } else {
      const int64_t blocks = (n_out_elems + 255) / 256;
      TQ_REQUIRE(blocks < ((int64_t)1 << 31) && sets < 65536, TQ_E_UNSUPPORTED, "tq_tn_contract: step too large");
      k_tn_step<R><<<dim3((unsigned)blocks, (unsigned)sets), 256, 0, st>>>(a, sa, b, sb, c, sc, d, n_out_elems);
    }
    TQ_CUDA_OK(cudaGetLastError());
    if (kernel != 2) g_tn_launches += kernel == 3 ? 2 : 1;   // tensor-core steps are counted in run_step_tc
    if (timed) {
      if (kernel != 2) TQ_CUDA_OK(cudaEventRecord(ev[3 * s + 1], st));
      TQ_CUDA_OK(cudaEventRecord(ev[3 * s + 2], st));
    }
    return TQ_OK;
  };
  auto run_item = [&](const SchedItem& it, int64_t slice) -> int {
    if (it.step >= 0) return run_step(it.step, slice, false);
    const int first = it.members.front();
    const bool timed = step_ms && (slice == s_begin || !p->dep_slice[first]);
    if (timed) TQ_CUDA_OK(cudaEventRecord(ev[3 * first], st));
    const int64_t sets = it.batched ? B : 1;
    if (it.chain >= 0) {
      const ChainRun& run = p->chains[it.chain];
      const cx<R>*src, *dst0;
      int64_t ss, sd;
      tensor_ptr(run.big_in, slice, src, ss);
      tensor_ptr(n_in + run.last, slice, dst0, sd);
      const size_t smem = (((size_t)1 << run.hdr.L) + (size_t)run.hdr.gate_entries) * sizeof(cx<R>);
      TQ_REQUIRE(sets < 65536, TQ_E_UNSUPPORTED, "tq_tn_contract: step too large");
      TQ_CUDA_OK(cudaFuncSetAttribute(k_tn_chain<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tn_chain<R><<<dim3(1u << run.hdr.n_glob, (unsigned)sets), 256, smem, st>>>(
          run.hdr, p->d_chain_steps + run.step_begin, (const InputRef*)table, shared, perset, p->arena_set, slice, src,
          ss, const_cast<cx<R>*>(dst0), sd);
      TQ_CUDA_OK(cudaGetLastError());
      g_tn_launches += 1;
      if (timed) {
        TQ_CUDA_OK(cudaEventRecord(ev[3 * first + 1], st));
        TQ_CUDA_OK(cudaEventRecord(ev[3 * first + 2], st));
      }
      return TQ_OK;
    }
    // one CTA per parameter set (2 CTAs per SM); a run shared by all sets gets a cluster of FUSE_CLUSTER CTAs
    if (sets == 1) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(FUSE_CLUSTER);
      cfg.blockDim = dim3(1024);
      cfg.stream = st;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = FUSE_CLUSTER;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      TQ_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tn_fused<R, 1024, FUSE_CLUSTER>, (const FusedStep*)(p->d_fsteps + it.fs_begin),
                                    (const int32_t*)(p->d_levels + it.lv_begin),
                                    (const int32_t*)(p->d_levels + it.lv_begin + it.n_levels + 1), it.n_levels,
                                    (const uint32_t*)p->d_micro, (const InputRef*)table, shared, perset, p->arena_set,
                                    slice));
    } else {
      k_tn_fused<R, 512, 1><<<(unsigned)sets, 512, 0, st>>>(p->d_fsteps + it.fs_begin, p->d_levels + it.lv_begin,
                                                            p->d_levels + it.lv_begin + it.n_levels + 1, it.n_levels,
                                                            p->d_micro, table, shared, perset, p->arena_set, slice);
    }
    TQ_CUDA_OK(cudaGetLastError());
    g_tn_launches += 1;
    if (timed) {
      TQ_CUDA_OK(cudaEventRecord(ev[3 * first + 1], st));
      TQ_CUDA_OK(cudaEventRecord(ev[3 * first + 2], st));
    }
    return TQ_OK;
  };
  int rc;
  if (backward) {  // reverse pass of slice s_begin over the intermediates its forward call left in this workspace
    for (const SchedItem& it : p->items[2])
      if ((rc = run_item(it, s_begin))) return rc;
    return TQ_OK;
  }
  // once-per-call items, then the slice loop
  if (stage != 2) {
    for (const SchedItem& it : p->items[0])
      if ((rc = run_item(it, 0))) return rc;
    if (step_ms) TQ_CUDA_OK(cudaEventRecord(ev[3 * n_steps], st));
    for (int s = 0; s < n_steps; ++s)
      if (p->dep_slice[s] && (rc = run_step(s, s_begin, true))) return rc;
    if (step_ms) TQ_CUDA_OK(cudaEventRecord(ev[3 * n_steps + 1], st));
  }
  if (stage == 1) return TQ_OK;
  FinalDev f;
  memset(&f, 0, sizeof(f));
  f.rank = p->n_out;
  for (int j = 0; j < p->n_out; ++j) f.pos[p->final_perm[j]] = (int8_t)j;
  const int64_t n_final = (int64_t)1 << p->n_out;
  const int n_fwd = p->n_fwd;  // the result is the last FORWARD tensor (a reverse pass may follow in the step list)
  const bool last_dep_slice = n_fwd ? p->dep_slice[n_fwd - 1] != 0 : false;
  for (int64_t slice = s_begin; slice < s_end; ++slice) {
    for (const SchedItem& it : p->items[1])
      if ((rc = run_item(it, slice))) return rc;
    const cx<R>* last;
    int64_t sl;
    tensor_ptr(n_in + n_fwd - 1, slice, last, sl);
    const int64_t sets = p->dep_batch[n_fwd - 1] ? B : 1;
    k_tn_final<R><<<dim3((unsigned)((n_final + 255) / 256), (unsigned)sets), 256, 0, st>>>(last, sl, (cx<R>*)out,
                                                                                          n_final, f, n_final);
    TQ_CUDA_OK(cudaGetLastError());
    g_tn_launches += 1;
    if (!last_dep_slice) break;  // nothing depends on the slice index: one pass is the whole sum
  }
  if (step_ms) {
    TQ_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<char> has_ev(n_steps, 1);
    for (int phase = 0; phase < 3; ++phase)
      for (const SchedItem& it : p->items[phase])
        for (size_t i = (phase == 2 ? 0 : 1); i < it.members.size(); ++i) has_ev[it.members[i]] = 0;
    for (size_t s2 = 0; s2 < p->phase.size(); ++s2)
      if (p->phase[s2] == 2) has_ev[s2] = 0;  // the profiling twin times the forward pass only
    for (int s = 0; s < n_steps; ++s) {
      float whole = 0, pack = 0;
      if (!has_ev[s]) {
        step_ms[2 * s] = step_ms[2 * s + 1] = 0.f;
        continue;
      }
      TQ_CUDA_OK(cudaEventElapsedTime(&whole, ev[3 * s], ev[3 * s + 2]));
      if (tq_tn_plan_step_kernel(p, s) == 2) TQ_CUDA_OK(cudaEventElapsedTime(&pack, ev[3 * s], ev[3 * s + 1]));
      step_ms[2 * s] = whole;
      step_ms[2 * s + 1] = pack;
    }
    TQ_CUDA_OK(cudaEventElapsedTime(&step_ms[2 * n_steps], ev[3 * n_steps], ev[3 * n_steps + 1]));
    step_ms[2 * n_steps + 1] = 0.f;
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return TQ_OK;
}

}  // namespace tq

static int tn_contract_any(const tq_tn_plan* p, const void* const* inputs, const int64_t* input_strides,
                           int64_t batch, int64_t slice_begin, int64_t slice_end, void* out, void* workspace,
                           size_t workspace_bytes, void* stream, float* step_ms, bool backward = false, int stage = 0) {
  TQ_REQUIRE(p && inputs && (out || stage == 1) && workspace && batch > 0, TQ_E_INVALID,
             "tq_tn_contract: null argument");
  TQ_REQUIRE(!p->steps.empty(), TQ_E_INVALID, "tq_tn_contract: empty plan");
  TQ_REQUIRE_DEVICE(p, "tq_tn_contract");
  const int64_t ns = (int64_t)1 << p->n_sliced;
  TQ_REQUIRE(slice_begin >= 0 && slice_begin <= slice_end && slice_end <= ns, TQ_E_INVALID,
             "tq_tn_contract: slice range [%lld, %lld) outside [0, %lld)", (long long)slice_begin,
             (long long)slice_end, (long long)ns);
  if (p->dtype == TQ_C64)
    return contract_impl<float>(p, inputs, input_strides, batch, slice_begin, slice_end, out, workspace,
                                workspace_bytes, (cudaStream_t)stream, step_ms, backward, stage);
  return contract_impl<double>(p, inputs, input_strides, batch, slice_begin, slice_end, out, workspace,
                               workspace_bytes, (cudaStream_t)stream, step_ms, backward, stage);
}

extern "C" int64_t tq_tn_launch_count(void) { return tq::g_tn_launches.load(); }

extern "C" int tq_tn_gather(const void* gate_mats, const void* adj_mats, int64_t src_stride, const int32_t* idx,
                            int64_t n, void* dst, int64_t batch, int32_t dtype, void* stream) {
  TQ_REQUIRE(gate_mats && adj_mats && idx && dst && n > 0 && batch > 0 && batch < 65536, TQ_E_INVALID,
             "tq_tn_gather: bad arguments");
  const dim3 grid((unsigned)((n + 255) / 256), (unsigned)batch);
  if (dtype == TQ_C64)
    k_tn_gather<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const cx<float>*)gate_mats, (const cx<float>*)adj_mats,
                                                               src_stride, idx, n, (cx<float>*)dst);
  else
    k_tn_gather<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const cx<double>*)gate_mats,
                                                                (const cx<double>*)adj_mats, src_stride, idx, n,
                                                                (cx<double>*)dst);
  TQ_CUDA_OK(cudaGetLastError());
  return TQ_OK;
}

extern "C" int tq_tn_contract(const tq_tn_plan* p, const void* const* inputs, const int64_t* input_strides,
                              int64_t batch, int64_t slice_begin, int64_t slice_end, void* out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return tn_contract_any(p, inputs, input_strides, batch, slice_begin, slice_end, out, workspace, workspace_bytes,
                         stream, nullptr);
}

extern "C" int tq_tn_contract_prepare(const tq_tn_plan* p, const void* const* inputs, const int64_t* input_strides,
                                      int64_t batch, int64_t slice_begin, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  return tn_contract_any(p, inputs, input_strides, batch, slice_begin, slice_begin, nullptr, workspace,
                         workspace_bytes, stream, nullptr, false, 1);
}

extern "C" int tq_tn_contract_slices(const tq_tn_plan* p, const void* const* inputs, const int64_t* input_strides,
                                     int64_t batch, int64_t slice_begin, int64_t slice_end, void* out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  return tn_contract_any(p, inputs, input_strides, batch, slice_begin, slice_end, out, workspace, workspace_bytes,
                         stream, nullptr, false, 2);
}

extern "C" int tq_tn_backward(const tq_tn_plan* p, const void* const* inputs, const int64_t* input_strides,
                              int64_t batch, int64_t slice, const void* grad_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  TQ_REQUIRE(p && p->seed_step >= 0, TQ_E_INVALID, "tq_tn_backward: call tq_tn_plan_enable_backward first");
  TQ_REQUIRE(grad_out, TQ_E_INVALID, "tq_tn_backward: grad_out is null");
  return tn_contract_any(p, inputs, input_strides, batch, slice, slice + 1, const_cast<void*>(grad_out), workspace,
                         workspace_bytes, stream, nullptr, true);
}

extern "C" int tq_tn_profile(const tq_tn_plan* p, const void* const* inputs, const int64_t* input_strides,
                             int64_t batch, int64_t slice, void* out, void* workspace, size_t workspace_bytes,
                             void* stream, float* step_ms) {
  TQ_REQUIRE(step_ms, TQ_E_INVALID, "tq_tn_profile: step_ms is null");
  return tn_contract_any(p, inputs, input_strides, batch, slice, slice + 1, out, workspace, workspace_bytes, stream,
                         step_ms);
}

// ---------------------------------------------------------------------------
// multi-GPU: slice ranges + one NCCL all-reduce, for hosts that do not go through torch.distributed
// ---------------------------------------------------------------------------
struct tq_dist {
  void* comm = nullptr;
  int rank = 0, world = 1;
  // ncclResult_t ncclAllReduce(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*error_string)(int) = nullptr;
};

extern "C" int tq_dist_create(void* nccl_comm, int32_t rank, int32_t world, tq_dist** out) {
  TQ_REQUIRE(out, TQ_E_INVALID, "tq_dist_create: out is null");
  *out = nullptr;
  TQ_REQUIRE(world >= 1 && rank >= 0 && rank < world, TQ_E_INVALID, "tq_dist_create: bad rank %d / world %d", rank, world);
  TQ_REQUIRE(nccl_comm || world == 1, TQ_E_INVALID, "tq_dist_create: a communicator is needed for world > 1");
  std::unique_ptr<tq_dist> d(new tq_dist());
  d->comm = nccl_comm;
  d->rank = rank;
  d->world = world;
  if (world > 1) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the NCCL the host already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    TQ_REQUIRE(h, TQ_E_UNSUPPORTED, "tq_dist_create: libnccl.so.2 not found (%s)", dlerror());
    d->all_reduce = reinterpret_cast<decltype(d->all_reduce)>(dlsym(h, "ncclAllReduce"));
    d->error_string = reinterpret_cast<decltype(d->error_string)>(dlsym(h, "ncclGetErrorString"));
    TQ_REQUIRE(d->all_reduce, TQ_E_UNSUPPORTED, "tq_dist_create: ncclAllReduce not found in libnccl");
  }
  *out = d.release();
  return TQ_OK;
}

extern "C" void tq_dist_destroy(tq_dist* d) { delete d; }

extern "C" int tq_dist_slice_range(int64_t n_slices, int32_t rank, int32_t world, int64_t* begin, int64_t* end) {
  TQ_REQUIRE(begin && end && n_slices >= 0 && world >= 1 && rank >= 0 && rank < world, TQ_E_INVALID,
             "tq_dist_slice_range: bad arguments");
  const int64_t per = (n_slices + world - 1) / world;
  *begin = std::min<int64_t>(n_slices, (int64_t)rank * per);
  *end = std::min<int64_t>(n_slices, (int64_t)(rank + 1) * per);
  return TQ_OK;
}

extern "C" int tq_dist_allreduce(const tq_dist* d, void* buf, int64_t count, int32_t dtype, void* stream) {
  TQ_REQUIRE(d && count >= 0, TQ_E_INVALID, "tq_dist_allreduce: bad arguments");
  if (d->world == 1 || count == 0) return TQ_OK;
  TQ_REQUIRE(buf, TQ_E_INVALID, "tq_dist_allreduce: buf is null");
  const int nccl_dtype = dtype == TQ_C64 ? 7 : 8;  // ncclFloat32 = 7, ncclFloat64 = 8; ncclSum = 0
  const int rc = d->all_reduce(buf, buf, (size_t)count, nccl_dtype, 0, d->comm, (cudaStream_t)stream);
  TQ_REQUIRE(rc == 0, TQ_E_CUDA, "ncclAllReduce failed: %s", d->error_string ? d->error_string(rc) : "?");
  return TQ_OK;
}

extern "C" int tq_tn_contract_sharded(const tq_tn_plan* p, const tq_dist* d, const void* const* inputs,
                                      const int64_t* input_strides, int64_t batch, void* out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  TQ_REQUIRE(p && d && out, TQ_E_INVALID, "tq_tn_contract_sharded: null argument");
  const int64_t ns = (int64_t)1 << p->n_sliced;
  int64_t b = 0, e = ns;
  if (ns > 1) {
    int rc = tq_dist_slice_range(ns, d->rank, d->world, &b, &e);
    if (rc) return rc;
  } else if (d->rank != 0) {
    e = 0;  // an unsliced plan: rank 0's contraction is the whole result
  }
  if (e > b) {
    int rc = tq_tn_contract(p, inputs, input_strides, batch, b, e, out, workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  bool any_batched = false;
  for (int t = 0; t < p->n_in; ++t) any_batched |= p->in_batched[t] != 0;
  const int64_t reals = 2 * (any_batched ? batch : 1) * ((int64_t)1 << p->n_out);
  return tq_dist_allreduce(d, out, reals, p->dtype, stream);
}
