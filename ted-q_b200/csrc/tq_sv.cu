// tq_sv.cu — host side of the state-vector engine: plan construction (gate structure,
// fusion, sweep scheduling, op/payload streams) and kernel orchestration behind the C ABI.
//
// Replaces, for backend="pytorch_b200", the reference's per-call Python loop
//   psi <- permute(tensordot(G, psi, axes))      tedq/backends/pytorch_backend.py:358-380
// its gate-tensor builders (:579-1188), measurements (:393-498) and autograd through them.
// Amplitude index bit b (0 = fastest) <-> qubit n-1-b  (pytorch_backend.py:513-522).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <complex>
#include <memory>
#include <vector>

#include "tq_sv_kernels.cuh"
#include "tq_sv_rg.cuh"

namespace tq {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

typedef std::complex<double> zc;

struct HostGate {  // one circuit gate, structure resolved
  int kind = 0, nq = 0;
  int qubits[3] = {0, 0, 0};
  int pidx[3] = {-1, -1, -1};
  double pconst[3] = {0, 0, 0};
  int nparams = 0;   // parameter slots of the kind
  int ntrain = 0;    // trainable slots
  bool noop = false;
  int full_off = -1;  // FIXED: full matrix in the fixed pool
  // structure of the gate on its own (single-member block)
  int cls = OP_DENSE;
  std::vector<int> targets, controls;  // qubits
  int red_off = -1;                    // FIXED: reduced matrix (dense block or diagonal) in the fixed pool
  int red_count = 0;
  bool real = false;   // every entry of the (full) matrix is real for every parameter value: RY, CRY, H, X, Z, CNOT, ...
  bool diag1 = false;  // one-qubit diagonal gate without controls: RZ, PhaseShift, S, T, Z — joins a diagonal layer
};

struct HostBlock {
  std::vector<int> qubits;   // block qubit order (MSB first)
  std::vector<int> members;  // gate indices, circuit order
  bool alive = true;
  int nderiv = 0;
  bool real = true;    // every member is real: the block's product is real (half the multiplies, only Re W needed)
  bool diag1 = false;  // a lone one-qubit diagonal gate kept out of dense blocks: merged into a diagonal layer op
  bool is_x = false;   // a lone fixed gate whose target block is exactly PauliX (CNOT, Toffoli, X): a permutation
  uint32_t gen = 0;    // register groups: generator codes (first | last << 4) when every trainable slot belongs to a
                       // Pauli rotation that is the block's first or last member (tq_sv_rg.cuh: rg_bwd_d1_gen)
  // resolved op
  int cls = OP_DENSE;
  std::vector<int> targets, controls;
  int count = 0;  // matrix entries
  int pidx[MAX_BLOCK_DERIV];
  int mat_block = -1;  // index into MatBlock table
};

struct Sweep {
  Geom geom;
  std::vector<int> bits;
  int op_begin = 0, op_end = 0, n_gates = 0;
  int slot_begin = 0, n_dslots = 0;
  int chunk_begin = 0, n_chunks = 0;
};

}  // namespace tq

using namespace tq;

struct tq_plan {
  int n = 0, n_params = 0, dtype = TQ_C64, n_gates = 0;
  int device = -1;  // CUDA device that owns the plan's tables
  std::vector<HostGate> gates;
  std::vector<HostBlock> blocks;  // alive blocks only, execution order
  std::vector<MatBlock> mblocks;
  std::vector<MatInstr> minstrs;
  int64_t stride_f = 0, stride_b = 0;  // payload entries per parameter set
  std::vector<tq_meas_desc> meas;
  std::vector<DevMeas> dmeas;
  int n_slots = 0;
  int64_t out_reals = 0;
  std::vector<zc> fixed;
  bool fwd_full = false, bwd_full = false, sv_ok = true;
  bool structure = false;  // tq_plan_opts.structure: the sweeps run the kernel instantiations with the real / diagonal-layer paths
  bool rg = false;         // tq_plan_opts.structure = 2 and the circuit qualifies: register-group sweeps (tq_sv_rg.cuh)
  int n_rg[2] = {0, 0};    // register groups emitted per direction
  int n_rg_folded[2] = {0, 0};  // X / CNOT blocks folded into group load / store addresses
  int m_f = 0, m_b = 0, coalesce = 0, threads_f = 256, threads_b = 256, fuse = 1;
  std::vector<Sweep> fwd, bwd;
  int n_ops[2] = {0, 0}, n_dl[2] = {0, 0}, n_dl_members[2] = {0, 0}, n_real[2] = {0, 0};  // emitted ops per direction
  // device
  void* d_fixed = nullptr;
  void* d_init = nullptr;
  OpDesc* d_ops_f = nullptr;
  OpDesc* d_ops_b = nullptr;
  ChunkInfo* d_chunks_f = nullptr;
  ChunkInfo* d_chunks_b = nullptr;
  MatBlock* d_mblocks = nullptr;
  MatInstr* d_minstrs = nullptr;
  DevMeas* d_meas = nullptr;
  int32_t* d_slot_pidx = nullptr;
  std::vector<GateT> gate_t;
  int64_t gate_t_total = 0;
  GateT* d_gate_t = nullptr;
  // scratch for tq_execute_host
  void* h_dev = nullptr;
  size_t h_dev_bytes = 0;
  cudaStream_t h_stream = nullptr;
};

namespace tq {

static size_t csize(int dtype) { return dtype == TQ_C64 ? 8 : 16; }
static size_t rsize(int dtype) { return dtype == TQ_C64 ? 4 : 8; }

// Peel control qubits off a fixed matrix, classify the rest (dense / diagonal).
static void classify_fixed(const zc* U, int k, const int* qubits, HostGate& g, std::vector<zc>& reduced) {
  const int D = 1 << k;
  bool ident = true;
  for (int r = 0; r < D && ident; ++r)
    for (int c = 0; c < D && ident; ++c)
      if (U[r * D + c] != (r == c ? zc(1, 0) : zc(0, 0))) ident = false;
  if (ident) {
    g.noop = true;
    return;
  }
  std::vector<int> is_ctrl(k, 0);
  for (int t = 0; t < k; ++t) {
    const int bt = k - 1 - t;
    bool ok = true;
    for (int r = 0; r < D && ok; ++r)
      for (int c = 0; c < D && ok; ++c) {
        int br = (r >> bt) & 1, bc = (c >> bt) & 1;
        zc v = U[r * D + c];
        if (br != bc) {
          if (v != zc(0, 0)) ok = false;
        } else if (br == 0) {
          if (v != (r == c ? zc(1, 0) : zc(0, 0))) ok = false;
        }
      }
    is_ctrl[t] = ok;
  }
  // a gate whose every qubit qualifies as a control (Z, S, T, CZ, ...) is a phase on the
  // all-ones subspace: keep the last qubit as the (diagonal) target
  bool all = true;
  for (int t = 0; t < k; ++t) all = all && is_ctrl[t];
  if (all) is_ctrl[k - 1] = 0;
  std::vector<int> tq_, cq_;
  for (int t = 0; t < k; ++t) (is_ctrl[t] ? cq_ : tq_).push_back(t);
  const int kt = (int)tq_.size();
  const int Dt = 1 << kt;
  reduced.assign((size_t)Dt * Dt, zc(0, 0));
  auto full_index = [&](int sub) {
    int idx = 0;
    for (int t : cq_) idx |= 1 << (k - 1 - t);
    for (int j = 0; j < kt; ++j)
      if ((sub >> (kt - 1 - j)) & 1) idx |= 1 << (k - 1 - tq_[j]);
    return idx;
  };
  for (int r = 0; r < Dt; ++r)
    for (int c = 0; c < Dt; ++c) reduced[r * Dt + c] = U[full_index(r) * D + full_index(c)];
  bool diag = true;
  for (int r = 0; r < Dt && diag; ++r)
    for (int c = 0; c < Dt && diag; ++c)
      if (r != c && reduced[r * Dt + c] != zc(0, 0)) diag = false;
  g.targets.clear();
  g.controls.clear();
  for (int t : tq_) g.targets.push_back(qubits[t]);
  for (int t : cq_) g.controls.push_back(qubits[t]);
  if (diag) {
    g.cls = OP_DIAG;
    std::vector<zc> d(Dt);
    for (int r = 0; r < Dt; ++r) d[r] = reduced[r * Dt + r];
    reduced = d;
  } else {
    g.cls = OP_DENSE;
  }
}

static void build_geom(int n, const std::vector<int>& bits, Geom& g) {
  memset(&g, 0, sizeof(g));
  g.n = n;
  g.m = (int)bits.size();
  int i = 0;
  while (i < g.m) {
    int j = i;
    while (j + 1 < g.m && bits[j + 1] == bits[j] + 1) ++j;
    g.lsrc[g.nl] = (int8_t)i;
    g.llen[g.nl] = (int8_t)(j - i + 1);
    g.ldst[g.nl] = (int8_t)bits[i];
    ++g.nl;
    i = j + 1;
  }
  std::vector<int> rest;
  std::vector<char> in(n, 0);
  for (int b : bits) in[b] = 1;
  for (int b = 0; b < n; ++b)
    if (!in[b]) rest.push_back(b);
  i = 0;
  const int nr = (int)rest.size();
  while (i < nr) {
    int j = i;
    while (j + 1 < nr && rest[j + 1] == rest[j] + 1) ++j;
    g.tsrc[g.nt] = (int8_t)i;
    g.tlen[g.nt] = (int8_t)(j - i + 1);
    g.tdst[g.nt] = (int8_t)rest[i];
    ++g.nt;
    i = j + 1;
  }
}

// Greedy segment scheduler over blocks: walk the not-yet-run blocks in order; a block joins the
// sweep when its qubits fit into the tile and none of them is blocked by an earlier deferred block.
static void schedule(int n, int m, int coalesce, const std::vector<HostBlock>& blocks, const std::vector<int>& order,
                     std::vector<std::vector<int>>& seg_ops, std::vector<std::vector<int>>& seg_bits) {
  seg_ops.clear();
  seg_bits.clear();
  const int G = (int)order.size();
  if (n <= m) {
    std::vector<int> bits(n);
    for (int b = 0; b < n; ++b) bits[b] = b;
    seg_bits.push_back(bits);
    seg_ops.push_back(order);
    return;
  }
  std::vector<char> done(G, 0);
  int remaining = G;
  while (remaining > 0) {
    std::vector<char> in_s(n, 0), blocked(n, 0);
    int ns = 0;
    for (int b = 0; b < coalesce; ++b) {
      in_s[b] = 1;
      ++ns;
    }
    int nblocked = 0;
    std::vector<int> gl;
    for (int oi = 0; oi < G && nblocked < n; ++oi) {
      if (done[oi]) continue;
      const HostBlock& op = blocks[order[oi]];
      bool blk = false;
      int need = 0;
      for (int q : op.qubits) {
        int b = n - 1 - q;
        if (blocked[b]) blk = true;
        if (!in_s[b]) ++need;
      }
      if (!blk && ns + need <= m) {
        for (int q : op.qubits) {
          int b = n - 1 - q;
          if (!in_s[b]) {
            in_s[b] = 1;
            ++ns;
          }
        }
        gl.push_back(order[oi]);
        done[oi] = 1;
        --remaining;
      } else {
        for (int q : op.qubits) {
          int b = n - 1 - q;
          if (!blocked[b]) {
            blocked[b] = 1;
            ++nblocked;
          }
        }
      }
    }
    for (int b = 0; b < n && ns < m; ++b)
      if (!in_s[b]) {
        in_s[b] = 1;
        ++ns;
      }
    std::vector<int> bits;
    for (int b = 0; b < n; ++b)
      if (in_s[b]) bits.push_back(b);
    seg_bits.push_back(bits);
    seg_ops.push_back(gl);
  }
}

static void make_desc(const HostBlock& h, int n, const std::vector<int>& bits, bool c64, OpDesc& d) {
  memset(&d, 0, sizeof(d));
  auto local = [&](int q) {
    int b = n - 1 - q;
    for (size_t j = 0; j < bits.size(); ++j)
      if (bits[j] == b) return (int)j;
    return -1;
  };
  std::vector<int> lt, lc;
  for (int q : h.targets) lt.push_back(local(q));
  for (int q : h.controls) lc.push_back(local(q));
  const int k = (int)lt.size();
  bool bit0_ctrl = false, bit0_tgt = false;
  for (int b : lc) bit0_ctrl |= (b == 0);
  for (int b : lt) bit0_tgt |= (b == 0);
  d.k = (uint8_t)k;
  d.nderiv = (uint8_t)h.nderiv;
  d.count = (uint32_t)h.count;
  d.pad = (uint32_t)h.cls;
  int shift = 0;
  std::vector<int> ins;
  if (h.cls == OP_DENSE && h.real && k <= 2) {  // real twins of the complex paths, chosen by the same rules
    if (k == 1) {
      if (c64 && lt[0] == 0) {
        d.path = P_R1P;
        shift = 1;
        ins = lc;
      } else if (c64 && !bit0_ctrl) {
        d.path = P_R1V;
        shift = 1;
        ins = lc;
        ins.push_back(lt[0]);
      } else {
        d.path = P_R1S;
        ins = lc;
        ins.push_back(lt[0]);
      }
    } else {
      d.path = (c64 && !bit0_ctrl && !bit0_tgt) ? P_R2V : P_R2S;
      shift = d.path == P_R2V ? 1 : 0;
      ins = lc;
      ins.push_back(lt[0]);
      ins.push_back(lt[1]);
    }
  } else if (h.cls == OP_DENSE) {
    if (k == 1) {
      if (c64 && lt[0] == 0) {
        d.path = P_D1P;
        shift = 1;
        ins = lc;
      } else if (c64 && !bit0_ctrl) {
        d.path = P_D1V;
        shift = 1;
        ins = lc;
        ins.push_back(lt[0]);
      } else {
        d.path = P_D1S;
        ins = lc;
        ins.push_back(lt[0]);
      }
    } else if (k == 2) {
      d.path = (c64 && !bit0_ctrl && !bit0_tgt) ? P_D2V : P_D2S;
      shift = d.path == P_D2V ? 1 : 0;
      ins = lc;
      ins.push_back(lt[0]);
      ins.push_back(lt[1]);
    } else {
      d.path = P_GEN;
      ins = lc;
      for (int b : lt) ins.push_back(b);
    }
  } else {
    if (k == 1) {
      if (c64 && !bit0_ctrl) {
        d.path = P_G1V;
        shift = 1;
        ins = lc;
      } else {
        d.path = P_G1S;
        ins = lc;
      }
    } else {
      d.path = P_GEN;
      ins = lc;
    }
  }
  std::sort(ins.begin(), ins.end());
  d.nins = (uint8_t)ins.size();
  for (size_t j = 0; j < ins.size() && j < 4; ++j) d.ins[j] = (uint8_t)(ins[j] - shift);
  for (int b : lc) d.cmask |= 1u << (b - shift);
  for (int t = 0; t < k && t < 4; ++t) {
    // P_G1V keeps the target in amplitude-bit units (it may be bit 0); every other path in path units
    d.tpos[t] = (uint8_t)(d.path == P_G1V ? lt[t] : lt[t] - shift);
  }
}

template <typename T>
static int upload(const std::vector<T>& v, T** dptr) {
  *dptr = nullptr;
  if (v.empty()) return TQ_OK;
  TQ_CUDA_OK(cudaMalloc((void**)dptr, v.size() * sizeof(T)));
  TQ_CUDA_OK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return TQ_OK;
}

static int upload_complex(const zc* src, size_t count, int dtype, void** dptr) {
  *dptr = nullptr;
  if (count == 0) return TQ_OK;
  TQ_CUDA_OK(cudaMalloc(dptr, count * csize(dtype)));
  if (dtype == TQ_C64) {
    std::vector<float> tmp(2 * count);
    for (size_t i = 0; i < count; ++i) {
      tmp[2 * i] = (float)src[i].real();  // same rounding as torch .type(complex64), pytorch_backend.py:573
      tmp[2 * i + 1] = (float)src[i].imag();
    }
    TQ_CUDA_OK(cudaMemcpy(*dptr, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    TQ_CUDA_OK(cudaMemcpy(*dptr, src, count * sizeof(zc), cudaMemcpyHostToDevice));
  }
  return TQ_OK;
}

// host mirror of the device embedding (tq_sv_kernels.cuh: embed_elem) for constant folding
static void host_embed(const zc* G, int gdim, int embed, int Dd, zc* E) {
  if (Dd == 2) {
    for (int i = 0; i < 4; ++i) E[i] = G[i];
    return;
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      const int r0 = r >> 1, r1 = r & 1, c0 = c >> 1, c1 = c & 1;
      if (gdim == 2) {
        if (embed == 0)
          E[r * 4 + c] = (r1 == c1) ? G[r0 * 2 + c0] : zc(0, 0);
        else
          E[r * 4 + c] = (r0 == c0) ? G[r1 * 2 + c1] : zc(0, 0);
      } else {
        int rr = r, cc = c;
        if (embed == 3) {
          rr = (r1 << 1) | r0;
          cc = (c1 << 1) | c0;
        }
        E[r * 4 + c] = G[rr * 4 + cc];
      }
    }
}

static int block_pay_entries(int count, int nderiv, bool backward) {
  int e = count * (1 + (backward ? nderiv : 0));
  return (e + 1) & ~1;  // 16-byte granularity for complex64
}

}  // namespace tq

extern "C" {

int tq_abi_version(void) { return TQ_ABI_VERSION; }
const char* tq_last_error(void) { return tq::g_err; }

int tq_sv_axes_perm(int32_t n_qubits, const int32_t* qubits, int32_t nq, int32_t* gate_pos, int32_t* perm) {
  // tensordot puts the gate's output axes first, followed by the untouched state axes in
  // ascending order; perm sends every qubit axis back to its place (compiled_circuit.py:126-198).
  TQ_REQUIRE(n_qubits > 0 && nq > 0 && nq <= n_qubits, TQ_E_INVALID, "tq_sv_axes_perm: bad sizes");
  for (int i = 0; i < nq; ++i) gate_pos[i] = nq == 1 ? 1 : nq + i;
  std::vector<int> where(n_qubits, -1);
  for (int i = 0; i < nq; ++i) {
    TQ_REQUIRE(qubits[i] >= 0 && qubits[i] < n_qubits, TQ_E_INVALID, "tq_sv_axes_perm: qubit out of range");
    where[qubits[i]] = i;
  }
  int next = nq;
  for (int q = 0; q < n_qubits; ++q)
    if (where[q] < 0) where[q] = next++;
  for (int q = 0; q < n_qubits; ++q) perm[q] = where[q];
  return TQ_OK;
}

void tq_plan_destroy(tq_plan* p) {
  if (!p) return;
  cudaFree(p->d_fixed);
  cudaFree(p->d_init);
  cudaFree(p->d_ops_f);
  cudaFree(p->d_ops_b);
  cudaFree(p->d_chunks_f);
  cudaFree(p->d_chunks_b);
  cudaFree(p->d_mblocks);
  cudaFree(p->d_minstrs);
  cudaFree(p->d_meas);
  cudaFree(p->d_slot_pidx);
  cudaFree(p->d_gate_t);
  cudaFree(p->h_dev);
  if (p->h_stream) cudaStreamDestroy(p->h_stream);
  delete p;
}

int tq_plan_create(const tq_gate_desc* gates, int32_t n_gates, const tq_meas_desc* meas, int32_t n_meas,
                   const double* pool, int64_t pool_len, int32_t n_qubits, int32_t n_params, int32_t dtype,
                   const double* init_state, const tq_plan_opts* opts, tq_plan** out) {
  TQ_REQUIRE(out, TQ_E_INVALID, "tq_plan_create: out is null");
  *out = nullptr;
  // beyond 30 qubits a state vector does not fit: the plan then only serves the tensor-network entry points
  // (tq_tn_operands); tq_forward / tq_backward refuse it
  TQ_REQUIRE(n_qubits >= 1 && n_qubits <= 4096, TQ_E_UNSUPPORTED, "tq_plan_create: n_qubits=%d unsupported", n_qubits);
  const bool sv_ok = n_qubits <= 30;
  TQ_REQUIRE(dtype == TQ_C64 || dtype == TQ_C128, TQ_E_INVALID, "tq_plan_create: bad dtype %d", dtype);
  TQ_REQUIRE(n_gates >= 0 && n_meas >= 1 && n_params >= 0, TQ_E_INVALID, "tq_plan_create: bad counts");
  std::unique_ptr<tq_plan, void (*)(tq_plan*)> P(new tq_plan(), tq_plan_destroy);
  tq_plan* p = P.get();
  p->n = n_qubits;
  p->n_params = n_params;
  p->dtype = dtype;
  p->n_gates = n_gates;
  p->sv_ok = sv_ok;
  p->device = tq::current_device();
  const int n = n_qubits;
  const zc* zpool = reinterpret_cast<const zc*>(pool);
  const bool c64 = dtype == TQ_C64;

  // ---- options -------------------------------------------------------------
  int full_f = c64 ? 14 : 13, full_b = c64 ? 13 : 12;  // whole state: <= 128 KiB (fwd) / 2 x 64 KiB (bwd)
  int m_f = c64 ? 13 : 12, m_b = c64 ? 12 : 11;        // tiled: 64 KiB of amplitudes per CTA -> 3 CTAs per SM
  int coalesce = c64 ? 3 : 2;                          // 64-byte contiguous runs
  int threads = 0;
  int fuse = 1;
  // structure-aware ops (tq_plan_opts.structure = 1): real blocks on half-cost paths, one-qubit diagonal gates merged
  // into diagonal-layer passes.  Off by default: on the 20-qubit HEA it halves the multiplies but needs 247 passes
  // over the tile instead of 190 and executes MORE instructions in total (2.70e9 vs 2.53e9 per 16 sets, ncu) — the
  // per-pass overhead (descriptor, matrix load, index arithmetic, barrier) outweighs the saved FMAs
  bool diag_layers = false, want_rg = false, auto_rg = true;
  if (opts) {
    diag_layers = opts->structure == 1;
    want_rg = opts->structure == 2;
    auto_rg = opts->structure == 0 && opts->threads == 0;  // -1: the default sweeps whatever the circuit
    if (opts->structure < -1 || opts->structure > 2) {
      set_error("tq_plan_create: structure = %d (expected -1 .. 2)", opts->structure);
      return TQ_E_INVALID;
    }
    if (opts->max_local_qubits_fwd > 0) m_f = full_f = opts->max_local_qubits_fwd;
    if (opts->max_local_qubits_bwd > 0) m_b = full_b = opts->max_local_qubits_bwd;
    if (opts->coalesce_bits >= 0) coalesce = opts->coalesce_bits;
    threads = opts->threads;
    if (opts->fuse >= 0) fuse = opts->fuse;
  }
  TQ_REQUIRE(m_f <= (c64 ? 14 : 13) && m_b <= (c64 ? 13 : 12), TQ_E_INVALID,
             "tq_plan_create: tile exceeds shared memory");
  if (!sv_ok) {
    m_f = std::min(m_f, 12);
    m_b = std::min(m_b, 11);
  }
  p->fwd_full = n <= full_f;
  p->bwd_full = n <= full_b;
  if (p->fwd_full) m_f = n;
  if (p->bwd_full) m_b = n;
  coalesce = std::min(coalesce, std::min(m_f, m_b) - TQ_MAX_GATE_QUBITS);
  if (coalesce < 0) coalesce = 0;
  TQ_REQUIRE(p->fwd_full || m_f - coalesce >= TQ_MAX_GATE_QUBITS, TQ_E_INVALID, "tq_plan_create: forward tile too small");
  TQ_REQUIRE(p->bwd_full || m_b - coalesce >= TQ_MAX_GATE_QUBITS, TQ_E_INVALID, "tq_plan_create: backward tile too small");
  p->m_f = m_f;
  p->m_b = m_b;
  p->coalesce = coalesce;
  p->fuse = fuse;
  p->structure = diag_layers;
  auto pick_threads = [&](int m) {
    if (threads > 0) return threads;
    int t = 1 << std::max(5, m - 3);
    return std::min(256, std::max(32, t));
  };
  p->threads_f = pick_threads(m_f);
  p->threads_b = pick_threads(m_b);
  // tiled complex64 adjoint sweeps: 128 threads per CTA, 3 CTAs per SM — every thread runs twice the amplitude groups
  // per op, which halves the per-op overhead (matrix loads, W reduction, barrier) per amplitude: 49.2 -> 45.7 ms on
  // the 20-qubit HEA (scripts/c3_threads.py); the forward sweep is faster with 256
  if (threads == 0 && c64 && !p->bwd_full) p->threads_b = 128;
  TQ_REQUIRE(p->threads_f % 32 == 0 && p->threads_f <= 256 && p->threads_b % 32 == 0 && p->threads_b <= 256,
             TQ_E_INVALID, "tq_plan_create: threads must be a multiple of 32, <= 256");

  // ---- gates ----------------------------------------------------------------
  p->gates.resize(n_gates);
  std::vector<char> param_seen(n_params, 0);
  for (int gi = 0; gi < n_gates; ++gi) {
    const tq_gate_desc& g = gates[gi];
    HostGate& h = p->gates[gi];
    TQ_REQUIRE(g.nq >= 1 && g.nq <= TQ_MAX_GATE_QUBITS, TQ_E_UNSUPPORTED, "gate %d: %d qubits unsupported", gi, g.nq);
    h.kind = g.kind;
    h.nq = g.nq;
    for (int i = 0; i < g.nq; ++i) {
      TQ_REQUIRE(g.qubits[i] >= 0 && g.qubits[i] < n, TQ_E_INVALID, "gate %d: qubit %d out of range", gi, g.qubits[i]);
      for (int j = 0; j < i; ++j)
        TQ_REQUIRE(g.qubits[i] != g.qubits[j], TQ_E_INVALID, "gate %d: repeated qubit", gi);
      h.qubits[i] = g.qubits[i];
    }
    if (g.kind == TQ_G_FIXED) {
      const int D = 1 << g.nq;
      TQ_REQUIRE(g.matrix_off >= 0 && g.matrix_off + (int64_t)D * D <= pool_len, TQ_E_INVALID,
                 "gate %d: matrix outside pool", gi);
      std::vector<zc> red;
      classify_fixed(zpool + g.matrix_off, g.nq, h.qubits, h, red);
      h.real = true;
      for (int i = 0; i < D * D; ++i) h.real = h.real && zpool[g.matrix_off + i].imag() == 0.0;
      h.diag1 = !h.noop && h.cls == OP_DIAG && h.targets.size() == 1 && h.controls.empty() && g.nq == 1;
      h.full_off = (int)p->fixed.size();
      p->fixed.insert(p->fixed.end(), zpool + g.matrix_off, zpool + g.matrix_off + D * D);
      if (!h.noop) {
        h.red_off = (int)p->fixed.size();
        h.red_count = (int)red.size();
        p->fixed.insert(p->fixed.end(), red.begin(), red.end());
        if (p->fixed.size() & 1) p->fixed.push_back(zc(0, 0));
      }
      continue;
    }
    int np = 1, nqexp = 1;
    bool ctrl = false;
    h.real = g.kind == TQ_G_RY || g.kind == TQ_G_CRY;
    h.diag1 = g.kind == TQ_G_RZ || g.kind == TQ_G_PHASESHIFT;
    switch (g.kind) {
      case TQ_G_RX: case TQ_G_RY: h.cls = OP_DENSE; break;
      case TQ_G_ROT: h.cls = OP_DENSE; np = 3; break;
      case TQ_G_RZ: case TQ_G_PHASESHIFT: h.cls = OP_DIAG; break;
      case TQ_G_CRX: case TQ_G_CRY: h.cls = OP_DENSE; ctrl = true; nqexp = 2; break;
      case TQ_G_CRZ: case TQ_G_CPHASE: h.cls = OP_DIAG; ctrl = true; nqexp = 2; break;
      default: TQ_REQUIRE(false, TQ_E_INVALID, "gate %d: unknown kind %d", gi, g.kind);
    }
    TQ_REQUIRE(g.nq == nqexp, TQ_E_INVALID, "gate %d: kind %d needs %d qubits", gi, g.kind, nqexp);
    if (ctrl) {
      h.controls.push_back(g.qubits[0]);
      h.targets.push_back(g.qubits[1]);
    } else {
      h.targets.push_back(g.qubits[0]);
    }
    h.nparams = np;
    h.red_count = h.cls == OP_DIAG ? 2 : 4;
    for (int i = 0; i < np; ++i) {
      h.pidx[i] = g.param_idx[i];
      h.pconst[i] = g.param_const[i];
      if (g.param_idx[i] >= 0) {
        TQ_REQUIRE(g.param_idx[i] < n_params, TQ_E_INVALID, "gate %d: parameter index %d out of range", gi,
                   g.param_idx[i]);
        TQ_REQUIRE(!param_seen[g.param_idx[i]], TQ_E_INVALID,
                   "gate %d: flat parameter %d bound twice (binding is positional, one slot each)", gi,
                   g.param_idx[i]);
        param_seen[g.param_idx[i]] = 1;
        ++h.ntrain;
      }
    }
  }

  // ---- operand table for the tensor-network path (every gate as a full [out..., in...] tensor) ------
  {
    int64_t off = 0;
    for (int gi = 0; gi < n_gates; ++gi) {
      const HostGate& g = p->gates[gi];
      GateT t;
      memset(&t, 0, sizeof(t));
      t.kind = g.kind;
      t.nq = g.nq;
      t.fixed_off = g.full_off;
      t.out_off = off;
      for (int i = 0; i < 3; ++i) {
        t.pidx[i] = g.pidx[i];
        t.pconst[i] = g.pconst[i];
      }
      p->gate_t.push_back(t);
      off += (int64_t)1 << (2 * g.nq);
    }
    p->gate_t_total = off;
  }
  if (!sv_ok) {
    int rc0;
    if ((rc0 = upload_complex(p->fixed.data(), p->fixed.size(), dtype, &p->d_fixed))) return rc0;
    if ((rc0 = upload(p->gate_t, &p->d_gate_t))) return rc0;
    *out = P.release();
    return TQ_OK;
  }

  // ---- register-group mode (structure = 2): complex64, tiles of >= 2^9 amplitudes, and every gate either a (controlled)
  // one-target block or a diagonal — anything else (SWAP, a dense two-qubit Unitary) keeps the default sweeps
  bool rg = (want_rg || auto_rg) && c64 && std::min(m_f, m_b) >= RG_MIN_TILE;
  for (int gi = 0; gi < n_gates && rg; ++gi) {
    const HostGate& g = p->gates[gi];
    if (g.noop) continue;
    if (g.cls == OP_DENSE && g.targets.size() != 1) rg = false;
    if (g.cls == OP_DIAG && g.targets.size() > 1 && g.ntrain > 0) rg = false;
  }

  // ---- fusion: runs of gates inside one qubit or one qubit pair become one dense block -------------
  const int pay_cap_entries = CHUNK_PAY_BYTES / (int)csize(dtype);
  std::vector<HostBlock> all;
  auto fuse_gates = [&](bool rg, std::vector<HostBlock>& all) {
    all.clear();
    std::vector<int> owner(n, -1);
    auto can_add = [&](const HostBlock& b, const HostGate& g, int extra_deriv) {
      if (!fuse) return false;
      int nd = b.nderiv + g.ntrain + extra_deriv;
      if (nd > MAX_BLOCK_DERIV) return false;
      int dim = 1 << std::max<int>((int)b.qubits.size(), g.nq);
      if (block_pay_entries(dim * dim, nd, true) > pay_cap_entries) return false;
      return b.members.size() < 256;
    };
    for (int gi = 0; gi < n_gates; ++gi) {
      const HostGate& g = p->gates[gi];
      if (g.noop) continue;
      auto fresh = [&]() {
        HostBlock b;
        for (int i = 0; i < g.nq; ++i) b.qubits.push_back(g.qubits[i]);
        b.members.push_back(gi);
        b.nderiv = g.ntrain;
        b.real = g.real;
        all.push_back(b);
        for (int i = 0; i < g.nq; ++i) owner[g.qubits[i]] = (int)all.size() - 1;
      };
      if (!fuse || g.nq == 3) {
        fresh();
        continue;
      }
      if (rg) {
        // register groups: only runs of one-qubit gates on the SAME qubit are multiplied together (2x2); entanglers
        // stay on their own (a CNOT is a register swap there, not a factor that turns its neighbours into a 4x4)
        const int o = g.nq == 1 ? owner[g.qubits[0]] : -1;
        if (o >= 0 && all[o].qubits.size() == 1 && can_add(all[o], g, 0)) {
          all[o].members.push_back(gi);
          all[o].nderiv += g.ntrain;
        } else {
          fresh();
        }
        continue;
      }
      if (g.nq == 1) {
        int o = owner[g.qubits[0]];
        if (diag_layers && g.diag1) {
          // a one-qubit diagonal gate costs nothing inside a block that is complex anyway; otherwise it stays on its
          // own and is merged with the other diagonal gates of its sweep into ONE phase-table pass (P_DL) — fused
          // into a real block it would double that block's multiplies
          if (o >= 0 && !all[o].diag1 && !all[o].real && all[o].qubits.size() <= 2 && can_add(all[o], g, 0)) {
            all[o].members.push_back(gi);
            all[o].nderiv += g.ntrain;
          } else {
            fresh();
            all.back().diag1 = true;
          }
          continue;
        }
        if (o >= 0 && !all[o].diag1 && all[o].qubits.size() <= 2 && can_add(all[o], g, 0)) {
          all[o].members.push_back(gi);
          all[o].nderiv += g.ntrain;
          all[o].real = all[o].real && g.real;
        } else {
          fresh();
        }
        continue;
      }
      // two-qubit gate
      int o0 = owner[g.qubits[0]], o1 = owner[g.qubits[1]];
      if (o0 >= 0 && o0 == o1 && all[o0].qubits.size() == 2 && can_add(all[o0], g, 0)) {
        all[o0].members.push_back(gi);
        all[o0].nderiv += g.ntrain;
        all[o0].real = all[o0].real && g.real;
        continue;
      }
      HostBlock b;
      b.qubits.push_back(g.qubits[0]);
      b.qubits.push_back(g.qubits[1]);
      b.real = g.real;
      for (int side = 0; side < 2; ++side) {
        int o = side == 0 ? o0 : o1;
        if (o >= 0 && all[o].alive && all[o].qubits.size() == 1 && !all[o].diag1) {
          if (can_add(b, g, all[o].nderiv)) {
            for (int mgi : all[o].members) b.members.push_back(mgi);
            b.nderiv += all[o].nderiv;
            b.real = b.real && all[o].real;
            all[o].alive = false;
          }
        }
      }
      b.members.push_back(gi);
      b.nderiv += g.ntrain;
      all.push_back(b);
      owner[g.qubits[0]] = owner[g.qubits[1]] = (int)all.size() - 1;
    }
  };
  // multiply-adds per evaluation of a fusion result (a fused block is a dense matrix on its qubits; a lone gate keeps
  // its own structure; X / CNOT cost nothing in a register group)
  auto fused_macs = [&](const std::vector<HostBlock>& blocks, bool rg_mode) {
    double macs = 0.0;
    for (const HostBlock& b : blocks) {
      if (!b.alive) continue;
      if (b.members.size() > 1) {
        macs += ldexp(1.0, n + (int)b.qubits.size());
        continue;
      }
      const HostGate& g = p->gates[b.members[0]];
      const bool x = g.kind == TQ_G_FIXED && g.cls == OP_DENSE && g.targets.size() == 1 && g.red_count == 4 &&
                     p->fixed[g.red_off] == zc(0, 0) && p->fixed[g.red_off + 1] == zc(1, 0) &&
                     p->fixed[g.red_off + 2] == zc(1, 0) && p->fixed[g.red_off + 3] == zc(0, 0);
      if (x) continue;  // a permutation: data movement only, in either mode
      double per = ldexp(1.0, (int)g.targets.size());          // dense: 2^t multiply-adds per amplitude it touches
      if (g.cls == OP_DIAG) per = rg_mode && g.targets.size() == 1 ? 2.0 : 1.0;  // (a register group runs it as a 2x2)
      macs += ldexp(1.0, n - (int)g.controls.size()) * per;
    }
    return macs;
  };
  if (rg && auto_rg && !want_rg) {
    // automatic choice: register groups where they cost no more arithmetic than the default fusion (layers of
    // one-qubit gates between sparse entanglers — QNN / HEA / lattice circuits: the same multiply-adds in half the
    // passes over the tile).  Circuits whose pair blocks fold many gates (the many-body-localisation Trotter steps: 18
    // gates per 4x4 block) would run each of those gates as its own 2x2 and keep the default sweeps.
    std::vector<HostBlock> a0, a1;
    fuse_gates(false, a0);
    fuse_gates(true, a1);
    rg = fused_macs(a1, true) <= 1.1 * fused_macs(a0, false);
    // measured on B200 (scripts/rg_small.py, HEA depth 10): the groups win from 12 qubits on (+32 % at 12, +24 % at 14,
    // +26 % at 16, +34 % at 20) and lose at 10 (64 threads per CTA: the whole-state kernels run out of warps)
    if (n < 12) rg = false;
  }
  if (rg && p->bwd_full && n > 12 && !(opts && opts->max_local_qubits_bwd > 0)) {
    // 13 qubits: a whole-state adjoint tile pair is 128 KiB = one CTA of 8 warps per SM; two tiled sweeps of 2^12 keep
    // two CTAs resident
    p->bwd_full = false;
    m_b = 12;
    p->m_b = m_b;
    coalesce = std::min(coalesce, m_b - TQ_MAX_GATE_QUBITS);
    p->coalesce = coalesce;
  }
  p->rg = rg;
  if (rg) {
    // the tile <-> state copies move pairs of amplitudes: at least one coalesce bit.  (Fewer coalesce bits make the
    // sweeps deeper — 20-qubit HEA: 9 + 12 sweeps with 2 bits, 7 + 11 with 1, against 13 + 14 with the default 3 — but
    // the register-group sweeps are bound by instruction issue, not by HBM: 1369 / 1333 / 1405 evaluations/s.)
    coalesce = std::max(coalesce, 1);
    p->coalesce = coalesce;
    p->threads_f = std::min(256, 1 << (m_f - RG_BITS));
    p->threads_b = std::min(256, 1 << (m_b - RG_BITS));
  }
  fuse_gates(rg, all);
  // resolve blocks -> ops + materialisation program
  for (HostBlock& b : all) {
    if (!b.alive) continue;
    MatBlock mb;
    memset(&mb, 0, sizeof(mb));
    mb.instr_begin = (int)p->minstrs.size();
    int dcount = 0;
    const bool single = b.members.size() == 1;
    const int Dd_blk = 1 << (int)b.qubits.size();
    std::vector<zc> fold;  // running product of consecutive FIXED members (host, double precision)
    int n_instr = 0;
    auto flush_fold = [&]() {
      if (fold.empty()) return;
      MatInstr mi;
      memset(&mi, 0, sizeof(mi));
      mi.kind = TQ_G_FIXED;
      mi.nq = (int)b.qubits.size();
      mi.embed = Dd_blk == 4 ? 2 : 0;
      mi.fixed_off = (int)p->fixed.size();
      for (int i = 0; i < 3; ++i) mi.pidx[i] = mi.dsel[i] = -1;
      p->fixed.insert(p->fixed.end(), fold.begin(), fold.end());
      if (p->fixed.size() & 1) p->fixed.push_back(zc(0, 0));
      p->minstrs.push_back(mi);
      ++n_instr;
      fold.clear();
    };
    for (int gi : b.members) {
      const HostGate& g = p->gates[gi];
      int embed = 0;
      if (!single) {
        if (b.qubits.size() == 1) {
          embed = 0;
        } else if (g.nq == 1) {
          embed = g.qubits[0] == b.qubits[0] ? 0 : 1;
        } else {
          embed = g.qubits[0] == b.qubits[0] ? 2 : 3;
        }
      }
      if (!single && g.kind == TQ_G_FIXED) {
        zc E[16], T[16];
        host_embed(p->fixed.data() + g.full_off, 1 << g.nq, embed, Dd_blk, E);
        if (fold.empty()) {
          fold.assign(E, E + Dd_blk * Dd_blk);
        } else {
          for (int r = 0; r < Dd_blk; ++r)
            for (int c = 0; c < Dd_blk; ++c) {
              zc acc(0, 0);
              for (int k = 0; k < Dd_blk; ++k) acc += E[r * Dd_blk + k] * fold[k * Dd_blk + c];
              T[r * Dd_blk + c] = acc;
            }
          fold.assign(T, T + Dd_blk * Dd_blk);
        }
        continue;
      }
      flush_fold();
      MatInstr mi;
      memset(&mi, 0, sizeof(mi));
      mi.kind = g.kind;
      mi.nq = g.nq;
      mi.fixed_off = single ? g.red_off : g.full_off;
      mi.embed = embed;
      for (int i = 0; i < 3; ++i) {
        mi.pidx[i] = g.pidx[i];
        mi.pconst[i] = g.pconst[i];
        mi.dsel[i] = -1;
        if (i < g.nparams && g.pidx[i] >= 0) {
          b.pidx[dcount] = g.pidx[i];
          mi.dsel[i] = dcount++;
        }
      }
      p->minstrs.push_back(mi);
      ++n_instr;
    }
    flush_fold();
    mb.instr_end = (int)p->minstrs.size();
    mb.nderiv = b.nderiv;
    if (p->rg && b.nderiv > 0) {
      auto code = [](const HostGate& g) -> uint32_t {
        switch (g.kind) {
          case TQ_G_RX: case TQ_G_CRX: return RG_PX;
          case TQ_G_RY: case TQ_G_CRY: return RG_PY;
          case TQ_G_RZ: case TQ_G_CRZ: return RG_PZ;
          case TQ_G_PHASESHIFT: case TQ_G_CPHASE: return RG_PP;
          default: return 0u;
        }
      };
      const int first = b.members.front(), last = b.members.back();
      bool ok = true;
      uint32_t gf = 0, gl = 0;
      for (int gi : b.members) {
        const HostGate& g = p->gates[gi];
        if (g.ntrain == 0) continue;
        const uint32_t c = (g.ntrain == 1 && g.nparams == 1) ? code(g) : 0u;
        if (!c)
          ok = false;
        else if (gi == last)
          gl = c;
        else if (gi == first)
          gf = c;
        else
          ok = false;
      }
      b.gen = ok ? (gf | (gl << 4)) : 0u;
    }
    if (!diag_layers) b.real = b.diag1 = false;
    if (single) {
      const HostGate& g = p->gates[b.members[0]];
      b.cls = g.cls;
      b.targets = g.targets;
      b.controls = g.controls;
      b.count = g.red_count;
      b.real = b.real && g.real && g.cls == OP_DENSE;
      if (g.kind == TQ_G_FIXED && g.cls == OP_DENSE && g.targets.size() == 1 && g.red_count == 4) {
        const zc* X = p->fixed.data() + g.red_off;
        b.is_x = X[0] == zc(0, 0) && X[1] == zc(1, 0) && X[2] == zc(1, 0) && X[3] == zc(0, 0);
      }
      mb.mode = g.kind == TQ_G_FIXED ? MB_FIXED : MB_NATIVE;
      mb.diag = g.cls == OP_DIAG;
      mb.dim = 1 << (int)g.targets.size();
    } else {
      b.cls = OP_DENSE;
      b.targets = b.qubits;
      b.controls.clear();
      mb.dim = 1 << (int)b.qubits.size();
      b.count = mb.dim * mb.dim;
      // a run of fixed gates folds into one constant matrix: nothing to compute per parameter set
      mb.mode = (n_instr == 1 && p->minstrs.back().kind == TQ_G_FIXED) ? MB_FIXED : MB_FUSED;
    }
    mb.count = b.count;
    b.mat_block = (int)p->mblocks.size();
    p->mblocks.push_back(mb);
    p->blocks.push_back(b);
  }

  // ---- measurements ------------------------------------------------------------
  p->meas.assign(meas, meas + n_meas);
  p->dmeas.resize(n_meas);
  for (int mi = 0; mi < n_meas; ++mi) {
    const tq_meas_desc& ms = meas[mi];
    DevMeas& d = p->dmeas[mi];
    memset(&d, 0, sizeof(d));
    d.kind = ms.kind;
    d.flags = ms.flags;
    d.nq = ms.nq;
    d.slot_base = -1;
    d.out_off = p->out_reals;
    TQ_REQUIRE(ms.nq >= 0 && ms.nq <= n && ms.nq <= TQ_MAX_MEAS_QUBITS, TQ_E_INVALID, "measurement %d: bad nq", mi);
    for (int t = 0; t < ms.nq; ++t)
      TQ_REQUIRE(ms.qubits[t] >= 0 && ms.qubits[t] < n, TQ_E_INVALID, "measurement %d: qubit out of range", mi);
    if (ms.kind == TQ_M_EXPVAL) {
      d.slot_base = p->n_slots++;
      p->out_reals += 1;
      if (ms.flags & TQ_MF_ZSTRING) {
        for (int t = 0; t < ms.nq; ++t) d.zmask ^= 1u << (n - 1 - ms.qubits[t]);
      } else {
        TQ_REQUIRE(ms.nq >= 1 && ms.nq <= TQ_MAX_OBS_QUBITS, TQ_E_UNSUPPORTED,
                   "measurement %d: dense observable on %d qubits unsupported", mi, ms.nq);
        const int D = 1 << ms.nq;
        TQ_REQUIRE(ms.matrix_off >= 0 && ms.matrix_off + (int64_t)D * D <= pool_len, TQ_E_INVALID,
                   "measurement %d: matrix outside pool", mi);
        d.mat_off = (int)p->fixed.size();
        p->fixed.insert(p->fixed.end(), zpool + ms.matrix_off, zpool + ms.matrix_off + D * D);
        for (int t = 0; t < ms.nq; ++t) d.pos[t] = (int8_t)(n - 1 - ms.qubits[t]);
      }
    } else if (ms.kind == TQ_M_PROBS) {
      if (ms.nq == 0 || ms.nq == n) {
        d.nq = 0;  // every qubit kept: torch.sum over no axis (pytorch_backend.py:468-471)
        p->out_reals += (int64_t)1 << n;
      } else {
        std::vector<int> kept(ms.qubits, ms.qubits + ms.nq);
        std::sort(kept.begin(), kept.end());  // torch.sum(dim=other axes) keeps ascending qubit order
        for (int t = 0; t + 1 < ms.nq; ++t)
          TQ_REQUIRE(kept[t] != kept[t + 1], TQ_E_INVALID, "measurement %d: repeated qubit", mi);
        TQ_REQUIRE(ms.nq <= 32, TQ_E_UNSUPPORTED, "measurement %d: too many kept qubits", mi);
        for (int t = 0; t < ms.nq; ++t) d.pos[t] = (int8_t)(n - 1 - kept[t]);
        if (ms.nq <= 6) {
          d.slot_base = p->n_slots;
          p->n_slots += 1 << ms.nq;
        }
        p->out_reals += (int64_t)1 << ms.nq;
      }
    } else if (ms.kind == TQ_M_STATE) {
      p->out_reals += (int64_t)2 << n;
    } else {
      TQ_REQUIRE(false, TQ_E_INVALID, "measurement %d: unknown kind %d", mi, ms.kind);
    }
  }

  // ---- schedule, descriptors, payload streams, chunk tables -----------------------------------------
  const int nb = (int)p->blocks.size();
  std::vector<int> order_f(nb), order_b(nb);
  for (int i = 0; i < nb; ++i) {
    order_f[i] = i;
    order_b[i] = nb - 1 - i;
  }
  std::vector<OpDesc> ops_f, ops_b;
  std::vector<ChunkInfo> chunks_f, chunks_b;
  std::vector<int32_t> slot_pidx;
  for (int dir = 0; dir < 2; ++dir) {
    std::vector<std::vector<int>> sg, sb;
    schedule(n, dir ? m_b : m_f, coalesce, p->blocks, dir ? order_b : order_f, sg, sb);
    std::vector<Sweep>& sweeps = dir ? p->bwd : p->fwd;
    std::vector<int> folded;  // register-group mode: X / CNOT blocks folded into load / store addresses
    std::vector<OpDesc>& ops = dir ? ops_b : ops_f;
    std::vector<ChunkInfo>& chunks = dir ? chunks_b : chunks_f;
    int64_t& stride = dir ? p->stride_b : p->stride_f;
    for (size_t s = 0; s < sg.size(); ++s) {
      Sweep sw;
      sw.bits = sb[s];
      build_geom(n, sb[s], sw.geom);
      TQ_REQUIRE(sw.geom.nl <= 16 && sw.geom.nt <= 32, TQ_E_UNSUPPORTED, "tile geometry too fragmented");
      sw.op_begin = (int)ops.size();
      sw.slot_begin = (int)slot_pidx.size();
      sw.chunk_begin = (int)chunks.size();
      ChunkInfo cur = {0, 0, 0, 0};
      bool open = false;
      if (p->rg) {
        // Register groups: walk the sweep's blocks in order; a block joins the open group when the group's tile bits plus
        // its own stay within four and none of its bits belongs to a block that was passed over (it commutes with every
        // passed-over block then); passed-over blocks start the next groups.
        const std::vector<int>& tb = sb[s];
        const int m_t = (int)tb.size();
        auto local = [&](int q) {
          const int b = n - 1 - q;
          for (int j = 0; j < m_t; ++j)
            if (tb[j] == b) return j;
          return -1;
        };
        // the sweep's blocks as the grouper sees them (rg_next_group, tq_sv_rg.cuh)
        std::vector<RgItem> items(p->blocks.size());
        for (int bi : sg[s]) {
          const HostBlock& hb = p->blocks[bi];
          RgItem& it = items[bi];
          it.nbits = 0;
          for (int q : hb.targets) it.bits[it.nbits++] = local(q);
          for (int q : hb.controls) it.bits[it.nbits++] = local(q);
          for (int k = 0; k < it.nbits; ++k)
            TQ_REQUIRE(it.bits[k] >= 0, TQ_E_INVALID, "tq_plan_create: block outside its sweep's tile");
          it.foldable = hb.is_x && hb.nderiv == 0 && hb.controls.size() <= 1;
          it.pay = block_pay_entries(hb.count, hb.nderiv, dir != 0);
        }
        std::vector<int> rest(sg[s]);
        while (!rest.empty()) {
          std::vector<int> pre, mid, post;
          std::vector<char> inb;
          rg_next_group(items, rest, m_t, pay_cap_entries, pre, mid, post, inb);
          int pay = 0;
          for (int bi : mid) pay += items[bi].pay;
          int used[RG_BITS], nu = 0, reg[RG_BITS];
          for (int b = 0; b < m_t; ++b)
            if (inb[b]) used[nu++] = b;
          rg_pick_bits(m_t, used, nu, reg);
          auto reg_index = [&](int q) {
            const int b = local(q);
            for (int i = 0; i < RG_BITS; ++i)
              if (reg[i] == b) return i;
            return -1;
          };
          auto fold = [&](const std::vector<int>& xs, bool forward_order) {
            std::vector<int> tg, ct;
            for (int bi : xs) {
              const HostBlock& hb = p->blocks[bi];
              tg.push_back(reg_index(hb.targets[0]));
              ct.push_back(hb.controls.empty() ? -1 : reg_index(hb.controls[0]));
            }
            return rg_affine_map(tg.data(), ct.data(), (int)xs.size(), forward_order);
          };
          // the header and its sub-ops travel in one prefetch chunk
          if (open && (cur.op_count + 1 + mid.size() > (size_t)CHUNK_OPS || (int)cur.pay_count + pay > pay_cap_entries)) {
            chunks.push_back(cur);
            open = false;
          }
          if (!open) {
            cur.op_begin = (uint32_t)ops.size();
            cur.op_count = 0;
            cur.pay_begin = (uint32_t)stride;
            cur.pay_count = 0;
            open = true;
          }
          OpDesc h;
          rg_pack_header(reg, (int)mid.size(), fold(pre, false), fold(post, true), h);
          ops.push_back(h);
          cur.op_count += 1;
          // folded blocks still own a (never read) slot in the payload streams, placed behind the last sweep's
          // payload: k_materialize writes every block
          for (const std::vector<int>* xs : {&pre, &post})
            for (int bi : *xs) {
              folded.push_back(bi);
              p->n_rg_folded[dir] += 1;
            }
          for (int bi : mid) {
            HostBlock& hb = p->blocks[bi];
            int treg[4], creg[4], nt = 0, nc = 0;
            for (int q : hb.targets) treg[nt++] = reg_index(q);
            for (int q : hb.controls) creg[nc++] = reg_index(q);
            TQ_REQUIRE(hb.cls == OP_DIAG || nt == 1, TQ_E_INVALID, "tq_plan_create: dense block with %d targets in a register group", nt);
            OpDesc d;
            rg_make_sub(hb.cls, nt, treg, nc, creg, hb.is_x, hb.count, hb.nderiv, d, hb.gen);
            TQ_REQUIRE(d.path != RG_GEN || hb.nderiv == 0, TQ_E_UNSUPPORTED, "tq_plan_create: trainable multi-target diagonal");
            const int pe = block_pay_entries(hb.count, hb.nderiv, dir != 0);
            (dir ? p->mblocks[hb.mat_block].off_b : p->mblocks[hb.mat_block].off_f) = (int32_t)stride;
            d.pay_off = (uint32_t)stride;
            if (dir) {
              d.dslot = (uint32_t)((int)slot_pidx.size() - sw.slot_begin);
              for (int k = 0; k < hb.nderiv; ++k) slot_pidx.push_back(hb.pidx[k]);
            }
            ops.push_back(d);
            cur.op_count += 1;
            cur.pay_count += (uint32_t)pe;
            stride += pe;
          }
          const std::vector<int>& group = mid;
          p->n_ops[dir] += (int)group.size();
          p->n_rg[dir] += 1;
        }
      } else {
        // Emission order: one-qubit diagonal blocks are held back and merged into diagonal-layer ops.  A held-back
        // block commutes with everything emitted before the layer is flushed (no shared qubit), and the layer is
        // flushed before the first block that touches one of its qubits.
        std::vector<std::vector<int>> items;
        {
          std::vector<int> group;
          std::vector<char> group_q(n, 0);
          auto flush = [&]() {
            if (group.empty()) return;
            items.push_back(group);
            group.clear();
            std::fill(group_q.begin(), group_q.end(), 0);
          };
          for (int bi : sg[s]) {
            const HostBlock& hb = p->blocks[bi];
            if (hb.diag1 && hb.cls == OP_DIAG) {
              const int q = hb.targets[0];
              if (group_q[q] || (int)group.size() == DL_MAX) flush();
              group.push_back(bi);
              group_q[q] = 1;
            } else {
              bool touches = false;
              for (int q : hb.qubits) touches = touches || group_q[q];
              if (touches) flush();
              items.push_back({bi});
            }
          }
          flush();
        }
        for (const std::vector<int>& item : items) {
          OpDesc d;
          int pe;
          const int64_t pay0 = stride;
          if (item.size() == 1) {
            HostBlock& hb = p->blocks[item[0]];
            make_desc(hb, n, sb[s], c64, d);
            pe = block_pay_entries(hb.count, hb.nderiv, dir != 0);
            (dir ? p->mblocks[hb.mat_block].off_b : p->mblocks[hb.mat_block].off_f) = (int32_t)stride;
            if (dir) {
              d.dslot = (uint32_t)((int)slot_pidx.size() - sw.slot_begin);
              for (int k = 0; k < hb.nderiv; ++k) slot_pidx.push_back(hb.pidx[k]);
            }
          } else {  // diagonal layer: the members' payloads back to back (2 entries each forward, 4 backward)
            memset(&d, 0, sizeof(d));
            d.path = P_DL;
            const int per = dir ? 4 : 2;
            pe = per * (int)item.size();
            uint32_t mask = 0;
            int nd = 0;
            uint8_t pos[DL_MAX];
            memset(pos, 0, sizeof(pos));
            if (dir) d.dslot = (uint32_t)((int)slot_pidx.size() - sw.slot_begin);
            for (size_t k = 0; k < item.size(); ++k) {
              HostBlock& hb = p->blocks[item[k]];
              const int bit = n - 1 - hb.targets[0];
              int lp = -1;
              for (size_t j = 0; j < sb[s].size(); ++j)
                if (sb[s][j] == bit) lp = (int)j;
              TQ_REQUIRE(lp >= 0, TQ_E_INVALID, "tq_plan_create: diagonal-layer member outside its sweep's tile");
              pos[k] = (uint8_t)lp;
              (dir ? p->mblocks[hb.mat_block].off_b : p->mblocks[hb.mat_block].off_f) = (int32_t)(stride + per * (int64_t)k);
              TQ_REQUIRE(hb.nderiv <= 1, TQ_E_UNSUPPORTED, "tq_plan_create: a diagonal gate with %d parameters", hb.nderiv);
              if (hb.nderiv == 1) {
                mask |= 1u << k;
                ++nd;
                if (dir) slot_pidx.push_back(hb.pidx[0]);
              }
            }
            d.nderiv = (uint8_t)nd;
            d.count = (uint32_t)item.size() | (mask << 8);
            for (int j = 0; j < 4; ++j) {
              d.ins[j] = pos[j];
              d.tpos[j] = pos[4 + j];
              d.cmask |= (uint32_t)pos[8 + j] << (8 * j);
              d.pad |= (uint32_t)pos[12 + j] << (8 * j);
            }
          }
          TQ_REQUIRE(pe <= pay_cap_entries, TQ_E_UNSUPPORTED, "block payload exceeds the prefetch buffer");
          d.pay_off = (uint32_t)pay0;
          if (open && (cur.op_count >= (uint32_t)CHUNK_OPS || (int)cur.pay_count + pe > pay_cap_entries)) {
            chunks.push_back(cur);
            open = false;
          }
          if (!open) {
            cur.op_begin = (uint32_t)ops.size();
            cur.op_count = 0;
            cur.pay_begin = (uint32_t)stride;
            cur.pay_count = 0;
            open = true;
          }
          cur.op_count += 1;
          cur.pay_count += (uint32_t)pe;
          stride += pe;
          ops.push_back(d);
          p->n_ops[dir] += 1;
          if (d.path == P_DL) {
            p->n_dl[dir] += 1;
            p->n_dl_members[dir] += (int)item.size();
          }
          if (d.path == P_R1S || d.path == P_R2S || d.path == P_R1V || d.path == P_R1P || d.path == P_R2V) p->n_real[dir] += 1;
        }
      }
      if (open) chunks.push_back(cur);
      sw.op_end = (int)ops.size();
      sw.n_gates = (int)sg[s].size();
      sw.n_dslots = dir ? (int)slot_pidx.size() - sw.slot_begin : 0;
      sw.n_chunks = (int)chunks.size() - sw.chunk_begin;
      sweeps.push_back(sw);
    }
    for (int bi : folded) {
      HostBlock& hb = p->blocks[bi];
      (dir ? p->mblocks[hb.mat_block].off_b : p->mblocks[hb.mat_block].off_f) = (int32_t)stride;
      stride += block_pay_entries(hb.count, hb.nderiv, dir != 0);
    }
    stride = (stride + 31) & ~(int64_t)31;  // keep every set's stream 256-byte aligned
  }
  TQ_REQUIRE(p->stride_f < ((int64_t)1 << 31) && p->stride_b < ((int64_t)1 << 31), TQ_E_UNSUPPORTED, "circuit too long");

  // ---- upload ---------------------------------------------------------------------
  int rc;
  if ((rc = upload_complex(p->fixed.data(), p->fixed.size(), dtype, &p->d_fixed))) return rc;
  if (init_state) {
    if ((rc = upload_complex(reinterpret_cast<const zc*>(init_state), (size_t)1 << n, dtype, &p->d_init))) return rc;
  }
  if ((rc = upload(ops_f, &p->d_ops_f))) return rc;
  if ((rc = upload(ops_b, &p->d_ops_b))) return rc;
  if ((rc = upload(chunks_f, &p->d_chunks_f))) return rc;
  if ((rc = upload(chunks_b, &p->d_chunks_b))) return rc;
  if ((rc = upload(p->mblocks, &p->d_mblocks))) return rc;
  if ((rc = upload(p->minstrs, &p->d_minstrs))) return rc;
  if ((rc = upload(p->dmeas, &p->d_meas))) return rc;
  if ((rc = upload(slot_pidx, &p->d_slot_pidx))) return rc;
  if ((rc = upload(p->gate_t, &p->d_gate_t))) return rc;
  *out = P.release();
  return TQ_OK;
}

int32_t tq_plan_num_qubits(const tq_plan* p) { return p ? p->n : -1; }
int32_t tq_plan_num_params(const tq_plan* p) { return p ? p->n_params : -1; }
int32_t tq_plan_num_sweeps(const tq_plan* p, int32_t backward) {
  return p ? (int32_t)(backward ? p->bwd.size() : p->fwd.size()) : -1;
}
int32_t tq_plan_sweep_bits(const tq_plan* p, int32_t backward, int32_t s, int32_t* bits, int32_t cap) {
  if (!p) return -1;
  const std::vector<Sweep>& v = backward ? p->bwd : p->fwd;
  if (s < 0 || s >= (int)v.size()) return -1;
  for (int i = 0; i < (int)v[s].bits.size() && i < cap; ++i) bits[i] = v[s].bits[i];
  return (int32_t)v[s].bits.size();
}
int32_t tq_plan_sweep_num_gates(const tq_plan* p, int32_t backward, int32_t s) {
  if (!p) return -1;
  const std::vector<Sweep>& v = backward ? p->bwd : p->fwd;
  if (s < 0 || s >= (int)v.size()) return -1;
  return v[s].n_gates;
}
int32_t tq_plan_num_blocks(const tq_plan* p) { return p ? (int32_t)p->blocks.size() : -1; }
/* what = 0: ops emitted into the sweeps, 1: diagonal-layer ops, 2: their members, 3: ops on the real paths,
 * 4: register groups (structure = 2; 0 when the plan fell back to the default sweeps), 5: X / CNOT blocks folded into
 * the groups' load / store addresses */
int32_t tq_plan_op_stats(const tq_plan* p, int32_t backward, int32_t what) {
  if (!p || what < 0 || what > 5) return -1;
  const int d = backward ? 1 : 0;
  if (what == 4) return p->n_rg[d];
  if (what == 5) return p->n_rg_folded[d];
  return what == 0 ? p->n_ops[d] : what == 1 ? p->n_dl[d] : what == 2 ? p->n_dl_members[d] : p->n_real[d];
}
int64_t tq_plan_out_reals(const tq_plan* p) { return p ? p->out_reals : -1; }

int64_t tq_plan_hbm_bytes(const tq_plan* p, int32_t backward) {
  if (!p) return -1;
  const int64_t cs = (int64_t)csize(p->dtype);
  const int64_t sv = ((int64_t)1 << p->n) * cs;
  const int64_t io = (int64_t)rsize(p->dtype) * (p->n_params + p->out_reals);
  if (!backward) {
    // payload stream written once and read once; first sweep synthesises |0..0> (no read);
    // every sweep writes; the measurement pass reads once
    const int64_t pay = 2 * p->stride_f * cs;
    if (p->fwd_full) return io + pay;  // (+ one write of psi when a backward follows, counted there)
    return io + pay + sv * (2 * (int64_t)p->fwd.size());
  }
  const int64_t pay = 2 * p->stride_b * cs;
  if (p->bwd_full) return io + pay + (int64_t)rsize(p->dtype) * p->n_params + 2 * sv;  // psi stored + re-read
  return io + pay + sv * (2 + 4 * (int64_t)p->bwd.size());
}

/* Algorithmic real flops of one evaluation on the fused-block schedule (complex MAC = 8 flops, complex multiply = 6):
 * forward: a dense block on t target qubits costs 8 * 2^t per amplitude it touches (2^(n - controls) of them), a
 * diagonal block 6 (a member of a diagonal layer included), a REAL dense block 4 * 2^t; backward (adjoint method): psi <- G^dagger psi, lambda <- G^dagger lambda and the accumulation of
 * W = psi (x) conj(lambda) cost one such product each, i.e. 3x the forward count.  Measurement passes not counted. */
double tq_plan_flops(const tq_plan* p, int32_t backward) {
  if (!p) return -1.0;
  double fl = 0.0;
  for (const HostBlock& b : p->blocks) {
    if (!b.alive) continue;
    const double amps = ldexp(1.0, p->n - (int)b.controls.size());
    // complex MAC = 8 flops; a real matrix entry times a complex amplitude, accumulated = 4
    const double per = b.cls == OP_DIAG ? 6.0 : (b.real ? 4.0 : 8.0) * ldexp(1.0, (int)b.targets.size());
    if (p->rg && b.is_x) continue;                  // a register swap / an address relabelling
    const double fwd = amps * (p->rg && b.cls == OP_DIAG && b.targets.size() == 1 ? 16.0 : per);  // diagonal applied as a 2x2 there
    // adjoint: psi <- G^dagger psi and lambda <- G^dagger lambda cost one forward product each; the gradient terms one
    // more (W = psi (x) conj(lambda)), or, in a register group whose slots have generator gradients, 4 multiply-adds
    // per amplitude pair and slot (a quarter of a 2x2 product each)
    double factor = 3.0;
    if (p->rg && b.gen) factor = 2.0 + 0.25 * b.nderiv;
    if (p->rg && b.nderiv == 0) factor = 2.0;
    fl += backward ? factor * fwd : fwd;
  }
  return fl;
}

int64_t tq_plan_launches(const tq_plan* p, int32_t backward) {
  if (!p) return -1;
  if (!backward) return 1 + (p->fwd_full ? 1 : (int64_t)p->fwd.size() + 1);
  return 1 + (p->bwd_full ? 1 : (int64_t)p->bwd.size() + 1);
}

int tq_tn_param_grads(const tq_plan* p, const void* params, int64_t batch, const void* arena, int64_t set_stride,
                      const int32_t* off_g, const int32_t* off_a, void* grad_params, void* stream) {
  TQ_REQUIRE(p && arena && off_g && off_a && grad_params && batch > 0, TQ_E_INVALID,
             "tq_tn_param_grads: null argument");
  TQ_REQUIRE(params || p->n_params == 0, TQ_E_INVALID, "tq_tn_param_grads: params is null");
  TQ_REQUIRE_DEVICE(p, "tq_tn_param_grads");
  const int ng = (int)p->gate_t.size();
  if (ng == 0 || p->n_params == 0) return TQ_OK;
  const int64_t total = batch * ng;
  const int threads = 128;
  const int64_t blocks = (total + threads - 1) / threads;
  if (p->dtype == TQ_C64)
    k_gate_tensor_grads<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const float*)params, p->n_params, batch, p->d_gate_t, ng, (const cx<float>*)arena, set_stride, off_g, off_a,
        (float*)grad_params);
  else
    k_gate_tensor_grads<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const double*)params, p->n_params, batch, p->d_gate_t, ng, (const cx<double>*)arena, set_stride, off_g,
        off_a, (double*)grad_params);
  TQ_CUDA_OK(cudaGetLastError());
  return TQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// execution
// ---------------------------------------------------------------------------
namespace tq {

struct WsLayout {
  size_t sf = 0, sb = 0, psi = 0, lam = 0, total = 0;
  bool has_psi = false, has_lam = false;
};

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static WsLayout ws_layout(const tq_plan* p, int64_t B, int with_backward) {
  WsLayout w;
  const size_t cs = csize(p->dtype);
  size_t off = 0;
  w.sf = off;
  off += align_up((size_t)B * p->stride_f * cs);
  w.sb = off;
  if (with_backward) off += align_up((size_t)B * p->stride_b * cs);
  w.has_psi = !p->fwd_full || with_backward;  // the adjoint pass starts from the stored final state
  w.has_lam = with_backward && !p->bwd_full;
  w.psi = off;
  if (w.has_psi) off += align_up(((size_t)B << p->n) * cs);
  w.lam = off;
  if (w.has_lam) off += align_up(((size_t)B << p->n) * cs);
  w.total = off + 256;
  return w;
}

template <typename K>
static int prep_kernel(K kernel, size_t smem) {
  TQ_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TQ_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
  return TQ_OK;
}

template <typename R>
static int run_materialize(const tq_plan* p, const void* params, int64_t B, char* ws, const WsLayout& L,
                           int with_deriv, cudaStream_t st) {
  const int nbk = (int)p->mblocks.size();
  if (nbk == 0) return TQ_OK;
  int64_t total = B * nbk * MAT_LANES;
  int threads = 128;
  int64_t blocks = (total + threads - 1) / threads;
  k_materialize<R><<<(unsigned)blocks, threads, 0, st>>>(
      (const R*)params, p->n_params, B, p->d_mblocks, nbk, p->d_minstrs, (const cx<R>*)p->d_fixed,
      (cx<R>*)(ws + L.sf), p->stride_f, (cx<R>*)(ws + L.sb), p->stride_b, with_deriv);
  TQ_CUDA_OK(cudaGetLastError());
  return TQ_OK;
}

static StreamRef stream_ref(const OpDesc* ops, const ChunkInfo* chunks, const Sweep& sw) {
  StreamRef r;
  r.ops = ops;
  r.chunks = chunks + sw.chunk_begin;
  r.n_chunks = sw.n_chunks;
  return r;
}

template <typename R>
static int forward_impl(const tq_plan* p, const void* params, int64_t B, void* out, void* workspace, size_t ws_bytes,
                        int with_backward, cudaStream_t st) {
  WsLayout L = ws_layout(p, B, with_backward);
  TQ_REQUIRE(ws_bytes >= L.total, TQ_E_WORKSPACE, "tq_forward: workspace %zu < required %zu", ws_bytes, L.total);
  char* ws = (char*)workspace;
  int rc = run_materialize<R>(p, params, B, ws, L, 0, st);
  if (rc) return rc;
  const int n = p->n;
  cx<R>* psi = L.has_psi ? (cx<R>*)(ws + L.psi) : nullptr;
  FwdArgs<R> a;
  memset(&a, 0, sizeof(a));
  a.psi = psi;
  a.init_state = (const cx<R>*)p->d_init;
  a.stream = (const cx<R>*)(ws + L.sf);
  a.stride = p->stride_f;
  a.fixed = (const cx<R>*)p->d_fixed;
  a.meas = p->d_meas;
  a.out = (R*)out;
  a.out_reals = p->out_reals;
  a.n_meas = (int)p->dmeas.size();
  a.n_slots = p->n_slots;
  TQ_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)B * p->out_reals * sizeof(R), st));
  if (p->fwd_full) {
    const Sweep& sw = p->fwd[0];
    a.st = stream_ref(p->d_ops_f, p->d_chunks_f, sw);
    a.geom = sw.geom;
    a.tiles_log2 = 0;
    a.flags = SW_INIT | SW_MEASURE | (L.has_psi ? SW_STORE : 0);
    size_t smem = ((size_t)sizeof(cx<R>) << n) + RING_BYTES + sizeof(R) * p->n_slots;
    if constexpr (sizeof(R) == 4) {
      if (p->rg) {
        if ((rc = prep_kernel(k_rg_fwd, smem))) return rc;
        k_rg_fwd<<<(unsigned)B, p->threads_f, smem, st>>>(a);
        TQ_CUDA_OK(cudaGetLastError());
        return TQ_OK;
      }
    }
    if (p->structure) {
      if ((rc = prep_kernel(k_sweep_fwd<R, true>, smem))) return rc;
      k_sweep_fwd<R, true><<<(unsigned)B, p->threads_f, smem, st>>>(a);
    } else {
      if ((rc = prep_kernel(k_sweep_fwd<R, false>, smem))) return rc;
      k_sweep_fwd<R, false><<<(unsigned)B, p->threads_f, smem, st>>>(a);
    }
    TQ_CUDA_OK(cudaGetLastError());
    return TQ_OK;
  }
  for (size_t s = 0; s < p->fwd.size(); ++s) {
    const Sweep& sw = p->fwd[s];
    a.st = stream_ref(p->d_ops_f, p->d_chunks_f, sw);
    a.geom = sw.geom;
    a.tiles_log2 = n - sw.geom.m;
    a.flags = SW_STORE | (s == 0 ? SW_INIT : 0);
    size_t smem = ((size_t)sizeof(cx<R>) << sw.geom.m) + RING_BYTES;
    int64_t blocks = B << a.tiles_log2;
    TQ_REQUIRE(blocks < ((int64_t)1 << 31), TQ_E_UNSUPPORTED, "tq_forward: batch too large for one launch");
    if constexpr (sizeof(R) == 4) {
      if (p->rg) {
        if ((rc = prep_kernel(k_rg_fwd, smem))) return rc;
        k_rg_fwd<<<(unsigned)blocks, p->threads_f, smem, st>>>(a);
        TQ_CUDA_OK(cudaGetLastError());
        continue;
      }
    }
    if ((rc = p->structure ? prep_kernel(k_sweep_fwd<R, true>, smem) : prep_kernel(k_sweep_fwd<R, false>, smem))) return rc;
    if (p->structure)
      k_sweep_fwd<R, true><<<(unsigned)blocks, p->threads_f, smem, st>>>(a);
    else
      k_sweep_fwd<R, false><<<(unsigned)blocks, p->threads_f, smem, st>>>(a);
    TQ_CUDA_OK(cudaGetLastError());
  }
  MeasArgs<R> ma;
  memset(&ma, 0, sizeof(ma));
  ma.psi = psi;
  ma.fixed = (const cx<R>*)p->d_fixed;
  ma.meas = p->d_meas;
  ma.out = (R*)out;
  ma.out_reals = p->out_reals;
  ma.n_meas = (int)p->dmeas.size();
  ma.n_slots = p->n_slots;
  ma.n = n;
  ma.chunk_log2 = std::min(n, 12);
  int64_t blocks = B << (n - ma.chunk_log2);
  TQ_REQUIRE(blocks < ((int64_t)1 << 31), TQ_E_UNSUPPORTED, "tq_forward: batch too large for one launch");
  k_measure<R><<<(unsigned)blocks, 256, sizeof(R) * p->n_slots, st>>>(ma);
  TQ_CUDA_OK(cudaGetLastError());
  return TQ_OK;
}

// Register-group adjoint sweep: shared memory for the gradient cells (tq_sv_rg.cuh: rg_grad_cell).  The fewest butterfly
// rounds whose cells still fit beside two resident CTAs per SM (one when the tiles alone exceed half an SM).
static int rg_grad_layout(size_t used, int n_dslots, int threads, size_t* smem) {
  const size_t sm_bytes = 227 * 1024, two = sm_bytes / 2 - 1024, one = sm_bytes - 2048;
  const size_t limit = used + 32 * (size_t)n_dslots <= two ? two : one;
  static const bool force_direct = getenv("TQ_RG_GRAD_DIRECT") != nullptr;  // tests: exercise the no-cells path
  for (int r = 0; r <= 5 && !force_direct; ++r) {
    const size_t need = used + sizeof(float) * (size_t)n_dslots * (size_t)(threads >> r);
    if (need <= limit) {
      *smem = need;
      return r;
    }
  }
  // thousands of trainable slots in one sweep: no cells, every warp adds its sums to the gradient in global memory
  *smem = used;
  return RG_GRAD_DIRECT;
}

template <typename R>
static int backward_impl(const tq_plan* p, const void* params, int64_t B, const void* grad_out, void* grad_params,
                         void* workspace, size_t ws_bytes, cudaStream_t st) {
  WsLayout L = ws_layout(p, B, 1);
  TQ_REQUIRE(ws_bytes >= L.total, TQ_E_WORKSPACE, "tq_backward: workspace %zu < required %zu", ws_bytes, L.total);
  char* ws = (char*)workspace;
  int rc = run_materialize<R>(p, params, B, ws, L, 1, st);
  if (rc) return rc;
  const int n = p->n;
  TQ_CUDA_OK(cudaMemsetAsync(grad_params, 0, (size_t)B * p->n_params * sizeof(R), st));
  BwdArgs<R> a;
  memset(&a, 0, sizeof(a));
  a.psi = L.has_psi ? (cx<R>*)(ws + L.psi) : nullptr;
  a.lam = L.has_lam ? (cx<R>*)(ws + L.lam) : nullptr;
  a.init_state = (const cx<R>*)p->d_init;
  a.stream_f = (const cx<R>*)(ws + L.sf);
  a.stream_b = (const cx<R>*)(ws + L.sb);
  a.stride_f = p->stride_f;
  a.stride_b = p->stride_b;
  a.fixed = (const cx<R>*)p->d_fixed;
  a.meas = p->d_meas;
  a.dy = (const R*)grad_out;
  a.grad = (R*)grad_params;
  a.out_reals = p->out_reals;
  a.n_meas = (int)p->dmeas.size();
  a.n_params = p->n_params;
  if (p->bwd_full) {
    const Sweep& sb = p->bwd[0];
    const Sweep& sf = p->fwd[0];
    a.st_b = stream_ref(p->d_ops_b, p->d_chunks_b, sb);
    a.st_f = stream_ref(p->d_ops_f, p->d_chunks_f, sf);
    a.geom = sb.geom;
    a.slot_pidx = p->d_slot_pidx + sb.slot_begin;
    a.n_dslots = sb.n_dslots;
    a.flags = SW_FULL;
    a.tiles_log2 = 0;
    size_t smem = ((size_t)2 * sizeof(cx<R>) << n) + RING_BYTES + sizeof(R) * sb.n_dslots;
    if constexpr (sizeof(R) == 4) {
      if (p->rg) {
        TQ_REQUIRE(a.psi, TQ_E_INVALID, "tq_backward: the workspace holds no final state");
        a.grad_rounds = rg_grad_layout(((size_t)2 * sizeof(cx<R>) << n) + RING_BYTES, sb.n_dslots, p->threads_b, &smem);
        TQ_REQUIRE(smem <= 227 * 1024 - 1024, TQ_E_UNSUPPORTED, "tq_backward: %d gradient slots exceed shared memory", sb.n_dslots);
        if (a.grad_rounds == RG_GRAD_DIRECT) {
          if ((rc = prep_kernel(k_rg_bwd<true>, smem))) return rc;
          k_rg_bwd<true><<<(unsigned)B, p->threads_b, smem, st>>>(a);
        } else {
          if ((rc = prep_kernel(k_rg_bwd<false>, smem))) return rc;
          k_rg_bwd<false><<<(unsigned)B, p->threads_b, smem, st>>>(a);
        }
        TQ_CUDA_OK(cudaGetLastError());
        return TQ_OK;
      }
    }
    if (p->structure) {
      if ((rc = prep_kernel(k_sweep_bwd<R, true>, smem))) return rc;
      k_sweep_bwd<R, true><<<(unsigned)B, p->threads_b, smem, st>>>(a);
    } else {
      if ((rc = prep_kernel(k_sweep_bwd<R, false>, smem))) return rc;
      k_sweep_bwd<R, false><<<(unsigned)B, p->threads_b, smem, st>>>(a);
    }
    TQ_CUDA_OK(cudaGetLastError());
    return TQ_OK;
  }
  SeedArgs<R> sa;
  memset(&sa, 0, sizeof(sa));
  sa.psi = a.psi;
  sa.lam = a.lam;
  sa.fixed = a.fixed;
  sa.meas = p->d_meas;
  sa.dy = (const R*)grad_out;
  sa.out_reals = p->out_reals;
  sa.total = B << n;
  sa.n_meas = a.n_meas;
  sa.n = n;
  int64_t sblocks = std::min<int64_t>((sa.total + 255) / 256, 148 * 32);
  k_seed<R><<<(unsigned)sblocks, 256, 0, st>>>(sa);
  TQ_CUDA_OK(cudaGetLastError());
  for (size_t s = 0; s < p->bwd.size(); ++s) {
    const Sweep& sb = p->bwd[s];
    a.st_b = stream_ref(p->d_ops_b, p->d_chunks_b, sb);
    a.geom = sb.geom;
    a.slot_pidx = p->d_slot_pidx + sb.slot_begin;
    a.n_dslots = sb.n_dslots;
    a.flags = SW_STORE;
    a.tiles_log2 = n - sb.geom.m;
    size_t smem = ((size_t)2 * sizeof(cx<R>) << sb.geom.m) + RING_BYTES + sizeof(R) * sb.n_dslots;
    int64_t blocks = B << a.tiles_log2;
    TQ_REQUIRE(blocks < ((int64_t)1 << 31), TQ_E_UNSUPPORTED, "tq_backward: batch too large for one launch");
    if constexpr (sizeof(R) == 4) {
      if (p->rg) {
        a.grad_rounds = rg_grad_layout(((size_t)2 * sizeof(cx<R>) << sb.geom.m) + RING_BYTES, sb.n_dslots, p->threads_b, &smem);
        TQ_REQUIRE(smem <= 227 * 1024 - 1024, TQ_E_UNSUPPORTED, "tq_backward: %d gradient slots exceed shared memory", sb.n_dslots);
        if (a.grad_rounds == RG_GRAD_DIRECT) {
          if ((rc = prep_kernel(k_rg_bwd<true>, smem))) return rc;
          k_rg_bwd<true><<<(unsigned)blocks, p->threads_b, smem, st>>>(a);
        } else {
          if ((rc = prep_kernel(k_rg_bwd<false>, smem))) return rc;
          k_rg_bwd<false><<<(unsigned)blocks, p->threads_b, smem, st>>>(a);
        }
        TQ_CUDA_OK(cudaGetLastError());
        continue;
      }
    }
    if ((rc = p->structure ? prep_kernel(k_sweep_bwd<R, true>, smem) : prep_kernel(k_sweep_bwd<R, false>, smem))) return rc;
    if (p->structure)
      k_sweep_bwd<R, true><<<(unsigned)blocks, p->threads_b, smem, st>>>(a);
    else
      k_sweep_bwd<R, false><<<(unsigned)blocks, p->threads_b, smem, st>>>(a);
    TQ_CUDA_OK(cudaGetLastError());
  }
  return TQ_OK;
}

}  // namespace tq

extern "C" {

size_t tq_workspace_bytes(const tq_plan* p, int64_t batch, int32_t with_backward) {
  if (!p || batch <= 0) return 0;
  return ws_layout(p, batch, with_backward).total;
}

int tq_forward(const tq_plan* p, const void* params, int64_t batch, void* out, void* workspace, size_t ws_bytes,
               int32_t with_backward, void* stream) {
  TQ_REQUIRE(p && out && workspace && batch > 0, TQ_E_INVALID, "tq_forward: null argument or empty batch");
  TQ_REQUIRE(params || p->n_params == 0, TQ_E_INVALID, "tq_forward: params is null");
  TQ_REQUIRE(p->sv_ok, TQ_E_UNSUPPORTED, "tq_forward: %d qubits exceed the state-vector limit of 30; use the tensor-network entry points", p->n);
  TQ_REQUIRE_DEVICE(p, "tq_forward");
  if (p->dtype == TQ_C64)
    return forward_impl<float>(p, params, batch, out, workspace, ws_bytes, with_backward, (cudaStream_t)stream);
  return forward_impl<double>(p, params, batch, out, workspace, ws_bytes, with_backward, (cudaStream_t)stream);
}

int tq_backward(const tq_plan* p, const void* params, int64_t batch, const void* grad_out, void* grad_params,
                void* workspace, size_t ws_bytes, void* stream) {
  TQ_REQUIRE(p && grad_out && grad_params && workspace && batch > 0, TQ_E_INVALID,
             "tq_backward: null argument or empty batch");
  TQ_REQUIRE(p->n_params > 0, TQ_E_INVALID, "tq_backward: circuit has no parameters");
  TQ_REQUIRE(p->sv_ok, TQ_E_UNSUPPORTED, "tq_backward: %d qubits exceed the state-vector limit of 30", p->n);
  TQ_REQUIRE_DEVICE(p, "tq_backward");
  if (p->dtype == TQ_C64)
    return backward_impl<float>(p, params, batch, grad_out, grad_params, workspace, ws_bytes, (cudaStream_t)stream);
  return backward_impl<double>(p, params, batch, grad_out, grad_params, workspace, ws_bytes, (cudaStream_t)stream);
}

void* tq_workspace_state(const tq_plan* p, void* workspace, int64_t batch) {
  if (!p || !workspace) return nullptr;
  WsLayout L = ws_layout(p, batch, 1);
  return L.has_psi ? (char*)workspace + L.psi : nullptr;
}

int tq_execute_host(tq_plan* p, const void* params, int64_t batch, void* out, const void* grad_out,
                    void* grad_params) {
  TQ_REQUIRE(p && out && batch > 0, TQ_E_INVALID, "tq_execute_host: null argument or empty batch");
  TQ_REQUIRE((grad_out == nullptr) == (grad_params == nullptr), TQ_E_INVALID,
             "tq_execute_host: grad_out and grad_params go together");
  TQ_REQUIRE_DEVICE(p, "tq_execute_host");
  const int with_b = grad_out != nullptr;
  const size_t rs = rsize(p->dtype);
  const size_t pb = align_up((size_t)batch * p->n_params * rs);
  const size_t ob = align_up((size_t)batch * p->out_reals * rs);
  const size_t wsb = tq_workspace_bytes(p, batch, with_b);
  const size_t need = 2 * pb + 2 * ob + wsb;
  if (!p->h_stream) TQ_CUDA_OK(cudaStreamCreateWithFlags(&p->h_stream, cudaStreamNonBlocking));
  if (p->h_dev_bytes < need) {
    if (p->h_dev) cudaFree(p->h_dev);
    p->h_dev = nullptr;
    p->h_dev_bytes = 0;
    TQ_CUDA_OK(cudaMalloc(&p->h_dev, need));
    p->h_dev_bytes = need;
  }
  char* d = (char*)p->h_dev;
  char* d_params = d;
  char* d_gparams = d + pb;
  char* d_out = d + 2 * pb;
  char* d_gout = d + 2 * pb + ob;
  char* d_ws = d + 2 * pb + 2 * ob;
  cudaStream_t st = p->h_stream;
  if (p->n_params)
    TQ_CUDA_OK(cudaMemcpyAsync(d_params, params, (size_t)batch * p->n_params * rs, cudaMemcpyHostToDevice, st));
  int rc = tq_forward(p, d_params, batch, d_out, d_ws, wsb, with_b, st);
  if (rc) return rc;
  TQ_CUDA_OK(cudaMemcpyAsync(out, d_out, (size_t)batch * p->out_reals * rs, cudaMemcpyDeviceToHost, st));
  if (with_b) {
    TQ_CUDA_OK(cudaMemcpyAsync(d_gout, grad_out, (size_t)batch * p->out_reals * rs, cudaMemcpyHostToDevice, st));
    rc = tq_backward(p, d_params, batch, d_gout, d_gparams, d_ws, wsb, st);
    if (rc) return rc;
    TQ_CUDA_OK(cudaMemcpyAsync(grad_params, d_gparams, (size_t)batch * p->n_params * rs, cudaMemcpyDeviceToHost, st));
  }
  TQ_CUDA_OK(cudaStreamSynchronize(st));
  return TQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// operand tensors for the tensor-network path
// ---------------------------------------------------------------------------
extern "C" {

int64_t tq_tn_gate_offset(const tq_plan* p, int32_t gate) {
  if (!p || gate < 0 || gate > (int)p->gate_t.size()) return -1;
  return gate == (int)p->gate_t.size() ? p->gate_t_total : p->gate_t[gate].out_off;
}

int tq_tn_operands(const tq_plan* p, const void* params, int64_t batch, void* gate_mats, void* adj_mats,
                   void* stream) {
  TQ_REQUIRE(p && gate_mats && adj_mats && batch > 0, TQ_E_INVALID, "tq_tn_operands: null argument");
  TQ_REQUIRE(params || p->n_params == 0, TQ_E_INVALID, "tq_tn_operands: params is null");
  TQ_REQUIRE_DEVICE(p, "tq_tn_operands");
  const int ng = (int)p->gate_t.size();
  if (ng == 0) return TQ_OK;
  const int64_t total = batch * ng;
  const int threads = 128;
  const int64_t blocks = (total + threads - 1) / threads;
  if (p->dtype == TQ_C64)
    k_gate_tensors<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const float*)params, p->n_params, batch, p->d_gate_t, ng, (const cx<float>*)p->d_fixed,
        (cx<float>*)gate_mats, (cx<float>*)adj_mats, p->gate_t_total);
  else
    k_gate_tensors<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const double*)params, p->n_params, batch, p->d_gate_t, ng, (const cx<double>*)p->d_fixed,
        (cx<double>*)gate_mats, (cx<double>*)adj_mats, p->gate_t_total);
  TQ_CUDA_OK(cudaGetLastError());
  return TQ_OK;
}

}  // extern "C"
