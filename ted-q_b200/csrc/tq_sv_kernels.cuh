// tq_sv_kernels.cuh — device side of the state-vector engine (sm_100a).
//
// Kernels (all hand-written, no library calls):
//   k_materialize  one thread per (parameter set, fused block): gate matrices from theta
//                  (pytorch_backend.py:866-1188), products of fused runs, and d/dtheta of the
//                  fused product for every trainable slot; written as a contiguous per-set
//                  *payload stream* in the exact order the sweeps consume it
//   k_sweep_fwd    a tile of 2^m amplitudes in shared memory; op descriptors and payload are
//                  prefetched chunk-by-chunk with cp.async (LDGSTS) into a 2-deep ring while the
//                  current chunk's gates run; 128-bit shared-memory accesses on the hot paths
//   k_sweep_bwd    adjoint sweep on the (psi, lambda) tile pair
//   k_measure / k_seed   measurements and cotangent seed over a state in HBM
#pragma once
#include "tq_common.h"

namespace tq {

// ---------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------
enum { OP_DENSE = 0, OP_DIAG = 1 };

// execution paths of one op on a shared-memory tile
enum {
  P_D1V = 0,  // dense 1 target, complex64, bit 0 free: two groups per thread, 128-bit accesses
  P_D1P = 1,  // dense 1 target on amplitude bit 0, complex64: the pair is one 128-bit word
  P_D1S = 2,  // dense 1 target, scalar accesses (complex128, or bit 0 is a control)
  P_D2V = 3,  // dense 2 targets, complex64, bit 0 free
  P_D2S = 4,  // dense 2 targets, scalar
  P_G1V = 5,  // diagonal 1 target, complex64, bit 0 not a control
  P_G1S = 6,  // diagonal 1 target, scalar
  P_GEN = 7,  // anything else (3-target dense, multi-target diagonal): correct, not tuned
  P_R1S = 8,  // REAL dense 1 target (RY, X, H, ... and fused products of such): half the multiplies of P_D1*
  P_R2S = 9,  // REAL dense 2 targets ((RY x RY) CNOT ...)
  P_R1V = 11, P_R1P = 12, P_R2V = 13,  // real twins of P_D1V / P_D1P / P_D2V (complex64, 128-bit accesses)
  P_DL = 10   // diagonal LAYER: up to 16 one-qubit diagonal gates (RZ, PhaseShift, S, T, Z) on different tile bits,
              // applied as ONE pass through two phase tables; payload = the members' own payloads, back to back
};
constexpr int DL_MAX = 16;     // members of a diagonal layer
constexpr int DL_LO_BITS = 7;  // tile bits resolved by the low phase table

struct __align__(16) OpDesc {  // 32 bytes
  uint8_t path, k, nins, nderiv;
  uint8_t ins[4];    // ascending bit positions to insert (units of the path: V/P paths count 128-bit words)
  uint8_t tpos[4];   // target bit positions, tpos[0] = most significant bit of the matrix index
  uint32_t cmask;    // control bits (same units as ins)
  uint32_t pay_off;  // payload offset inside the per-set stream (complex entries)
  uint32_t dslot;    // first gradient slot inside the sweep
  uint32_t count;    // entries per matrix in the payload (4, 16, 64 dense; 2, 4, 8 diagonal)
  uint32_t pad;
};
static_assert(sizeof(OpDesc) == 32, "OpDesc must be 32 bytes");

constexpr int CHUNK_OPS = 16;          // ops per prefetch chunk
constexpr int CHUNK_PAY_BYTES = 2048;  // payload bytes per chunk buffer
constexpr int MAX_CHUNKS_SMEM = 128;   // chunk table entries staged in shared memory
constexpr int MAX_BLOCK_DERIV = 8;     // trainable slots per fused block

struct ChunkInfo {  // 16 bytes
  uint32_t op_begin, op_count, pay_begin, pay_count;  // payload in complex entries
};

struct MatInstr {  // one member gate of a block
  int32_t kind, nq, embed, fixed_off;
  int32_t pidx[3];
  int32_t dsel[3];  // derivative slot inside the block, -1 = not trainable
  double pconst[3];
};

enum { MB_FIXED = 0, MB_NATIVE = 1, MB_FUSED = 2 };

struct MatBlock {
  int32_t mode, dim, count, nderiv;
  int32_t instr_begin, instr_end;
  int32_t off_f, off_b;  // payload offsets in the forward / backward streams (complex entries)
  int32_t diag;          // MB_NATIVE: payload is the diagonal (2 entries)
  int32_t pad[3];
};

struct DevMeas {
  int32_t kind, flags, nq;
  int32_t slot_base;
  int64_t out_off;
  uint32_t zmask;
  int32_t mat_off;
  int8_t pos[32];
};

struct Geom {
  int32_t m, n;
  int32_t nl;
  int8_t lsrc[16], llen[16], ldst[16];
  int32_t nt;
  int8_t tsrc[32], tlen[32], tdst[32];
};

__device__ __forceinline__ uint32_t dep_local(const Geom& g, uint32_t l) {
  uint32_t r = 0;
  for (int i = 0; i < g.nl; ++i) r |= ((l >> g.lsrc[i]) & ((1u << g.llen[i]) - 1u)) << g.ldst[i];
  return r;
}
__device__ __forceinline__ uint32_t dep_tile(const Geom& g, uint32_t t) {
  uint32_t r = 0;
  for (int i = 0; i < g.nt; ++i) r |= ((t >> g.tsrc[i]) & ((1u << g.tlen[i]) - 1u)) << g.tdst[i];
  return r;
}

enum { SW_INIT = 1, SW_STORE = 2, SW_MEASURE = 4, SW_FULL = 8 };

// ---------------------------------------------------------------------------
// k_materialize
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincos_(double a, double* s, double* c) { sincos(a, s, c); }

// 2x2 target block M (row-major) and d/dp_i of it for a parametrised kind.
template <typename R>
__device__ void param_gate(int kind, const R* p, cx<R>* M, cx<R> (*D)[4]) {
  const R h = (R)0.5;
  const cx<R> z = mk<R>(0, 0);
  R s, c;
  switch (kind) {
    case TQ_G_RX:
    case TQ_G_CRX:  // [[c, -i s], [-i s, c]]
      sincos_(p[0] * h, &s, &c);
      M[0] = mk<R>(c, 0); M[1] = mk<R>(0, -s); M[2] = mk<R>(0, -s); M[3] = mk<R>(c, 0);
      D[0][0] = mk<R>(-h * s, 0); D[0][1] = mk<R>(0, -h * c); D[0][2] = mk<R>(0, -h * c); D[0][3] = mk<R>(-h * s, 0);
      break;
    case TQ_G_RY:
    case TQ_G_CRY:  // [[c, -s], [s, c]]
      sincos_(p[0] * h, &s, &c);
      M[0] = mk<R>(c, 0); M[1] = mk<R>(-s, 0); M[2] = mk<R>(s, 0); M[3] = mk<R>(c, 0);
      D[0][0] = mk<R>(-h * s, 0); D[0][1] = mk<R>(-h * c, 0); D[0][2] = mk<R>(h * c, 0); D[0][3] = mk<R>(-h * s, 0);
      break;
    case TQ_G_RZ:
    case TQ_G_CRZ:  // diag(e^{-i t/2}, e^{+i t/2})
      sincos_(p[0] * h, &s, &c);
      M[0] = mk<R>(c, -s); M[1] = z; M[2] = z; M[3] = mk<R>(c, s);
      D[0][0] = mk<R>(-h * s, -h * c); D[0][1] = z; D[0][2] = z; D[0][3] = mk<R>(-h * s, h * c);
      break;
    case TQ_G_PHASESHIFT:
    case TQ_G_CPHASE:  // diag(1, e^{i phi})
      sincos_(p[0], &s, &c);
      M[0] = mk<R>(1, 0); M[1] = z; M[2] = z; M[3] = mk<R>(c, s);
      D[0][0] = z; D[0][1] = z; D[0][2] = z; D[0][3] = mk<R>(-s, c);
      break;
    case TQ_G_ROT: {
      // [[e^{-i(a+w)/2} c, -e^{i(a-w)/2} s], [e^{-i(a-w)/2} s, e^{i(a+w)/2} c]],  c = cos(b/2)
      R sp, cp, sm, cm;
      sincos_(p[1] * h, &s, &c);
      sincos_((p[0] + p[2]) * h, &sp, &cp);
      sincos_((p[0] - p[2]) * h, &sm, &cm);
      M[0] = mk<R>(cp * c, -sp * c);
      M[1] = mk<R>(-cm * s, -sm * s);
      M[2] = mk<R>(cm * s, -sm * s);
      M[3] = mk<R>(cp * c, sp * c);
      // (x, y) * (+i/2) = (-y/2, x/2);  (x, y) * (-i/2) = (y/2, -x/2)
      D[0][0] = mk<R>(h * M[0].y, -h * M[0].x);   // d/da: -i/2, +i/2, -i/2, +i/2
      D[0][1] = mk<R>(-h * M[1].y, h * M[1].x);
      D[0][2] = mk<R>(h * M[2].y, -h * M[2].x);
      D[0][3] = mk<R>(-h * M[3].y, h * M[3].x);
      D[1][0] = mk<R>(-h * cp * s, h * sp * s);   // d/db
      D[1][1] = mk<R>(-h * cm * c, -h * sm * c);
      D[1][2] = mk<R>(h * cm * c, -h * sm * c);
      D[1][3] = mk<R>(-h * cp * s, -h * sp * s);
      D[2][0] = mk<R>(h * M[0].y, -h * M[0].x);   // d/dw: -i/2, -i/2, +i/2, +i/2
      D[2][1] = mk<R>(h * M[1].y, -h * M[1].x);
      D[2][2] = mk<R>(-h * M[2].y, h * M[2].x);
      D[2][3] = mk<R>(-h * M[3].y, h * M[3].x);
    } break;
    default:
      M[0] = mk<R>(1, 0); M[1] = z; M[2] = z; M[3] = mk<R>(1, 0);
      break;
  }
}

__device__ __forceinline__ bool kind_controlled(int kind) {
  return kind == TQ_G_CRX || kind == TQ_G_CRY || kind == TQ_G_CRZ || kind == TQ_G_CPHASE;
}

// One (parameter set, block) item is handled by a 16-lane group: lane e owns element (e / Dd, e % Dd)
// of every Dd x Dd matrix (Dd = 2 or 4), products go through intra-group shuffles.
constexpr int MAT_LANES = 16;

template <typename R>
__device__ __forceinline__ cx<R> shfl16(cx<R> v, int src, unsigned mask) {
  return mk<R>(__shfl_sync(mask, v.x, src, MAT_LANES), __shfl_sync(mask, v.y, src, MAT_LANES));
}

// my element of A*B, where a / b are my elements of A / B
template <typename R>
__device__ __forceinline__ cx<R> mm_elem(cx<R> a, cx<R> b, int Dd, int r, int c, unsigned mask) {
  cx<R> acc = mk<R>(0, 0);
  for (int k = 0; k < 4; ++k) {  // Dd <= 4; every lane of the group runs all 4 shuffles
    cx<R> x = shfl16(a, (r * Dd + k) & 15, mask);
    cx<R> y = shfl16(b, (k * Dd + c) & 15, mask);
    if (k < Dd) acc = cfma(x, y, acc);
  }
  return acc;
}

template <typename R>
__device__ __forceinline__ cx<R> pick4(const cx<R>* G, int i) {
  cx<R> v = G[0];
  if (i == 1) v = G[1];
  if (i == 2) v = G[2];
  if (i == 3) v = G[3];
  return v;
}

// Element (r, c) of the Dd x Dd embedding of one member gate.
//   G2: the member's 2x2 (1-qubit gate, or target block of a controlled gate when ctl)
//   fx: FIXED member, full matrix in the pool (gdim = 2 or 4)
//   embed: gdim 2 -> 0 acts on block qubit 0 (MSB), 1 on block qubit 1; gdim 4 -> 2 same order, 3 swapped
template <typename R>
__device__ __forceinline__ cx<R> embed_elem(const cx<R>* G2, const cx<R>* fx, int gdim, bool ctl, bool deriv,
                                            int embed, int Dd, int r, int c) {
  const cx<R> z = mk<R>(0, 0);
  if (Dd == 2) return fx ? fx[r * 2 + c] : pick4(G2, r * 2 + c);
  if (gdim == 2 && !ctl) {
    const int r0 = r >> 1, r1 = r & 1, c0 = c >> 1, c1 = c & 1;
    if (embed == 0) {
      if (r1 != c1) return z;
      return fx ? fx[r0 * 2 + c0] : pick4(G2, r0 * 2 + c0);
    }
    if (r0 != c0) return z;
    return fx ? fx[r1 * 2 + c1] : pick4(G2, r1 * 2 + c1);
  }
  int rr = r, cc = c;
  if (embed == 3) {
    rr = ((r & 1) << 1) | (r >> 1);
    cc = ((c & 1) << 1) | (c >> 1);
  }
  if (!ctl) return fx[rr * 4 + cc];
  // controlled: diag(I, G2) in member order (control, target); the identity part has zero derivative
  if (rr < 2 || cc < 2) return (rr == cc && !deriv) ? mk<R>(1, 0) : z;
  return pick4(G2, (rr - 2) * 2 + (cc - 2));
}

template <typename R>
__global__ void k_materialize(const R* __restrict__ params, int n_params, int64_t batch,
                              const MatBlock* __restrict__ blocks, int n_blocks, const MatInstr* __restrict__ instrs,
                              const cx<R>* __restrict__ fixed, cx<R>* __restrict__ stream_f, int64_t stride_f,
                              cx<R>* __restrict__ stream_b, int64_t stride_b, int with_deriv) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t item = t / MAT_LANES;
  const int e = (int)(t % MAT_LANES);
  if (item >= batch * n_blocks) return;  // whole 16-lane groups leave together
  const unsigned mask = 0xFFFFu << (threadIdx.x & 16);
  const int64_t b = item / n_blocks;
  const MatBlock blk = blocks[(int)(item - b * n_blocks)];
  const R* pb = params + b * n_params;
  cx<R>* of = stream_f + b * stride_f + blk.off_f;
  cx<R>* ob = with_deriv ? stream_b + b * stride_b + blk.off_b : nullptr;

  if (blk.mode == MB_FIXED) {
    const MatInstr& ins = instrs[blk.instr_begin];
    for (int i = e; i < blk.count; i += MAT_LANES) {
      cx<R> v = fixed[ins.fixed_off + i];
      of[i] = v;
      if (ob) ob[i] = v;
    }
    return;
  }
  if (blk.mode == MB_NATIVE) {
    if (e != 0) return;
    const MatInstr& ins = instrs[blk.instr_begin];
    cx<R> G[4], D[3][4];
    R p[3];
    for (int i = 0; i < 3; ++i) p[i] = ins.pidx[i] >= 0 ? pb[ins.pidx[i]] : (R)ins.pconst[i];
    param_gate<R>(ins.kind, p, G, D);
    if (blk.diag) {
      of[0] = G[0];
      of[1] = G[3];
      if (ob) {
        ob[0] = G[0];
        ob[1] = G[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (ins.dsel[i] >= 0) {
            ob[2 + 2 * ins.dsel[i]] = D[i][0];
            ob[3 + 2 * ins.dsel[i]] = D[i][3];
          }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) of[i] = G[i];
      if (ob) {
#pragma unroll
        for (int i = 0; i < 4; ++i) ob[i] = G[i];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (ins.dsel[i] >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ob[4 + 4 * ins.dsel[i] + j] = D[i][j];
          }
      }
    }
    return;
  }
  // MB_FUSED: U = E_k ... E_1 ; dU_s = (E_k ... E_{i+1}) dE_i (E_{i-1} ... E_1)
  const int Dd = blk.dim;
  const int DD = Dd * Dd;
  const bool live = e < DD;
  const int r = live ? e / Dd : 0, c = live ? e % Dd : 0;
  const cx<R> ident = mk<R>(r == c ? (R)1 : (R)0, 0);
  cx<R> U = ident;
  cx<R> Pre[MAX_BLOCK_DERIV];
#pragma unroll
  for (int s = 0; s < MAX_BLOCK_DERIV; ++s) Pre[s] = mk<R>(0, 0);
  for (int mi = blk.instr_begin; mi < blk.instr_end; ++mi) {
    const MatInstr ins = instrs[mi];
    cx<R> G[4], D[3][4];
    const cx<R>* fx = nullptr;
    bool ctl = false;
    int gdim = 2;
    if (ins.kind == TQ_G_FIXED) {
      fx = fixed + ins.fixed_off;
      gdim = 1 << ins.nq;
    } else {
      R p[3];
      for (int i = 0; i < 3; ++i) p[i] = ins.pidx[i] >= 0 ? pb[ins.pidx[i]] : (R)ins.pconst[i];
      param_gate<R>(ins.kind, p, G, D);
      ctl = kind_controlled(ins.kind);
      if (ob) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int s = 0; s < MAX_BLOCK_DERIV; ++s)
            if (ins.dsel[i] == s) Pre[s] = U;
      }
    }
    cx<R> E = embed_elem<R>(G, fx, gdim, ctl, false, ins.embed, Dd, r, c);
    U = mm_elem<R>(E, U, Dd, r, c, mask);
  }
  if (live) {
    of[e] = U;
    if (ob) ob[e] = U;
  }
  if (!ob || blk.nderiv == 0) return;
  cx<R> S = ident;
  for (int mi = blk.instr_end - 1; mi >= blk.instr_begin; --mi) {
    const MatInstr ins = instrs[mi];
    cx<R> G[4], D[3][4];
    const cx<R>* fx = nullptr;
    bool ctl = false;
    int gdim = 2;
    if (ins.kind == TQ_G_FIXED) {
      fx = fixed + ins.fixed_off;
      gdim = 1 << ins.nq;
    } else {
      R p[3];
      for (int i = 0; i < 3; ++i) p[i] = ins.pidx[i] >= 0 ? pb[ins.pidx[i]] : (R)ins.pconst[i];
      param_gate<R>(ins.kind, p, G, D);
      ctl = kind_controlled(ins.kind);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (ins.dsel[i] >= 0) {  // uniform across the 16-lane group
          cx<R> pre = mk<R>(0, 0);
#pragma unroll
          for (int s = 0; s < MAX_BLOCK_DERIV; ++s)
            if (ins.dsel[i] == s) pre = Pre[s];
          cx<R> dE = embed_elem<R>(D[i], nullptr, 2, ctl, true, ins.embed, Dd, r, c);
          cx<R> T = mm_elem<R>(dE, pre, Dd, r, c, mask);
          cx<R> dU = mm_elem<R>(S, T, Dd, r, c, mask);
          if (live) ob[DD * (1 + ins.dsel[i]) + e] = dU;
        }
      }
    }
    cx<R> E = embed_elem<R>(G, fx, gdim, ctl, false, ins.embed, Dd, r, c);
    S = mm_elem<R>(S, E, Dd, r, c, mask);
  }
}

// ---------------------------------------------------------------------------
// full gate tensors G [out..., in...] and G^dagger for the tensor-network operands
// (compiled_circuit.py:442-467; adjoint = reshape(d,d).T.conj(), pytorch_backend.py:524-546)
// ---------------------------------------------------------------------------
struct GateT {
  int32_t kind, nq, fixed_off, pad;
  int32_t pidx[3];
  int32_t pad2;
  double pconst[3];
  int64_t out_off;
};

template <typename R>
__global__ void k_gate_tensors(const R* __restrict__ params, int n_params, int64_t batch,
                               const GateT* __restrict__ tab, int n_gates, const cx<R>* __restrict__ fixed,
                               cx<R>* __restrict__ gm, cx<R>* __restrict__ am, int64_t total) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * n_gates) return;
  const int64_t b = t / n_gates;
  const GateT g = tab[(int)(t - b * n_gates)];
  const int Dd = 1 << g.nq;
  cx<R>* G = gm + b * total + g.out_off;
  cx<R>* A = am + b * total + g.out_off;
  if (g.kind == TQ_G_FIXED) {
    for (int r = 0; r < Dd; ++r)
      for (int c = 0; c < Dd; ++c) {
        cx<R> v = fixed[g.fixed_off + r * Dd + c];
        G[r * Dd + c] = v;
        A[c * Dd + r] = conj_(v);
      }
    return;
  }
  R p[3];
  for (int i = 0; i < 3; ++i) p[i] = g.pidx[i] >= 0 ? params[b * n_params + g.pidx[i]] : (R)g.pconst[i];
  cx<R> M[4], D[3][4];
  param_gate<R>(g.kind, p, M, D);
  if (!kind_controlled(g.kind)) {
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 2; ++c) {
        G[r * 2 + c] = M[r * 2 + c];
        A[c * 2 + r] = conj_(M[r * 2 + c]);
      }
    return;
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      cx<R> v = mk<R>(0, 0);
      if (r < 2 || c < 2) {
        if (r == c) v = mk<R>(1, 0);
      } else {
        v = M[(r - 2) * 2 + (c - 2)];
      }
      G[r * 4 + c] = v;
      A[c * 4 + r] = conj_(v);
    }
}

// Chain rule from the gradients of a network's gate operands to the flat parameters (tensor-network mode
// backward): one thread per (parameter set, gate).  g holds CONJUGATED operand gradients inside the per-set
// arena of tq_tn_backward; off_g[gate][e] / off_a[gate][e] = element offset of entry e of the gradient of G
// (ket half) / G^dagger (bra half), -1 where the network has no such entry (constant operands, entries removed
// by tn_simplify).  dL/dtheta = Re sum_e g_e * dT_e/dtheta; the thread owns its parameter slots (one flat slot
// belongs to exactly one gate), so += needs no atomics and accumulates over measurement networks.
template <typename R>
__global__ void k_gate_tensor_grads(const R* __restrict__ params, int n_params, int64_t batch,
                                    const GateT* __restrict__ tab, int n_gates, const cx<R>* __restrict__ arena,
                                    int64_t set_stride, const int32_t* __restrict__ off_g,
                                    const int32_t* __restrict__ off_a, R* __restrict__ grad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * n_gates) return;
  const int64_t b = t / n_gates;
  const int gi = (int)(t - b * n_gates);
  const GateT g = tab[gi];
  if (g.kind == TQ_G_FIXED || (g.pidx[0] < 0 && g.pidx[1] < 0 && g.pidx[2] < 0)) return;
  R p[3];
  for (int i = 0; i < 3; ++i) p[i] = g.pidx[i] >= 0 ? params[b * n_params + g.pidx[i]] : (R)g.pconst[i];
  cx<R> M[4], D[3][4];
  for (int i = 0; i < 3; ++i)
    for (int e = 0; e < 4; ++e) D[i][e] = mk<R>(0, 0);
  param_gate<R>(g.kind, p, M, D);
  const bool ctl = kind_controlled(g.kind);
  const int Dd = ctl ? 4 : 2, base = ctl ? 2 : 0;
  const cx<R>* a = arena + b * set_stride;
  const int32_t* og = off_g + gi * 16;
  const int32_t* oa = off_a + gi * 16;
  for (int i = 0; i < 3; ++i) {
    if (g.pidx[i] < 0) continue;
    R acc = 0;
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 2; ++c) {
        const cx<R> d = D[i][r * 2 + c];
        const int eg = (base + r) * Dd + (base + c), ea = (base + c) * Dd + (base + r);
        if (og[eg] >= 0) {  // Re(g * d)
          const cx<R> gv = a[og[eg]];
          acc += gv.x * d.x - gv.y * d.y;
        }
        if (oa[ea] >= 0) {  // adjoint operand entry = conj(d): Re(g * conj(d))
          const cx<R> gv = a[oa[ea]];
          acc += gv.x * d.x + gv.y * d.y;
        }
      }
    grad[b * n_params + g.pidx[i]] += acc;
  }
}

// ---------------------------------------------------------------------------
// op application on shared-memory tiles
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand_ins(const OpDesc& d, uint32_t g) {
  uint32_t idx = g;
  if (d.nins > 0) idx = insert_zero_bit(idx, d.ins[0]);
  if (d.nins > 1) idx = insert_zero_bit(idx, d.ins[1]);
  if (d.nins > 2) idx = insert_zero_bit(idx, d.ins[2]);
  if (d.nins > 3) idx = insert_zero_bit(idx, d.ins[3]);
  return idx | d.cmask;
}

typedef cx<float> cf;

__device__ __forceinline__ cf lo(const float4& v) { return mk<float>(v.x, v.y); }
__device__ __forceinline__ cf hi(const float4& v) { return mk<float>(v.z, v.w); }
__device__ __forceinline__ float4 pack(cf a, cf b) { return make_float4(a.x, a.y, b.x, b.y); }

// ---- forward paths (ADJ: apply the conjugate transpose) -------------------------
template <bool ADJ>
__device__ __forceinline__ void ld2x2(const cf* M, cf* m) {
  if (ADJ) {
    m[0] = conj_(M[0]); m[1] = conj_(M[2]); m[2] = conj_(M[1]); m[3] = conj_(M[3]);
  } else {
    m[0] = M[0]; m[1] = M[1]; m[2] = M[2]; m[3] = M[3];
  }
}

__device__ __forceinline__ void mv2(const cf* m, cf a0, cf a1, cf& b0, cf& b1) {
  b0 = cfma(m[1], a1, cmul(m[0], a0));
  b1 = cfma(m[3], a1, cmul(m[2], a0));
}

template <bool ADJ>
__device__ __forceinline__ void fwd_d1v(float4* s4, const OpDesc& d, const cf* M, int m) {
  cf mm[4];
  ld2x2<ADJ>(M, mm);
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  // low target bits: half of every 8-lane group starts on the partner word so the 8 lanes of one
  // 128-bit wavefront fall into 8 distinct bank groups
  const uint32_t sw = (d.tpos[0] < 3 && (threadIdx.x & 4)) ? tb : 0u;
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = s4[c ^ sw], y = s4[c ^ sw ^ tb];
    if (sw) {
      float4 t = x;
      x = y;
      y = t;
    }
    cf b0, b1, c0, c1;
    mv2(mm, lo(x), lo(y), b0, b1);
    mv2(mm, hi(x), hi(y), c0, c1);
    s4[c] = pack(b0, c0);
    s4[c | tb] = pack(b1, c1);
  }
}

template <bool ADJ>
__device__ __forceinline__ void fwd_d1p(float4* s4, const OpDesc& d, const cf* M, int m) {
  cf mm[4];
  ld2x2<ADJ>(M, mm);
  const uint32_t ng = 1u << (m - 1 - d.nins);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = s4[c];
    cf b0, b1;
    mv2(mm, lo(x), hi(x), b0, b1);
    s4[c] = pack(b0, b1);
  }
}

template <typename R, bool ADJ>
__device__ __forceinline__ void fwd_d1s(cx<R>* s, const OpDesc& d, const cx<R>* M, int m) {
  cx<R> m0, m1, m2, m3;
  if (ADJ) {
    m0 = conj_(M[0]); m1 = conj_(M[2]); m2 = conj_(M[1]); m3 = conj_(M[3]);
  } else {
    m0 = M[0]; m1 = M[1]; m2 = M[2]; m3 = M[3];
  }
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    cx<R> a0 = s[i], a1 = s[i | tb];
    s[i] = cfma(m1, a1, cmul(m0, a0));
    s[i | tb] = cfma(m3, a1, cmul(m2, a0));
  }
}

template <typename R, bool ADJ>
__device__ __forceinline__ void ld4x4(const cx<R>* M, cx<R>* mm) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) mm[r * 4 + c] = ADJ ? conj_(M[c * 4 + r]) : M[r * 4 + c];
}

template <typename R>
__device__ __forceinline__ void mv4(const cx<R>* mm, const cx<R>* a, cx<R>* b) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    cx<R> acc = cmul(mm[r * 4], a[0]);
#pragma unroll
    for (int c = 1; c < 4; ++c) acc = cfma(mm[r * 4 + c], a[c], acc);
    b[r] = acc;
  }
}

template <bool ADJ>
__device__ __forceinline__ void fwd_d2v(float4* s4, const OpDesc& d, const cf* M, int m) {
  cf mm[16];
  ld4x4<float, ADJ>(M, mm);
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 v0 = s4[c], v1 = s4[c | o1], v2 = s4[c | o2], v3 = s4[c | o1 | o2];
    cf a[4] = {lo(v0), lo(v1), lo(v2), lo(v3)}, b[4];
    cf e[4] = {hi(v0), hi(v1), hi(v2), hi(v3)}, f[4];
    mv4<float>(mm, a, b);
    mv4<float>(mm, e, f);
    s4[c] = pack(b[0], f[0]);
    s4[c | o1] = pack(b[1], f[1]);
    s4[c | o2] = pack(b[2], f[2]);
    s4[c | o1 | o2] = pack(b[3], f[3]);
  }
}

template <typename R, bool ADJ>
__device__ __forceinline__ void fwd_d2s(cx<R>* s, const OpDesc& d, const cx<R>* M, int m) {
  cx<R> mm[16];
  ld4x4<R, ADJ>(M, mm);
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    cx<R> a[4] = {s[i], s[i | o1], s[i | o2], s[i | o1 | o2]}, b[4];
    mv4<R>(mm, a, b);
    s[i] = b[0];
    s[i | o1] = b[1];
    s[i | o2] = b[2];
    s[i | o1 | o2] = b[3];
  }
}

template <bool ADJ>
__device__ __forceinline__ void fwd_g1v(float4* s4, const OpDesc& d, const cf* M, int m) {
  cf d0 = M[0], d1 = M[1];
  if (ADJ) {
    d0 = conj_(d0);
    d1 = conj_(d1);
  }
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const int tp = d.tpos[0];  // amplitude-bit position of the target
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = s4[c];
    cf da, db;
    if (tp == 0) {
      da = d0;
      db = d1;
    } else {
      da = db = ((c >> (tp - 1)) & 1u) ? d1 : d0;
    }
    s4[c] = pack(cmul(da, lo(x)), cmul(db, hi(x)));
  }
}

template <typename R, bool ADJ>
__device__ __forceinline__ void fwd_g1s(cx<R>* s, const OpDesc& d, const cx<R>* M, int m) {
  cx<R> d0 = M[0], d1 = M[1];
  if (ADJ) {
    d0 = conj_(d0);
    d1 = conj_(d1);
  }
  const uint32_t ng = 1u << (m - d.nins);
  const int tp = d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    s[i] = cmul(((i >> tp) & 1u) ? d1 : d0, s[i]);
  }
}

// ---- real blocks: the matrix has no imaginary part (payload entries are complex with .y == 0) ------------------
// y = M x with real M costs 2 real multiplies per complex amplitude entry instead of 4.
template <typename R>
__device__ __forceinline__ cx<R> rmul(R m, cx<R> a) { return mk<R>(m * a.x, m * a.y); }
template <typename R>
__device__ __forceinline__ cx<R> rfma(R m, cx<R> a, cx<R> acc) {
  acc.x += m * a.x;
  acc.y += m * a.y;
  return acc;
}

template <typename R, bool ADJ>
__device__ __forceinline__ void fwd_r1s(cx<R>* s, const OpDesc& d, const cx<R>* M, int m) {
  const R m0 = M[0].x, m1 = ADJ ? M[2].x : M[1].x, m2 = ADJ ? M[1].x : M[2].x, m3 = M[3].x;
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    const cx<R> a0 = s[i], a1 = s[i | tb];
    s[i] = rfma(m1, a1, rmul(m0, a0));
    s[i | tb] = rfma(m3, a1, rmul(m2, a0));
  }
}

template <typename R, bool ADJ>
__device__ __forceinline__ void ld4x4_real(const cx<R>* M, R* mm) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) mm[r * 4 + c] = ADJ ? M[c * 4 + r].x : M[r * 4 + c].x;
}
template <typename R>
__device__ __forceinline__ void mv4_real(const R* mm, const cx<R>* a, cx<R>* b) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    cx<R> acc = rmul(mm[r * 4], a[0]);
#pragma unroll
    for (int c = 1; c < 4; ++c) acc = rfma(mm[r * 4 + c], a[c], acc);
    b[r] = acc;
  }
}

template <typename R, bool ADJ>
__device__ __forceinline__ void fwd_r2s(cx<R>* s, const OpDesc& d, const cx<R>* M, int m) {
  R mm[16];
  ld4x4_real<R, ADJ>(M, mm);
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    cx<R> a[4] = {s[i], s[i | o1], s[i | o2], s[i | o1 | o2]}, b[4];
    mv4_real<R>(mm, a, b);
    s[i] = b[0];
    s[i | o1] = b[1];
    s[i | o2] = b[2];
    s[i | o1 | o2] = b[3];
  }
}

// real twins of the 128-bit complex64 paths: same addressing, real matrix entries
__device__ __forceinline__ void mv2r(const float* m, cf a0, cf a1, cf& b0, cf& b1) {
  b0 = rfma(m[1], a1, rmul(m[0], a0));
  b1 = rfma(m[3], a1, rmul(m[2], a0));
}
template <bool ADJ>
__device__ __forceinline__ void ld2x2r(const cf* M, float* m) {
  m[0] = M[0].x;
  m[1] = ADJ ? M[2].x : M[1].x;
  m[2] = ADJ ? M[1].x : M[2].x;
  m[3] = M[3].x;
}
template <bool ADJ>
__device__ __forceinline__ void fwd_r1v(float4* s4, const OpDesc& d, const cf* M, int m) {
  float mm[4];
  ld2x2r<ADJ>(M, mm);
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  const uint32_t sw = (d.tpos[0] < 3 && (threadIdx.x & 4)) ? tb : 0u;
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = s4[c ^ sw], y = s4[c ^ sw ^ tb];
    if (sw) {
      float4 t = x;
      x = y;
      y = t;
    }
    cf b0, b1, c0, c1;
    mv2r(mm, lo(x), lo(y), b0, b1);
    mv2r(mm, hi(x), hi(y), c0, c1);
    s4[c] = pack(b0, c0);
    s4[c | tb] = pack(b1, c1);
  }
}
template <bool ADJ>
__device__ __forceinline__ void fwd_r1p(float4* s4, const OpDesc& d, const cf* M, int m) {
  float mm[4];
  ld2x2r<ADJ>(M, mm);
  const uint32_t ng = 1u << (m - 1 - d.nins);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = s4[c];
    cf b0, b1;
    mv2r(mm, lo(x), hi(x), b0, b1);
    s4[c] = pack(b0, b1);
  }
}
template <bool ADJ>
__device__ __forceinline__ void fwd_r2v(float4* s4, const OpDesc& d, const cf* M, int m) {
  float mm[16];
  ld4x4_real<float, ADJ>(M, mm);
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 v0 = s4[c], v1 = s4[c | o1], v2 = s4[c | o2], v3 = s4[c | o1 | o2];
    cf a[4] = {lo(v0), lo(v1), lo(v2), lo(v3)}, b[4];
    cf e[4] = {hi(v0), hi(v1), hi(v2), hi(v3)}, f[4];
    mv4_real<float>(mm, a, b);
    mv4_real<float>(mm, e, f);
    s4[c] = pack(b[0], f[0]);
    s4[c | o1] = pack(b[1], f[1]);
    s4[c | o2] = pack(b[2], f[2]);
    s4[c | o1 | o2] = pack(b[3], f[3]);
  }
}

// ---- diagonal layer ----------------------------------------------------------------------------------------------
// Member k acts on tile bit pos[k] with phases (d0, d1) = pay[stride * k], pay[stride * k + 1] (stride 2 in the
// forward stream; 4 in the backward stream, where entries 2, 3 are d(d0), d(d1) of a trainable member).
// phase(i) = prod_k d_k[bit_k(i)] = T_lo[i & 127] * T_hi[i >> 7]: two tables built once per op by the CTA.
struct DlDesc {
  int n, train_mask;
  uint8_t pos[DL_MAX];
};
__device__ __forceinline__ DlDesc dl_decode(const OpDesc& d) {
  DlDesc r;
  r.n = (int)(d.count & 0xffu);
  r.train_mask = (int)((d.count >> 8) & 0xffffu);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.pos[j] = d.ins[j];
    r.pos[4 + j] = d.tpos[j];
    r.pos[8 + j] = (uint8_t)((d.cmask >> (8 * j)) & 0xffu);
    r.pos[12 + j] = (uint8_t)((d.pad >> (8 * j)) & 0xffu);
  }
  return r;
}

// ONE static shared allocation for every diagonal-layer path of a kernel (forward, adjoint and the forward recompute
// inside the adjoint kernel): 2 x 128 phases + 17 reduction slots.  (Separate arrays per path pushed the adjoint
// sweep over the shared-memory budget of 3 CTAs per SM.)
template <typename R>
__device__ __noinline__ cx<R>* dl_smem() {
  __shared__ cx<R> tab[2 * (1 << DL_LO_BITS) + (DL_MAX + 2) / 2 + 1];
  return tab;
}

template <typename R, bool ADJ>
__device__ __forceinline__ void dl_tables(const DlDesc& L, const cx<R>* pay, int stride, int m, cx<R>* t_lo, cx<R>* t_hi) {
  const int hi_bits = m > DL_LO_BITS ? m - DL_LO_BITS : 0;
  const int n_lo = 1 << (m < DL_LO_BITS ? m : DL_LO_BITS), n_hi = 1 << hi_bits;
  for (int e = threadIdx.x; e < n_lo + n_hi; e += blockDim.x) {
    const bool hi = e >= n_lo;
    const uint32_t v = hi ? (uint32_t)(e - n_lo) << DL_LO_BITS : (uint32_t)e;
    cx<R> acc = mk<R>(1, 0);
    for (int k = 0; k < L.n; ++k) {
      const int b = L.pos[k];
      if ((b >= DL_LO_BITS) != hi) continue;
      cx<R> ph = pay[stride * k + ((v >> b) & 1u)];
      if (ADJ) ph = conj_(ph);
      acc = cmul(acc, ph);
    }
    (hi ? t_hi[e - n_lo] : t_lo[e]) = acc;
  }
  __syncthreads();
}

template <typename R, bool ADJ>
__device__ __forceinline__ void fwd_dl(cx<R>* s, const OpDesc& d, const cx<R>* pay, int stride, int m) {
  cx<R>* t_lo = dl_smem<R>();
  cx<R>* t_hi = t_lo + (1 << DL_LO_BITS);
  const DlDesc L = dl_decode(d);
  dl_tables<R, ADJ>(L, pay, stride, m, t_lo, t_hi);
  const uint32_t n = 1u << m;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    s[i] = cmul(cmul(t_lo[i & ((1u << DL_LO_BITS) - 1u)], t_hi[i >> DL_LO_BITS]), s[i]);
}

// generic: dense with k targets (k <= 3) or diagonal with k targets; ins = all inserted bits (amplitude units)
template <typename R, bool ADJ>
__device__ void fwd_gen(cx<R>* s, const OpDesc& d, const cx<R>* M, int m, int cls) {
  const int k = d.k;
  const int Dd = 1 << k;
  const uint32_t ng = 1u << (m - d.nins);
  if (cls == OP_DIAG) {
    for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
      const uint32_t i = expand_ins(d, g);
      int di = 0;
      for (int t = 0; t < k; ++t) di = (di << 1) | ((i >> d.tpos[t]) & 1u);
      cx<R> v = M[di];
      if (ADJ) v = conj_(v);
      s[i] = cmul(v, s[i]);
    }
    return;
  }
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    cx<R> a[8], b[8];
    for (int r = 0; r < Dd; ++r) {
      uint32_t off = 0;
      for (int t = 0; t < k; ++t) off |= ((r >> (k - 1 - t)) & 1u) << d.tpos[t];
      a[r] = s[i | off];
    }
    for (int r = 0; r < Dd; ++r) {
      cx<R> acc = mk<R>(0, 0);
      for (int c = 0; c < Dd; ++c) acc = cfma(ADJ ? conj_(M[c * Dd + r]) : M[r * Dd + c], a[c], acc);
      b[r] = acc;
    }
    for (int r = 0; r < Dd; ++r) {
      uint32_t off = 0;
      for (int t = 0; t < k; ++t) off |= ((r >> (k - 1 - t)) & 1u) << d.tpos[t];
      s[i | off] = b[r];
    }
  }
}

// ST: the structure-aware paths (real blocks, diagonal layers; tq_plan_opts.structure) are compiled into separate
// kernel instantiations, so that the default kernels keep the code (registers, spills) they had without them
template <typename R, bool ADJ, bool ST>
__device__ __forceinline__ void apply_op(cx<R>* s, const OpDesc& d, const cx<R>* pay, int m) {
  const cx<R>* M = pay;
  if constexpr (ST) {
    switch (d.path) {
      case P_R1S: fwd_r1s<R, ADJ>(s, d, M, m); return;
      case P_R2S: fwd_r2s<R, ADJ>(s, d, M, m); return;
      case P_DL: fwd_dl<R, ADJ>(s, d, M, 2, m); return;  // (forward payload stream: 2 entries per member)
      default: break;
    }
  }
  switch (d.path) {
    case P_D1S: fwd_d1s<R, ADJ>(s, d, M, m); break;
    case P_D2S: fwd_d2s<R, ADJ>(s, d, M, m); break;
    case P_G1S: fwd_g1s<R, ADJ>(s, d, M, m); break;
    default: fwd_gen<R, ADJ>(s, d, M, m, d.pad); break;
  }
}
template <bool ADJ, bool ST>
__device__ __forceinline__ void apply_op_f32(cf* s, const OpDesc& d, const cf* pay, int m) {
  float4* s4 = reinterpret_cast<float4*>(s);
  if constexpr (ST) {
    switch (d.path) {
      case P_R1V: fwd_r1v<ADJ>(s4, d, pay, m); return;
      case P_R1P: fwd_r1p<ADJ>(s4, d, pay, m); return;
      case P_R2V: fwd_r2v<ADJ>(s4, d, pay, m); return;
      default: break;
    }
  }
  switch (d.path) {
    case P_D1V: fwd_d1v<ADJ>(s4, d, pay, m); break;
    case P_D1P: fwd_d1p<ADJ>(s4, d, pay, m); break;
    case P_D2V: fwd_d2v<ADJ>(s4, d, pay, m); break;
    case P_G1V: fwd_g1v<ADJ>(s4, d, pay, m); break;
    default: apply_op<float, ADJ, ST>(s, d, pay, m); break;
  }
}
template <typename R, bool ADJ, bool ST>
__device__ __forceinline__ void run_op(cx<R>* s, const OpDesc& d, const cx<R>* pay, int m) {
  if constexpr (sizeof(R) == 4)
    apply_op_f32<ADJ, ST>(s, d, pay, m);
  else
    apply_op<R, ADJ, ST>(s, d, pay, m);
}

// ---- adjoint step: psi <- G^dag psi ; grad_d += Re <lambda | dG_d psi> ; lambda <- G^dag lambda -----------
// The derivative matrices are the same for every amplitude group of the tile, so the groups only
// accumulate W[r][c] = sum_groups psi_prev[c] * conj(lambda[r]); the per-parameter contraction
// grad_d = Re sum_rc dG_d[r][c] W[r][c] happens once per thread and op, whatever the number of
// trainable slots fused into the block.
template <typename R>
__device__ __forceinline__ void wacc(cx<R>& w, cx<R> p, cx<R> l) {  // w += p * conj(l)
  w.x += p.x * l.x;
  w.x += p.y * l.y;
  w.y += p.y * l.x;
  w.y -= p.x * l.y;
}

template <typename R, int DD>
__device__ __forceinline__ void grad_contract(const cx<R>* W, const cx<R>* Dm, int nd, R* s_grad, uint32_t dslot) {
  for (int e = 0; e < nd; ++e) {
    const cx<R>* De = Dm + DD * e;
    R v = 0;
#pragma unroll
    for (int i = 0; i < DD; ++i) v += De[i].x * W[i].x - De[i].y * W[i].y;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_grad[dslot + e], v);
  }
}

// one 2-amplitude group
template <typename R>
__device__ __forceinline__ void bwd2_group(const cx<R>* mh, cx<R>& a0, cx<R>& a1, cx<R>& l0, cx<R>& l1, cx<R>* W,
                                           bool has_d) {
  cx<R> p0 = cfma(mh[1], a1, cmul(mh[0], a0));
  cx<R> p1 = cfma(mh[3], a1, cmul(mh[2], a0));
  if (has_d) {  // uniform per op: fixed blocks carry no gradient
    wacc(W[0], p0, l0);
    wacc(W[1], p1, l0);
    wacc(W[2], p0, l1);
    wacc(W[3], p1, l1);
  }
  cx<R> q0 = cfma(mh[1], l1, cmul(mh[0], l0));
  cx<R> q1 = cfma(mh[3], l1, cmul(mh[2], l0));
  a0 = p0;
  a1 = p1;
  l0 = q0;
  l1 = q1;
}

template <typename R>
__device__ __forceinline__ void bwd4_group(const cx<R>* mh, cx<R>* a, cx<R>* l, cx<R>* W, bool has_d) {
  cx<R> p[4], q[4];
  mv4<R>(mh, a, p);
  if (has_d) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) wacc(W[r * 4 + c], p[c], l[r]);
  }
  mv4<R>(mh, l, q);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    a[r] = p[r];
    l[r] = q[r];
  }
}

template <typename R, int DD>
__device__ __forceinline__ void wzero(cx<R>* W) {
#pragma unroll
  for (int i = 0; i < DD; ++i) W[i] = mk<R>(0, 0);
}

__device__ __forceinline__ void bwd_d1v(float4* p4, float4* l4, const OpDesc& d, const cf* pay, float* s_grad, int m) {
  cf mh[4], W[4];
  ld2x2<true>(pay, mh);
  wzero<float, 4>(W);
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  const uint32_t sw = (d.tpos[0] < 3 && (threadIdx.x & 4)) ? tb : 0u;
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = p4[c ^ sw], y = p4[c ^ sw ^ tb], u = l4[c ^ sw], v = l4[c ^ sw ^ tb];
    if (sw) {
      float4 t = x; x = y; y = t;
      t = u; u = v; v = t;
    }
    cf a0 = lo(x), a1 = lo(y), b0 = hi(x), b1 = hi(y);
    cf k0 = lo(u), k1 = lo(v), n0 = hi(u), n1 = hi(v);
    bwd2_group<float>(mh, a0, a1, k0, k1, W, has_d);
    bwd2_group<float>(mh, b0, b1, n0, n1, W, has_d);
    p4[c] = pack(a0, b0);
    p4[c | tb] = pack(a1, b1);
    l4[c] = pack(k0, n0);
    l4[c | tb] = pack(k1, n1);
  }
  grad_contract<float, 4>(W, pay + 4, d.nderiv, s_grad, d.dslot);
}

__device__ __forceinline__ void bwd_d1p(float4* p4, float4* l4, const OpDesc& d, const cf* pay, float* s_grad, int m) {
  cf mh[4], W[4];
  ld2x2<true>(pay, mh);
  wzero<float, 4>(W);
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - 1 - d.nins);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = p4[c], u = l4[c];
    cf a0 = lo(x), a1 = hi(x), k0 = lo(u), k1 = hi(u);
    bwd2_group<float>(mh, a0, a1, k0, k1, W, has_d);
    p4[c] = pack(a0, a1);
    l4[c] = pack(k0, k1);
  }
  grad_contract<float, 4>(W, pay + 4, d.nderiv, s_grad, d.dslot);
}

template <typename R>
__device__ __forceinline__ void bwd_d1s(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  cx<R> mh[4] = {conj_(pay[0]), conj_(pay[2]), conj_(pay[1]), conj_(pay[3])};
  cx<R> W[4];
  wzero<R, 4>(W);
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    cx<R> a0 = sp[i], a1 = sp[i | tb], l0 = sl[i], l1 = sl[i | tb];
    bwd2_group<R>(mh, a0, a1, l0, l1, W, has_d);
    sp[i] = a0;
    sp[i | tb] = a1;
    sl[i] = l0;
    sl[i | tb] = l1;
  }
  grad_contract<R, 4>(W, pay + 4, d.nderiv, s_grad, d.dslot);
}

__device__ __forceinline__ void bwd_d2v(float4* p4, float4* l4, const OpDesc& d, const cf* pay, float* s_grad, int m) {
  cf mh[16], W[16];
  ld4x4<float, true>(pay, mh);
  wzero<float, 16>(W);
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 v0 = p4[c], v1 = p4[c | o1], v2 = p4[c | o2], v3 = p4[c | o1 | o2];
    float4 w0 = l4[c], w1 = l4[c | o1], w2 = l4[c | o2], w3 = l4[c | o1 | o2];
    cf a[4] = {lo(v0), lo(v1), lo(v2), lo(v3)}, l[4] = {lo(w0), lo(w1), lo(w2), lo(w3)};
    bwd4_group<float>(mh, a, l, W, has_d);
    cf e[4] = {hi(v0), hi(v1), hi(v2), hi(v3)}, f[4] = {hi(w0), hi(w1), hi(w2), hi(w3)};
    bwd4_group<float>(mh, e, f, W, has_d);
    p4[c] = pack(a[0], e[0]);
    p4[c | o1] = pack(a[1], e[1]);
    p4[c | o2] = pack(a[2], e[2]);
    p4[c | o1 | o2] = pack(a[3], e[3]);
    l4[c] = pack(l[0], f[0]);
    l4[c | o1] = pack(l[1], f[1]);
    l4[c | o2] = pack(l[2], f[2]);
    l4[c | o1 | o2] = pack(l[3], f[3]);
  }
  grad_contract<float, 16>(W, pay + 16, d.nderiv, s_grad, d.dslot);
}

// y = G^dag x with G read from shared memory on every use (complex128: 16 matrix entries would
// cost 64 registers on top of the 64 of W and spill)
template <typename R>
__device__ __forceinline__ void mv4_adj_smem(const cx<R>* M, const cx<R>* x, cx<R>* y) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    cx<R> acc = mk<R>(0, 0);
#pragma unroll
    for (int c = 0; c < 4; ++c) acc = cfma_conj(M[c * 4 + r], x[c], acc);
    y[r] = acc;
  }
}

template <typename R>
__device__ __forceinline__ void bwd_d2s(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  cx<R> W[16];
  wzero<R, 16>(W);
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  if (sizeof(R) == 8) {
    for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
      const uint32_t i = expand_ins(d, g);
      cx<R> a[4] = {sp[i], sp[i | o1], sp[i | o2], sp[i | o1 | o2]}, pa[4];
      mv4_adj_smem<R>(pay, a, pa);
      sp[i] = pa[0]; sp[i | o1] = pa[1]; sp[i | o2] = pa[2]; sp[i | o1 | o2] = pa[3];
      cx<R> l[4] = {sl[i], sl[i | o1], sl[i | o2], sl[i | o1 | o2]};
      if (has_d) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) wacc(W[r * 4 + c], pa[c], l[r]);
      }
      mv4_adj_smem<R>(pay, l, a);
      sl[i] = a[0]; sl[i | o1] = a[1]; sl[i | o2] = a[2]; sl[i | o1 | o2] = a[3];
    }
  } else {
    cx<R> mh[16];
    ld4x4<R, true>(pay, mh);
    for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
      const uint32_t i = expand_ins(d, g);
      cx<R> a[4] = {sp[i], sp[i | o1], sp[i | o2], sp[i | o1 | o2]};
      cx<R> l[4] = {sl[i], sl[i | o1], sl[i | o2], sl[i | o1 | o2]};
      bwd4_group<R>(mh, a, l, W, has_d);
      sp[i] = a[0]; sp[i | o1] = a[1]; sp[i | o2] = a[2]; sp[i | o1 | o2] = a[3];
      sl[i] = l[0]; sl[i | o1] = l[1]; sl[i | o2] = l[2]; sl[i | o1 | o2] = l[3];
    }
  }
  grad_contract<R, 16>(W, pay + 16, d.nderiv, s_grad, d.dslot);
}

template <typename R>
__device__ __forceinline__ void bwd_g1_amp(cx<R> dh, cx<R>& a, cx<R>& l, cx<R>& w) {
  cx<R> pa = cmul(dh, a);
  wacc(w, pa, l);
  a = pa;
  l = cmul(dh, l);
}

__device__ __forceinline__ void bwd_g1v(float4* p4, float4* l4, const OpDesc& d, const cf* pay, float* s_grad, int m) {
  const cf h0 = conj_(pay[0]), h1 = conj_(pay[1]);
  cf W[2];
  wzero<float, 2>(W);
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const int tp = d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = p4[c], u = l4[c];
    cf a0 = lo(x), a1 = hi(x), k0 = lo(u), k1 = hi(u);
    if (tp == 0) {
      bwd_g1_amp<float>(h0, a0, k0, W[0]);
      bwd_g1_amp<float>(h1, a1, k1, W[1]);
    } else if ((c >> (tp - 1)) & 1u) {
      bwd_g1_amp<float>(h1, a0, k0, W[1]);
      bwd_g1_amp<float>(h1, a1, k1, W[1]);
    } else {
      bwd_g1_amp<float>(h0, a0, k0, W[0]);
      bwd_g1_amp<float>(h0, a1, k1, W[0]);
    }
    p4[c] = pack(a0, a1);
    l4[c] = pack(k0, k1);
  }
  grad_contract<float, 2>(W, pay + 2, d.nderiv, s_grad, d.dslot);
}

template <typename R>
__device__ __forceinline__ void bwd_g1s(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  const cx<R> h0 = conj_(pay[0]), h1 = conj_(pay[1]);
  cx<R> W[2];
  wzero<R, 2>(W);
  const uint32_t ng = 1u << (m - d.nins);
  const int tp = d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    cx<R> a = sp[i], l = sl[i];
    if ((i >> tp) & 1u)
      bwd_g1_amp<R>(h1, a, l, W[1]);
    else
      bwd_g1_amp<R>(h0, a, l, W[0]);
    sp[i] = a;
    sl[i] = l;
  }
  grad_contract<R, 2>(W, pay + 2, d.nderiv, s_grad, d.dslot);
}

// ---- adjoint step of a REAL block: psi <- M^T psi, lambda <- M^T lambda, and only Re W is needed because the
// derivative matrices of a real block are real: grad_d = sum_rc dM_d[r][c] * Re W[r][c] ---------------------------
template <typename R>
__device__ __forceinline__ void wacc_re(R& w, cx<R> p, cx<R> l) { w += p.x * l.x + p.y * l.y; }

template <typename R, int DD>
__device__ __forceinline__ void grad_contract_real(const R* W, const cx<R>* Dm, int nd, R* s_grad, uint32_t dslot) {
  for (int e = 0; e < nd; ++e) {
    const cx<R>* De = Dm + DD * e;
    R v = 0;
#pragma unroll
    for (int i = 0; i < DD; ++i) v += De[i].x * W[i];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_grad[dslot + e], v);
  }
}

template <typename R>
__device__ __forceinline__ void bwd_r1s(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  const R m0 = pay[0].x, m1 = pay[2].x, m2 = pay[1].x, m3 = pay[3].x;  // M^T
  R W[4] = {0, 0, 0, 0};
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    const cx<R> a0 = sp[i], a1 = sp[i | tb], l0 = sl[i], l1 = sl[i | tb];
    const cx<R> p0 = rfma(m1, a1, rmul(m0, a0)), p1 = rfma(m3, a1, rmul(m2, a0));
    if (has_d) {
      wacc_re(W[0], p0, l0);
      wacc_re(W[1], p1, l0);
      wacc_re(W[2], p0, l1);
      wacc_re(W[3], p1, l1);
    }
    sp[i] = p0;
    sp[i | tb] = p1;
    sl[i] = rfma(m1, l1, rmul(m0, l0));
    sl[i | tb] = rfma(m3, l1, rmul(m2, l0));
  }
  grad_contract_real<R, 4>(W, pay + 4, d.nderiv, s_grad, d.dslot);
}

template <typename R>
__device__ __forceinline__ void bwd_r2s(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  R mh[16], W[16];
  ld4x4_real<R, true>(pay, mh);
#pragma unroll
  for (int i = 0; i < 16; ++i) W[i] = 0;
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t i = expand_ins(d, g);
    cx<R> a[4] = {sp[i], sp[i | o1], sp[i | o2], sp[i | o1 | o2]}, p[4];
    cx<R> l[4] = {sl[i], sl[i | o1], sl[i | o2], sl[i | o1 | o2]}, q[4];
    mv4_real<R>(mh, a, p);
    if (has_d) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) wacc_re(W[r * 4 + c], p[c], l[r]);
    }
    mv4_real<R>(mh, l, q);
    sp[i] = p[0]; sp[i | o1] = p[1]; sp[i | o2] = p[2]; sp[i | o1 | o2] = p[3];
    sl[i] = q[0]; sl[i | o1] = q[1]; sl[i | o2] = q[2]; sl[i | o1 | o2] = q[3];
  }
  grad_contract_real<R, 16>(W, pay + 16, d.nderiv, s_grad, d.dslot);
}

// real twins of the 128-bit complex64 adjoint paths
__device__ __forceinline__ void bwd2r_group(const float* mh, cf& a0, cf& a1, cf& l0, cf& l1, float* W, bool has_d) {
  cf p0, p1, q0, q1;
  mv2r(mh, a0, a1, p0, p1);
  if (has_d) {
    wacc_re(W[0], p0, l0);
    wacc_re(W[1], p1, l0);
    wacc_re(W[2], p0, l1);
    wacc_re(W[3], p1, l1);
  }
  mv2r(mh, l0, l1, q0, q1);
  a0 = p0; a1 = p1; l0 = q0; l1 = q1;
}
__device__ __forceinline__ void bwd4r_group(const float* mh, cf* a, cf* l, float* W, bool has_d) {
  cf p[4], q[4];
  mv4_real<float>(mh, a, p);
  if (has_d) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) wacc_re(W[r * 4 + c], p[c], l[r]);
  }
  mv4_real<float>(mh, l, q);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    a[r] = p[r];
    l[r] = q[r];
  }
}
__device__ __forceinline__ void bwd_r1v(float4* p4, float4* l4, const OpDesc& d, const cf* pay, float* s_grad, int m) {
  float mh[4], W[4] = {0, 0, 0, 0};
  ld2x2r<true>(pay, mh);
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t tb = 1u << d.tpos[0];
  const uint32_t sw = (d.tpos[0] < 3 && (threadIdx.x & 4)) ? tb : 0u;
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = p4[c ^ sw], y = p4[c ^ sw ^ tb], u = l4[c ^ sw], v = l4[c ^ sw ^ tb];
    if (sw) {
      float4 t = x; x = y; y = t;
      t = u; u = v; v = t;
    }
    cf a0 = lo(x), a1 = lo(y), b0 = hi(x), b1 = hi(y);
    cf k0 = lo(u), k1 = lo(v), n0 = hi(u), n1 = hi(v);
    bwd2r_group(mh, a0, a1, k0, k1, W, has_d);
    bwd2r_group(mh, b0, b1, n0, n1, W, has_d);
    p4[c] = pack(a0, b0);
    p4[c | tb] = pack(a1, b1);
    l4[c] = pack(k0, n0);
    l4[c | tb] = pack(k1, n1);
  }
  grad_contract_real<float, 4>(W, pay + 4, d.nderiv, s_grad, d.dslot);
}
__device__ __forceinline__ void bwd_r1p(float4* p4, float4* l4, const OpDesc& d, const cf* pay, float* s_grad, int m) {
  float mh[4], W[4] = {0, 0, 0, 0};
  ld2x2r<true>(pay, mh);
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - 1 - d.nins);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 x = p4[c], u = l4[c];
    cf a0 = lo(x), a1 = hi(x), k0 = lo(u), k1 = hi(u);
    bwd2r_group(mh, a0, a1, k0, k1, W, has_d);
    p4[c] = pack(a0, a1);
    l4[c] = pack(k0, k1);
  }
  grad_contract_real<float, 4>(W, pay + 4, d.nderiv, s_grad, d.dslot);
}
__device__ __forceinline__ void bwd_r2v(float4* p4, float4* l4, const OpDesc& d, const cf* pay, float* s_grad, int m) {
  float mh[16], W[16];
  ld4x4_real<float, true>(pay, mh);
#pragma unroll
  for (int i = 0; i < 16; ++i) W[i] = 0.f;
  const bool has_d = d.nderiv > 0;
  const uint32_t ng = 1u << (m - 1 - d.nins);
  const uint32_t o1 = 1u << d.tpos[1], o2 = 1u << d.tpos[0];
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t c = expand_ins(d, g);
    float4 v0 = p4[c], v1 = p4[c | o1], v2 = p4[c | o2], v3 = p4[c | o1 | o2];
    float4 w0 = l4[c], w1 = l4[c | o1], w2 = l4[c | o2], w3 = l4[c | o1 | o2];
    cf a[4] = {lo(v0), lo(v1), lo(v2), lo(v3)}, l[4] = {lo(w0), lo(w1), lo(w2), lo(w3)};
    bwd4r_group(mh, a, l, W, has_d);
    cf e[4] = {hi(v0), hi(v1), hi(v2), hi(v3)}, f[4] = {hi(w0), hi(w1), hi(w2), hi(w3)};
    bwd4r_group(mh, e, f, W, has_d);
    p4[c] = pack(a[0], e[0]);
    p4[c | o1] = pack(a[1], e[1]);
    p4[c | o2] = pack(a[2], e[2]);
    p4[c | o1 | o2] = pack(a[3], e[3]);
    l4[c] = pack(l[0], f[0]);
    l4[c | o1] = pack(l[1], f[1]);
    l4[c | o2] = pack(l[2], f[2]);
    l4[c | o1 | o2] = pack(l[3], f[3]);
  }
  grad_contract_real<float, 16>(W, pay + 16, d.nderiv, s_grad, d.dslot);
}

// ---- adjoint step of a diagonal layer ------------------------------------------------------------------------------
// psi <- D^dag psi, lambda <- D^dag lambda.  Every member is a unit-modulus phase d_b = e^{i alpha_b(phi)}, so with
// t_i = Im(psi_i conj(lambda_i)) (the same before and after the layer):
//   dL/dphi_k = -alpha_0' W0 - alpha_1' W1,   W_b = sum over amplitudes with bit_k = b of t_i,
// and W0, W1 follow from the plain sum T = sum t_i and the SIGNED sum S_k = sum (-1)^{bit_k(i)} t_i.  A thread's
// amplitudes are i = tid + it * blockDim: a tile bit below log2(blockDim) is fixed per thread (S_k contribution =
// +-T_thread), one above it is an iteration bit (kept in a per-thread signed accumulator).
template <typename R>
__device__ __forceinline__ void bwd_dl(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  cx<R>* t_lo = dl_smem<R>();
  cx<R>* t_hi = t_lo + (1 << DL_LO_BITS);
  R* s_acc = reinterpret_cast<R*>(t_hi + (1 << DL_LO_BITS));  // DL_MAX + 1 reals
  const DlDesc L = dl_decode(d);
  if (threadIdx.x <= DL_MAX) s_acc[threadIdx.x] = 0;
  dl_tables<R, true>(L, pay, 4, m, t_lo, t_hi);  // (ends with a barrier)
  const uint32_t n = 1u << m;
  const int lbd = 31 - __clz((int)blockDim.x);  // blockDim is a power of two
  R T = 0, S_it[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool has_d = L.train_mask != 0;
  uint32_t it = 0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x, ++it) {
    const cx<R> ph = cmul(t_lo[i & ((1u << DL_LO_BITS) - 1u)], t_hi[i >> DL_LO_BITS]);
    const cx<R> a = sp[i], l = sl[i];
    if (has_d) {
      const R t = a.y * l.x - a.x * l.y;
      T += t;
#pragma unroll
      for (int j = 0; j < 8; ++j) S_it[j] += ((it >> j) & 1u) ? -t : t;
    }
    sp[i] = cmul(ph, a);
    sl[i] = cmul(ph, l);
  }
  if (!has_d) return;
  for (int k = 0; k < L.n; ++k) {
    if (!((L.train_mask >> k) & 1)) continue;
    const int b = L.pos[k];
    R v;
    if (b < lbd) {
      v = ((threadIdx.x >> b) & 1u) ? -T : T;
    } else {
      v = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (b - lbd == j) v = S_it[j];
    }
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[k], v);
  }
  {
    const R v = warp_sum(T);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[DL_MAX], v);
  }
  __syncthreads();
  if ((int)threadIdx.x < L.n && ((L.train_mask >> threadIdx.x) & 1)) {
    const int k = threadIdx.x;
    const cx<R> d0 = pay[4 * k], d1 = pay[4 * k + 1], e0 = pay[4 * k + 2], e1 = pay[4 * k + 3];
    const R a0 = e0.y * d0.x - e0.x * d0.y, a1 = e1.y * d1.x - e1.x * d1.y;  // alpha_b' = Im(d(d_b) conj(d_b))
    const R Tt = s_acc[DL_MAX], S = s_acc[k];
    const R w0 = (R)0.5 * (Tt + S), w1 = (R)0.5 * (Tt - S);
    const int slot = __popc(L.train_mask & ((1 << k) - 1));
    atomicAdd(&s_grad[d.dslot + slot], -a0 * w0 - a1 * w1);
  }
}

template <typename R, bool ST>
__device__ __forceinline__ void bwd_op_scalar(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  if constexpr (ST) {
    switch (d.path) {
      case P_R1S: bwd_r1s<R>(sp, sl, d, pay, s_grad, m); return;
      case P_R2S: bwd_r2s<R>(sp, sl, d, pay, s_grad, m); return;
      case P_DL: bwd_dl<R>(sp, sl, d, pay, s_grad, m); return;
      default: break;
    }
  }
  switch (d.path) {
    case P_D1S: bwd_d1s<R>(sp, sl, d, pay, s_grad, m); break;
    case P_D2S: bwd_d2s<R>(sp, sl, d, pay, s_grad, m); break;
    case P_G1S: bwd_g1s<R>(sp, sl, d, pay, s_grad, m); break;
    default:  // generic ops are fixed gates (no parameters): un-apply on both tiles
      fwd_gen<R, true>(sp, d, pay, m, d.pad);
      fwd_gen<R, true>(sl, d, pay, m, d.pad);
      break;
  }
}
template <typename R, bool ST>
__device__ __forceinline__ void bwd_op(cx<R>* sp, cx<R>* sl, const OpDesc& d, const cx<R>* pay, R* s_grad, int m) {
  if constexpr (sizeof(R) == 4) {
    float4* p4 = reinterpret_cast<float4*>(sp);
    float4* l4 = reinterpret_cast<float4*>(sl);
    if constexpr (ST) {
      switch (d.path) {
        case P_R1V: bwd_r1v(p4, l4, d, pay, s_grad, m); return;
        case P_R1P: bwd_r1p(p4, l4, d, pay, s_grad, m); return;
        case P_R2V: bwd_r2v(p4, l4, d, pay, s_grad, m); return;
        default: break;
      }
    }
    switch (d.path) {
      case P_D1V: bwd_d1v(p4, l4, d, pay, s_grad, m); break;
      case P_D1P: bwd_d1p(p4, l4, d, pay, s_grad, m); break;
      case P_D2V: bwd_d2v(p4, l4, d, pay, s_grad, m); break;
      case P_G1V: bwd_g1v(p4, l4, d, pay, s_grad, m); break;
      default: bwd_op_scalar<float, ST>(sp, sl, d, pay, s_grad, m); break;
    }
  } else {
    bwd_op_scalar<R, ST>(sp, sl, d, pay, s_grad, m);
  }
}

// ---------------------------------------------------------------------------
// chunked op stream: descriptors + payload prefetched with cp.async into a 2-deep ring
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

struct StreamRef {
  const OpDesc* ops;        // descriptors of the whole direction (global)
  const ChunkInfo* chunks;  // chunk table of this sweep (global)
  int32_t n_chunks;
};

template <typename R>
struct Ring {
  OpDesc* desc[2];
  cx<R>* pay[2];
  ChunkInfo* table;  // staged chunk table (first MAX_CHUNKS_SMEM entries)
};

template <typename R>
__device__ __forceinline__ Ring<R> ring_carve(unsigned char* base) {
  Ring<R> r;
  r.desc[0] = reinterpret_cast<OpDesc*>(base);
  r.desc[1] = r.desc[0] + CHUNK_OPS;
  r.pay[0] = reinterpret_cast<cx<R>*>(base + 2 * CHUNK_OPS * sizeof(OpDesc));
  r.pay[1] = reinterpret_cast<cx<R>*>(base + 2 * CHUNK_OPS * sizeof(OpDesc) + CHUNK_PAY_BYTES);
  r.table = reinterpret_cast<ChunkInfo*>(base + 2 * CHUNK_OPS * sizeof(OpDesc) + 2 * CHUNK_PAY_BYTES);
  return r;
}
constexpr int RING_BYTES = 2 * CHUNK_OPS * 32 + 2 * CHUNK_PAY_BYTES + MAX_CHUNKS_SMEM * 16;

template <typename R>
__device__ __forceinline__ ChunkInfo chunk_info(const Ring<R>& ring, const StreamRef& st, int c) {
  return c < MAX_CHUNKS_SMEM ? ring.table[c] : st.chunks[c];
}

template <typename R>
__device__ __forceinline__ void ring_issue(const Ring<R>& ring, const StreamRef& st, const cx<R>* pay_b, int c) {
  // payload offsets are multiples of 16 bytes by construction (host pads every op)
  const ChunkInfo ci = chunk_info<R>(ring, st, c);
  const int buf = c & 1;
  const char* dsrc = reinterpret_cast<const char*>(st.ops + ci.op_begin);
  char* ddst = reinterpret_cast<char*>(ring.desc[buf]);
  const int dbytes = ci.op_count * (int)sizeof(OpDesc);
  for (int o = threadIdx.x * 16; o < dbytes; o += blockDim.x * 16) cp_async16(ddst + o, dsrc + o);
  const char* psrc = reinterpret_cast<const char*>(pay_b + ci.pay_begin);
  char* pdst = reinterpret_cast<char*>(ring.pay[buf]);
  const int pbytes = ci.pay_count * (int)sizeof(cx<R>);
  for (int o = threadIdx.x * 16; o < pbytes; o += blockDim.x * 16) cp_async16(pdst + o, psrc + o);
  cp_async_commit();
}

template <typename R>
__device__ __forceinline__ void ring_start(const Ring<R>& ring, const StreamRef& st, const cx<R>* pay_b) {
  const int nt = st.n_chunks < MAX_CHUNKS_SMEM ? st.n_chunks : MAX_CHUNKS_SMEM;
  for (int i = threadIdx.x; i < nt; i += blockDim.x) ring.table[i] = st.chunks[i];
  __syncthreads();
  if (st.n_chunks > 0) ring_issue<R>(ring, st, pay_b, 0);
}

// ---------------------------------------------------------------------------
// measurements / cotangent seed (unchanged semantics: pytorch_backend.py:393-498)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gather_bits(uint32_t i, const int8_t* pos, int nq) {
  uint32_t r = 0;
  for (int t = 0; t < nq; ++t) r = (r << 1) | ((i >> pos[t]) & 1u);
  return r;
}

template <typename R>
__device__ __forceinline__ cx<R> obs_row_dot(const cx<R>* __restrict__ arr, uint32_t i, const DevMeas& ms,
                                             const cx<R>* __restrict__ O) {
  const int nq = ms.nq;
  const int Dd = 1 << nq;
  uint32_t r = gather_bits(i, ms.pos, nq);
  uint32_t base = i;
  for (int t = 0; t < nq; ++t) base &= ~(1u << ms.pos[t]);
  cx<R> acc = mk<R>(0, 0);
  for (int j = 0; j < Dd; ++j) {
    uint32_t idx = base;
    for (int t = 0; t < nq; ++t) idx |= ((j >> (nq - 1 - t)) & 1u) << ms.pos[t];
    acc = cfma(O[r * Dd + j], arr[idx], acc);
  }
  return acc;
}

template <typename R>
__device__ void measure_block(const cx<R>* __restrict__ arr, uint32_t i0, uint32_t cnt,
                              const DevMeas* __restrict__ meas, int n_meas, const cx<R>* __restrict__ fixed,
                              R* __restrict__ out_b, R* s_acc, bool atomic_out) {
  for (int mi = 0; mi < n_meas; ++mi) {
    const DevMeas& ms = meas[mi];
    if (ms.kind == TQ_M_EXPVAL) {
      R acc = 0;
      if (ms.flags & TQ_MF_ZSTRING) {
        const uint32_t zm = ms.zmask;
        for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
          uint32_t i = i0 + j;
          cx<R> a = arr[i];
          R p = a.x * a.x + a.y * a.y;
          acc += (__popc(i & zm) & 1) ? -p : p;
        }
      } else {
        const cx<R>* O = fixed + ms.mat_off;
        for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
          uint32_t i = i0 + j;
          acc += re_conj_mul(arr[i], obs_row_dot<R>(arr, i, ms, O));
        }
      }
      acc = warp_sum(acc);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[ms.slot_base], acc);
    } else if (ms.kind == TQ_M_PROBS) {
      if (ms.slot_base >= 0) {
        const int nb = 1 << ms.nq;
        for (int bin = 0; bin < nb; ++bin) {
          R acc = 0;
          for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
            uint32_t i = i0 + j;
            if ((int)gather_bits(i, ms.pos, ms.nq) == bin) {
              cx<R> a = arr[i];
              acc += a.x * a.x + a.y * a.y;
            }
          }
          acc = warp_sum(acc);
          if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[ms.slot_base + bin], acc);
        }
      } else {
        R* o = out_b + ms.out_off;
        for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
          uint32_t i = i0 + j;
          cx<R> a = arr[i];
          R p = a.x * a.x + a.y * a.y;
          if (ms.nq == 0)
            o[i] = p;
          else
            atomicAdd(&o[gather_bits(i, ms.pos, ms.nq)], p);
        }
      }
    } else {
      R* o = out_b + ms.out_off;
      for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
        uint32_t i = i0 + j;
        cx<R> a = arr[i];
        o[2 * (size_t)i] = a.x;
        o[2 * (size_t)i + 1] = a.y;
      }
    }
  }
  __syncthreads();
  for (int mi = 0; mi < n_meas; ++mi) {
    const DevMeas& ms = meas[mi];
    if (ms.slot_base < 0) continue;
    int ns = ms.kind == TQ_M_EXPVAL ? 1 : (1 << ms.nq);
    for (int s = threadIdx.x; s < ns; s += blockDim.x) {
      R v = s_acc[ms.slot_base + s];
      if (atomic_out)
        atomicAdd(&out_b[ms.out_off + s], v);
      else
        out_b[ms.out_off + s] = v;
    }
  }
}

template <typename R>
__device__ __forceinline__ cx<R> seed_amp(const cx<R>* __restrict__ arr, uint32_t i,
                                          const DevMeas* __restrict__ meas, int n_meas,
                                          const cx<R>* __restrict__ fixed, const R* __restrict__ dy_b) {
  cx<R> g = mk<R>(0, 0);
  const cx<R> a = arr[i];
  for (int mi = 0; mi < n_meas; ++mi) {
    const DevMeas& ms = meas[mi];
    if (ms.kind == TQ_M_EXPVAL) {
      R w = (R)2 * dy_b[ms.out_off];
      if (ms.flags & TQ_MF_ZSTRING) {
        if (__popc(i & ms.zmask) & 1) w = -w;
        g.x += w * a.x;
        g.y += w * a.y;
      } else {
        cx<R> v = obs_row_dot<R>(arr, i, ms, fixed + ms.mat_off);
        g.x += w * v.x;
        g.y += w * v.y;
      }
    } else if (ms.kind == TQ_M_PROBS) {
      uint32_t bin = ms.nq == 0 ? i : gather_bits(i, ms.pos, ms.nq);
      R w = (R)2 * dy_b[ms.out_off + bin];
      g.x += w * a.x;
      g.y += w * a.y;
    } else {
      g.x += dy_b[ms.out_off + 2 * (size_t)i];
      g.y += dy_b[ms.out_off + 2 * (size_t)i + 1];
    }
  }
  return g;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
template <typename R>
struct FwdArgs {
  cx<R>* psi;
  const cx<R>* init_state;
  const cx<R>* stream;  // forward payload stream [batch, stride]
  const cx<R>* fixed;
  const DevMeas* meas;
  R* out;
  int64_t out_reals;
  int64_t stride;
  StreamRef st;
  int32_t n_meas, n_slots;
  int32_t flags;
  int32_t tiles_log2;
  Geom geom;
};

extern __shared__ __align__(16) unsigned char tq_smem[];

template <typename R, bool ST>
__device__ __forceinline__ void run_stream_fwd(cx<R>* sm, const Ring<R>& ring, const StreamRef& st, const cx<R>* pay_b,
                                               int m) {
  ring_start<R>(ring, st, pay_b);
  for (int c = 0; c < st.n_chunks; ++c) {
    cp_async_wait_all();
    __syncthreads();  // chunk c landed; everyone is done with the buffer chunk c+1 will overwrite
    if (c + 1 < st.n_chunks) ring_issue<R>(ring, st, pay_b, c + 1);
    const ChunkInfo ci = chunk_info<R>(ring, st, c);
    const OpDesc* dd = ring.desc[c & 1];
    const cx<R>* pp = ring.pay[c & 1];
    for (uint32_t o = 0; o < ci.op_count; ++o) {
      const OpDesc d = dd[o];
      run_op<R, false, ST>(sm, d, pp + (d.pay_off - ci.pay_begin), m);
      __syncthreads();
    }
  }
}

template <typename R, bool ST = false>
// complex64: 80 registers -> 3 CTAs per SM (the shared-memory limit), +17 % on the 20-qubit sweeps; complex128
// would spill at that cap and keeps 2
__global__ void __launch_bounds__(256, sizeof(R) == 4 ? 3 : 2) k_sweep_fwd(const __grid_constant__ FwdArgs<R> a) {
  const int m = a.geom.m;
  const uint32_t tile_n = 1u << m;
  cx<R>* sm = reinterpret_cast<cx<R>*>(tq_smem);
  Ring<R> ring = ring_carve<R>(tq_smem + sizeof(cx<R>) * tile_n);
  const int64_t b = (int64_t)(blockIdx.x >> a.tiles_log2);
  const uint32_t tile = blockIdx.x & ((1u << a.tiles_log2) - 1u);
  const uint32_t tbase = dep_tile(a.geom, tile);
  const size_t sv = (size_t)1 << a.geom.n;
  cx<R>* psi_b = a.psi ? a.psi + (size_t)b * sv : nullptr;

  if (a.flags & SW_INIT) {
    if (a.init_state) {
      for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sm[l] = a.init_state[tbase | dep_local(a.geom, l)];
    } else {
      for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sm[l] = mk<R>(0, 0);
      __syncthreads();
      if (threadIdx.x == 0 && tbase == 0) sm[0] = mk<R>(1, 0);
    }
  } else {
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sm[l] = psi_b[tbase | dep_local(a.geom, l)];
  }
  run_stream_fwd<R, ST>(sm, ring, a.st, a.stream + b * a.stride, m);  // begins with a __syncthreads

  if (a.flags & SW_STORE) {
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) psi_b[tbase | dep_local(a.geom, l)] = sm[l];
  }
  if (a.flags & SW_MEASURE) {
    R* s_acc = reinterpret_cast<R*>(tq_smem + sizeof(cx<R>) * tile_n + RING_BYTES);
    for (int s = threadIdx.x; s < a.n_slots; s += blockDim.x) s_acc[s] = 0;
    __syncthreads();
    measure_block<R>(sm, 0, tile_n, a.meas, a.n_meas, a.fixed, a.out + (size_t)b * a.out_reals, s_acc, false);
  }
}

template <typename R>
struct MeasArgs {
  const cx<R>* psi;
  const cx<R>* fixed;
  const DevMeas* meas;
  R* out;
  int64_t out_reals;
  int32_t n_meas, n_slots, n;
  int32_t chunk_log2;
};

template <typename R>
__global__ void __launch_bounds__(256) k_measure(const __grid_constant__ MeasArgs<R> a) {
  R* s_acc = reinterpret_cast<R*>(tq_smem);
  const int cl = a.n - a.chunk_log2;
  const int64_t b = (int64_t)(blockIdx.x >> cl);
  const uint32_t chunk = blockIdx.x & ((1u << cl) - 1u);
  for (int s = threadIdx.x; s < a.n_slots; s += blockDim.x) s_acc[s] = 0;
  __syncthreads();
  measure_block<R>(a.psi + ((size_t)b << a.n), chunk << a.chunk_log2, 1u << a.chunk_log2, a.meas, a.n_meas,
                   a.fixed, a.out + (size_t)b * a.out_reals, s_acc, true);
}

template <typename R>
struct SeedArgs {
  const cx<R>* psi;
  cx<R>* lam;
  const cx<R>* fixed;
  const DevMeas* meas;
  const R* dy;
  int64_t out_reals;
  int64_t total;
  int32_t n_meas, n;
};

template <typename R>
__global__ void __launch_bounds__(256) k_seed(const __grid_constant__ SeedArgs<R> a) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t >> a.n;
    uint32_t i = (uint32_t)(t & (((int64_t)1 << a.n) - 1));
    a.lam[t] = seed_amp<R>(a.psi + ((size_t)b << a.n), i, a.meas, a.n_meas, a.fixed, a.dy + b * a.out_reals);
  }
}

template <typename R>
struct BwdArgs {
  cx<R>* psi;
  cx<R>* lam;
  const cx<R>* init_state;
  const cx<R>* stream_f;  // forward payload stream (full mode: recompute)
  const cx<R>* stream_b;  // backward payload stream
  const cx<R>* fixed;
  const DevMeas* meas;
  const R* dy;
  R* grad;
  const int32_t* slot_pidx;
  int64_t out_reals;
  int64_t stride_f, stride_b;
  StreamRef st_f, st_b;
  int32_t n_meas, n_params, n_dslots;
  int32_t flags;
  int32_t tiles_log2;
  int32_t grad_rounds;  // register-group sweeps: butterfly steps before a gradient term goes to its shared-memory cell
  Geom geom;
};

template <typename R, bool ST = false>
__global__ void __launch_bounds__(256, 2) k_sweep_bwd(const __grid_constant__ BwdArgs<R> a) {
  const int m = a.geom.m;
  const uint32_t tile_n = 1u << m;
  cx<R>* sp = reinterpret_cast<cx<R>*>(tq_smem);
  cx<R>* sl = sp + tile_n;
  Ring<R> ring = ring_carve<R>(tq_smem + 2 * sizeof(cx<R>) * tile_n);
  R* s_grad = reinterpret_cast<R*>(tq_smem + 2 * sizeof(cx<R>) * tile_n + RING_BYTES);
  const int64_t b = (int64_t)(blockIdx.x >> a.tiles_log2);
  const uint32_t tile = blockIdx.x & ((1u << a.tiles_log2) - 1u);
  const uint32_t tbase = dep_tile(a.geom, tile);
  const size_t sv = (size_t)1 << a.geom.n;

  for (int s = threadIdx.x; s < a.n_dslots; s += blockDim.x) s_grad[s] = 0;

  if (a.flags & SW_FULL) {
    if (a.psi) {  // tq_forward(with_backward) left psi_final in the workspace: 2^n * 8 B per set, no recompute
      const cx<R>* psi_b = a.psi + (size_t)b * sv;
      for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sp[l] = psi_b[l];
      __syncthreads();
    } else {
      if (a.init_state) {
        for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sp[l] = a.init_state[l];
      } else {
        for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sp[l] = mk<R>(0, 0);
        __syncthreads();
        if (threadIdx.x == 0) sp[0] = mk<R>(1, 0);
      }
      run_stream_fwd<R, ST>(sp, ring, a.st_f, a.stream_f + b * a.stride_f, m);
    }
    const R* dy_b = a.dy + b * a.out_reals;
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x)
      sl[l] = seed_amp<R>(sp, l, a.meas, a.n_meas, a.fixed, dy_b);
  } else {
    cx<R>* psi_b = a.psi + (size_t)b * sv;
    cx<R>* lam_b = a.lam + (size_t)b * sv;
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) {
      uint32_t gi = tbase | dep_local(a.geom, l);
      sp[l] = psi_b[gi];
      sl[l] = lam_b[gi];
    }
  }
  __syncthreads();

  const cx<R>* pay_b = a.stream_b + b * a.stride_b;
  ring_start<R>(ring, a.st_b, pay_b);
  for (int c = 0; c < a.st_b.n_chunks; ++c) {
    cp_async_wait_all();
    __syncthreads();
    if (c + 1 < a.st_b.n_chunks) ring_issue<R>(ring, a.st_b, pay_b, c + 1);
    const ChunkInfo ci = chunk_info<R>(ring, a.st_b, c);
    const OpDesc* dd = ring.desc[c & 1];
    const cx<R>* pp = ring.pay[c & 1];
    for (uint32_t o = 0; o < ci.op_count; ++o) {
      const OpDesc d = dd[o];
      bwd_op<R, ST>(sp, sl, d, pp + (d.pay_off - ci.pay_begin), s_grad, m);
      __syncthreads();
    }
  }

  if (!(a.flags & SW_FULL) && (a.flags & SW_STORE)) {
    cx<R>* psi_b = a.psi + (size_t)b * sv;
    cx<R>* lam_b = a.lam + (size_t)b * sv;
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) {
      uint32_t gi = tbase | dep_local(a.geom, l);
      psi_b[gi] = sp[l];
      lam_b[gi] = sl[l];
    }
  }
  R* grad_b = a.grad + b * a.n_params;
  for (int s = threadIdx.x; s < a.n_dslots; s += blockDim.x) {
    if (a.flags & SW_FULL)
      grad_b[a.slot_pidx[s]] = s_grad[s];
    else
      atomicAdd(&grad_b[a.slot_pidx[s]], s_grad[s]);
  }
}

}  // namespace tq
