// tq_sv_rg.cuh — register-group sweeps of the state-vector engine (complex64, sm_100a).
//
// The default sweeps (tq_sv_kernels.cuh) make one pass over the shared-memory tile per fused block; on circuits made of
// one-qubit layers and sparse entanglers (the hardware-efficient ansatz, tedq/templates/layers.py:96-115) the blocks
// are dense 4x4 products and the pass overhead (descriptor, matrix load, index arithmetic, barrier) is a third of the
// instructions.  Here a thread owns the 16 amplitudes spanned by FOUR tile bits (the group's "register bits"), keeps
// them in registers, and applies a whole run of blocks that live inside those four bits before it writes them back:
//   * one shared-memory round trip per GROUP instead of per block,
//   * one-qubit runs stay 2x2 (8 FMAs per amplitude instead of 16 for the 4x4 product they would be fused into),
//   * CNOT / X blocks at the start or the end of a group cost nothing: a network of them is an affine map over GF(2)
//     on the 4-bit register pattern, folded into the addresses the group loads from / stores to; controlled-X blocks
//     in the middle of a group (and Toffoli) are register swaps,
//   * the adjoint pass accumulates the 2x2 W of every block in registers: one warp reduction per block and thread.
// The tile is stored XOR-swizzled (rg_phys) so that the 16 lanes of a half-warp always hit 16 different 8-byte bank
// pairs whichever four bits are register bits.  Replaces pytorch_backend.py:365-379 + autograd like the default sweeps.
#pragma once
#include <vector>

#include "tq_sv_kernels.cuh"

namespace tq {

enum { P_RG = 14 };                                // header op of a register group (OpDesc.path in the op stream)
enum { RG_D1 = 0, RG_X1 = 1, RG_GEN = 2 };        // sub-op kinds (OpDesc.path of the descriptors after a header)
constexpr int RG_BITS = 4;                         // register bits of a group
constexpr int RG_MAX_SUB = CHUNK_OPS - 1;          // header + sub-ops travel in one prefetch chunk
constexpr int RG_MIN_TILE = 9;                     // 2^(m-4) >= 32 items: every lane of a warp owns a group
constexpr uint32_t RG_MAP_ID = 0x8421u;            // identity relabelling
enum { RG_PX = 1, RG_PY = 2, RG_PZ = 3, RG_PP = 4 };  // generator of a trainable member: RX, RY, RZ, PhaseShift (and controlled)

// Header:  path = P_RG, nins = number of sub-ops, tpos[0..3] = register bits (tile-local amplitude-bit positions,
//          ascending), and the swizzled word offsets of the LOAD and STORE relabellings, five 16-bit values each
//          (RgAddr: offset of register pattern j = c ^ XOR_{b in j} b_b), packed by rg_pack_header.  A relabelling
//          is an affine map over GF(2) on the register pattern j (bits 4b..4b+3 = image of the unit pattern e_b,
//          bits 16..19 = constant): register j is loaded from the tile pattern load(j), stored to store(j).
// Sub-op:  path = kind, k = register-bit INDEX (0..3) of the target (RG_D1, RG_X1), cmask = 16-bit "live" mask (bit j
//          set when register pattern j satisfies the block's controls), nderiv / dslot / pay_off / count as in the
//          default ops (count = 2: the payload is a diagonal, applied as a 2x2 with zero off-diagonals).
//          pad = generator codes: (first) | (last) << 4 when every trainable slot of the block belongs to a Pauli
//          rotation that is the block's first or last member (RG_PX .. RG_PP); 0: gradients through W.
//          RG_GEN (multi-target diagonal, no trainable slot): ins[0..3] | tpos[0..3] hold sixteen 4-bit diagonal
//          indices, one per register pattern.

__host__ __device__ __forceinline__ uint32_t rg_phys(uint32_t i) {
  return i ^ ((i >> 4) & 15u) ^ ((i >> 8) & 15u) ^ ((i >> 12) & 15u);
}

// the two 16-byte halves of an OpDesc, decoded
struct RgSub {
  uint32_t kind, k, nsub, nderiv, count, live, tlo, thi, pay_off, dslot, gen;
};
__host__ __device__ __forceinline__ RgSub rg_decode(uint4 w0, uint4 w1) {
  RgSub d;
  d.kind = w0.x & 255u;            // path
  d.k = (w0.x >> 8) & 255u;        // k
  d.nsub = (w0.x >> 16) & 255u;    // nins
  d.nderiv = w0.x >> 24;           // nderiv
  d.tlo = w0.y;                    // ins[0..3]
  d.thi = w0.z;                    // tpos[0..3]
  d.live = w0.w;                   // cmask
  d.pay_off = w1.x;
  d.dslot = w1.y;
  d.count = w1.z;
  d.gen = w1.w;                    // pad: generator codes of the block's trainable slots (RG_GEN_*), 0 = use W
  return d;
}

// word offsets (swizzled) of the 16 register patterns under an affine relabelling: off(j) = c ^ XOR_{b in j} b_b
struct RgAddr {
  uint32_t b1, b2, b4, b8, c;
};
__host__ __device__ __forceinline__ uint32_t rg_img(const uint32_t* o, uint32_t nib) {
  return ((nib & 1u) ? o[0] : 0u) ^ ((nib & 2u) ? o[1] : 0u) ^ ((nib & 4u) ? o[2] : 0u) ^ ((nib & 8u) ? o[3] : 0u);
}
__host__ __device__ __forceinline__ RgAddr rg_addr(const uint32_t* o, uint32_t map) {
  RgAddr A;
  A.b1 = rg_img(o, map & 15u);
  A.b2 = rg_img(o, (map >> 4) & 15u);
  A.b4 = rg_img(o, (map >> 8) & 15u);
  A.b8 = rg_img(o, (map >> 12) & 15u);
  A.c = rg_img(o, (map >> 16) & 15u);
  return A;
}
#define RG_OFF(A, j) ((A).c ^ (((j) & 1) ? (A).b1 : 0u) ^ (((j) & 2) ? (A).b2 : 0u) ^ (((j) & 4) ? (A).b4 : 0u) ^ (((j) & 8) ? (A).b8 : 0u))

// header words: w0 = (x: path | k | nins | nderiv, y: ins, z: tpos, w: cmask), w1 = (x: pay_off, y: dslot, z: count)
__host__ __device__ __forceinline__ RgAddr rg_header_load(const uint4 w0, const uint4 w1) {
  RgAddr A;
  A.b1 = w0.y & 0xffffu;
  A.b2 = w0.y >> 16;
  A.b4 = w0.w & 0xffffu;
  A.b8 = w0.w >> 16;
  A.c = w1.x & 0xffffu;
  return A;
}
__host__ __device__ __forceinline__ RgAddr rg_header_store(const uint4 w1) {
  RgAddr A;
  A.b1 = w1.y & 0xffffu;
  A.b2 = w1.y >> 16;
  A.b4 = w1.z & 0xffffu;
  A.b8 = w1.z >> 16;
  A.c = w1.x >> 16;
  return A;
}

// swizzled word index of the group's pattern-0 amplitude for item g (the m-4 non-register bits of the tile index)
__host__ __device__ __forceinline__ uint32_t rg_base(uint32_t regbits /* tpos[0..3] packed */, uint32_t g) {
  uint32_t b = insert_zero_bit(g, (int)(regbits & 255u));
  b = insert_zero_bit(b, (int)((regbits >> 8) & 255u));
  b = insert_zero_bit(b, (int)((regbits >> 16) & 255u));
  b = insert_zero_bit(b, (int)(regbits >> 24));
  return rg_phys(b);
}
__host__ __device__ __forceinline__ void rg_bit_offsets(uint32_t regbits, uint32_t* o) {
  o[0] = rg_phys(1u << (regbits & 255u));
  o[1] = rg_phys(1u << ((regbits >> 8) & 255u));
  o[2] = rg_phys(1u << ((regbits >> 16) & 255u));
  o[3] = rg_phys(1u << (regbits >> 24));
}

// ---- arithmetic on the 16 register amplitudes (host-callable: tests/native/rg_check.cu runs the same code) ---------
__host__ __device__ __forceinline__ void rg_mv2(const cx<float>* m, cx<float>& a0, cx<float>& a1) {
  const cx<float> b0 = cfma(m[1], a1, cmul(m[0], a0));
  const cx<float> b1 = cfma(m[3], a1, cmul(m[2], a0));
  a0 = b0;
  a1 = b1;
}

template <int T>
__host__ __device__ __forceinline__ void rg_d1(cx<float> (&a)[16], const cx<float>* m, uint32_t live) {
  constexpr int tb = 1 << T;
  if (live == 0xffffu) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (!(j & tb)) rg_mv2(m, a[j], a[j | tb]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (!(j & tb) && ((live >> j) & 1u)) rg_mv2(m, a[j], a[j | tb]);
  }
}

template <int T>
__host__ __device__ __forceinline__ void rg_x1(cx<float> (&a)[16], uint32_t live) {
  constexpr int tb = 1 << T;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (!(j & tb) && ((live >> j) & 1u)) {
      const cx<float> t = a[j];
      a[j] = a[j | tb];
      a[j | tb] = t;
    }
}

template <bool ADJ>
__host__ __device__ __forceinline__ void rg_gen(cx<float> (&a)[16], uint32_t tlo, uint32_t thi, uint32_t live,
                                                const cx<float>* pay) {
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if ((live >> j) & 1u) {
      const uint32_t idx = ((j < 8 ? tlo : thi) >> (4 * (j & 7))) & 15u;
      cx<float> v = pay[idx];
      if (ADJ) v = conj_(v);
      a[j] = cmul(v, a[j]);
    }
}

// the 2x2 of a sub-op from its first four payload entries p[0..3] (count = 2: p[0], p[1] are a diagonal); ADJ:
// conjugate transpose
template <bool ADJ>
__host__ __device__ __forceinline__ void rg_ld2x2(const cx<float>* p, uint32_t count, cx<float>* m) {
  const cx<float> z = mk<float>(0.f, 0.f);
  if (count == 2) {
    m[0] = ADJ ? conj_(p[0]) : p[0];
    m[1] = z;
    m[2] = z;
    m[3] = ADJ ? conj_(p[1]) : p[1];
  } else if (ADJ) {
    m[0] = conj_(p[0]); m[1] = conj_(p[2]); m[2] = conj_(p[1]); m[3] = conj_(p[3]);
  } else {
    m[0] = p[0]; m[1] = p[1]; m[2] = p[2]; m[3] = p[3];
  }
}

// forward sub-op; m = the 2x2 (RG_D1), pay = the payload (RG_GEN only)
__host__ __device__ __forceinline__ void rg_fwd_sub(cx<float> (&a)[16], const RgSub& d, const cx<float>* m,
                                                    const cx<float>* pay) {
  switch (d.kind * 4u + (d.kind == RG_GEN ? 0u : d.k)) {
    case RG_D1 * 4 + 0: rg_d1<0>(a, m, d.live); break;
    case RG_D1 * 4 + 1: rg_d1<1>(a, m, d.live); break;
    case RG_D1 * 4 + 2: rg_d1<2>(a, m, d.live); break;
    case RG_D1 * 4 + 3: rg_d1<3>(a, m, d.live); break;
    case RG_X1 * 4 + 0: rg_x1<0>(a, d.live); break;
    case RG_X1 * 4 + 1: rg_x1<1>(a, d.live); break;
    case RG_X1 * 4 + 2: rg_x1<2>(a, d.live); break;
    case RG_X1 * 4 + 3: rg_x1<3>(a, d.live); break;
    default: rg_gen<false>(a, d.tlo, d.thi, d.live, pay); break;
  }
}

// adjoint step of a 2x2 block on the register amplitudes: psi <- G^dag psi, W += psi_prev (x) conj(lambda),
// lambda <- G^dag lambda (the conventions of bwd2_group / grad_contract in tq_sv_kernels.cuh)
__host__ __device__ __forceinline__ void rg_wacc(cx<float>& w, cx<float> p, cx<float> l) {  // w += p * conj(l)
  w.x += p.x * l.x;
  w.x += p.y * l.y;
  w.y += p.y * l.x;
  w.y -= p.x * l.y;
}
__host__ __device__ __forceinline__ void rg_bwd2(const cx<float>* mh, cx<float>& a0, cx<float>& a1, cx<float>& l0,
                                                 cx<float>& l1, cx<float>* W, bool has_d) {
  rg_mv2(mh, a0, a1);
  if (has_d) {  // uniform per sub-op: fixed blocks carry no gradient
    rg_wacc(W[0], a0, l0);
    rg_wacc(W[1], a1, l0);
    rg_wacc(W[2], a0, l1);
    rg_wacc(W[3], a1, l1);
  }
  rg_mv2(mh, l0, l1);
}

template <int T>
__host__ __device__ __forceinline__ void rg_bwd_d1(cx<float> (&a)[16], cx<float> (&l)[16], const cx<float>* mh,
                                                   uint32_t live, bool has_d, cx<float>* W) {
  constexpr int tb = 1 << T;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (!(j & tb) && ((live >> j) & 1u)) rg_bwd2(mh, a[j], a[j | tb], l[j], l[j | tb], W, has_d);
}

// Gradient of a Pauli-rotation slot without W.  A block U whose LAST member is L = exp(-i theta P / 2) has
// dU/dtheta = (-i/2 P) U, so dL/dtheta = Re <lambda | (-i/2 P) | psi> with both vectors at the block's OUTPUT (before
// the un-apply); one whose FIRST member is such a rotation has dU/dtheta = U (-i/2 P): the same expression at the
// block's INPUT (after the un-apply of psi and lambda).  PhaseShift: generator i |1><1|.  Four multiply-adds per pair
// and slot instead of the sixteen of W (the factor 1/2 is applied once, by the caller).
template <int P>
__host__ __device__ __forceinline__ float rg_gen_term(cx<float> l0, cx<float> l1, cx<float> a0, cx<float> a1) {
  if (P == RG_PX) return (l0.x * a1.y - l0.y * a1.x) + (l1.x * a0.y - l1.y * a0.x);   // Im(conj(l0) a1 + conj(l1) a0)
  if (P == RG_PY) return (l1.x * a0.x + l1.y * a0.y) - (l0.x * a1.x + l0.y * a1.y);   // Re(conj(l1) a0 - conj(l0) a1)
  if (P == RG_PZ) return (l0.x * a0.y - l0.y * a0.x) - (l1.x * a1.y - l1.y * a1.x);   // Im(conj(l0) a0 - conj(l1) a1)
  return -2.f * (l1.x * a1.y - l1.y * a1.x);                                         // 2 Re(conj(l1) i a1)
}
template <int T, int P>
__host__ __device__ __forceinline__ float rg_gen_sum(const cx<float> (&a)[16], const cx<float> (&l)[16], uint32_t live) {
  constexpr int tb = 1 << T;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (!(j & tb) && ((live >> j) & 1u)) s += rg_gen_term<P>(l[j], l[j | tb], a[j], a[j | tb]);
  return 0.5f * s;
}
template <int T>
__host__ __device__ __forceinline__ float rg_gen_sum_dyn(const cx<float> (&a)[16], const cx<float> (&l)[16], uint32_t live,
                                                         uint32_t P) {
  switch (P) {
    case RG_PX: return rg_gen_sum<T, RG_PX>(a, l, live);
    case RG_PY: return rg_gen_sum<T, RG_PY>(a, l, live);
    case RG_PZ: return rg_gen_sum<T, RG_PZ>(a, l, live);
    default: return rg_gen_sum<T, RG_PP>(a, l, live);
  }
}
// adjoint step with generator gradients: W[0].x / W[0].y receive the sums of the block's slots in slot order
template <int T>
__host__ __device__ __forceinline__ void rg_bwd_d1_gen(cx<float> (&a)[16], cx<float> (&l)[16], const cx<float>* mh,
                                                       uint32_t live, uint32_t gen, cx<float>* W) {
  const uint32_t gf = gen & 15u, gl = gen >> 4;
  float s_last = 0.f, s_first = 0.f;
  if (gl) s_last = rg_gen_sum_dyn<T>(a, l, live, gl);
  rg_bwd_d1<T>(a, l, mh, live, false, W);
  if (gf) s_first = rg_gen_sum_dyn<T>(a, l, live, gf);
  if (gf) {
    W[0].x = s_first;
    W[0].y = s_last;
  } else {
    W[0].x = s_last;
  }
}

// what one thread adds to a gradient slot of a 2x2 sub-op: Re sum_rc dG[r][c] W[r][c]; De = the slot's derivative
// entries (4 dense, 2 diagonal)
__host__ __device__ __forceinline__ float rg_grad_term(const cx<float>* W, const cx<float>* De, uint32_t count) {
  if (count == 2) return De[0].x * W[0].x - De[0].y * W[0].y + De[1].x * W[3].x - De[1].y * W[3].y;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) v += De[i].x * W[i].x - De[i].y * W[i].y;
  return v;
}

// adjoint sub-op; mh = conjugate transpose of the 2x2 (RG_D1).  Returns true when W holds gradient contributions the
// caller has to reduce (RG_D1 with trainable slots)
__host__ __device__ __forceinline__ bool rg_bwd_sub(cx<float> (&a)[16], cx<float> (&l)[16], const RgSub& d,
                                                    const cx<float>* mh, const cx<float>* pay, cx<float>* W) {
  const bool has_d = d.nderiv > 0;
  if (d.kind == RG_D1 && d.gen) {
    switch (d.k) {
      case 0: rg_bwd_d1_gen<0>(a, l, mh, d.live, d.gen, W); break;
      case 1: rg_bwd_d1_gen<1>(a, l, mh, d.live, d.gen, W); break;
      case 2: rg_bwd_d1_gen<2>(a, l, mh, d.live, d.gen, W); break;
      default: rg_bwd_d1_gen<3>(a, l, mh, d.live, d.gen, W); break;
    }
    return true;
  }
  switch (d.kind * 4u + (d.kind == RG_GEN ? 0u : d.k)) {
    case RG_D1 * 4 + 0: rg_bwd_d1<0>(a, l, mh, d.live, has_d, W); return has_d;
    case RG_D1 * 4 + 1: rg_bwd_d1<1>(a, l, mh, d.live, has_d, W); return has_d;
    case RG_D1 * 4 + 2: rg_bwd_d1<2>(a, l, mh, d.live, has_d, W); return has_d;
    case RG_D1 * 4 + 3: rg_bwd_d1<3>(a, l, mh, d.live, has_d, W); return has_d;
    // a permutation is its own adjoint
    case RG_X1 * 4 + 0: rg_x1<0>(a, d.live); rg_x1<0>(l, d.live); return false;
    case RG_X1 * 4 + 1: rg_x1<1>(a, d.live); rg_x1<1>(l, d.live); return false;
    case RG_X1 * 4 + 2: rg_x1<2>(a, d.live); rg_x1<2>(l, d.live); return false;
    case RG_X1 * 4 + 3: rg_x1<3>(a, d.live); rg_x1<3>(l, d.live); return false;
    default:
      rg_gen<true>(a, d.tlo, d.thi, d.live, pay);
      rg_gen<true>(l, d.tlo, d.thi, d.live, pay);
      return false;
  }
}

// ---- descriptors (host) ------------------------------------------------------------------------------------------------
// treg / creg: register-bit INDEX (0..3) of every target (most significant bit of the matrix index first) / control
inline void rg_make_sub(int cls, int ntargets, const int* treg, int nctrl, const int* creg, bool is_x, int count,
                        int nderiv, OpDesc& d, uint32_t gen = 0) {
  memset(&d, 0, sizeof(d));
  d.pad = ntargets == 1 && nderiv > 0 ? gen : 0u;
  uint32_t cm = 0;
  for (int c = 0; c < nctrl; ++c) cm |= 1u << creg[c];
  uint32_t live = 0;
  for (uint32_t j = 0; j < 16; ++j)
    if ((j & cm) == cm) live |= 1u << j;
  d.cmask = live;
  d.nderiv = (uint8_t)nderiv;
  d.count = (uint32_t)count;
  if (ntargets == 1) {
    d.path = (is_x && nderiv == 0) ? RG_X1 : RG_D1;
    d.k = (uint8_t)treg[0];
    return;
  }
  d.path = RG_GEN;  // diagonal on several targets
  d.k = (uint8_t)ntargets;
  (void)cls;
  for (uint32_t j = 0; j < 16; ++j) {
    uint32_t idx = 0;
    for (int t = 0; t < ntargets; ++t) idx = (idx << 1) | ((j >> treg[t]) & 1u);
    uint8_t* byte = (j < 8 ? d.ins : d.tpos) + ((j & 7) >> 1);
    *byte |= (uint8_t)(idx << (4 * (j & 1)));
  }
}

// four register bits for a group whose blocks touch the tile bits in `used` (<= 4 of the m tile bits): padded so that the
// four lowest NON-register bits — the ones consecutive lanes run through — sit in four different columns of the
// swizzle (bit position mod 4), which makes every shared-memory access of the group bank-conflict free
inline void rg_pick_bits(int m, const int* used, int n_used, int* reg) {
  auto conflict_free = [&](const int* r) {
    int seen = 0, cnt = 0;
    for (int b = 0; b < m && cnt < 4; ++b) {
      bool is_reg = false;
      for (int i = 0; i < 4; ++i) is_reg = is_reg || r[i] == b;
      if (is_reg) continue;
      if (seen & (1 << (b & 3))) return false;
      seen |= 1 << (b & 3);
      ++cnt;
    }
    return true;
  };
  int freeb[32], nf = 0;
  for (int b = 0; b < m; ++b) {
    bool u = false;
    for (int i = 0; i < n_used; ++i) u = u || used[i] == b;
    if (!u) freeb[nf++] = b;
  }
  const int need = 4 - n_used;
  int best[4], cand[4];
  // enumerate pad choices (at most C(13, 3) = 286), keep the first conflict-free one, else the lowest free bits
  int idx[3] = {0, 1, 2};
  auto fill = [&](int* out) {
    for (int i = 0; i < n_used; ++i) out[i] = used[i];
    for (int i = 0; i < need; ++i) out[n_used + i] = freeb[idx[i]];
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j)
        if (out[j] < out[i]) {
          int t = out[i];
          out[i] = out[j];
          out[j] = t;
        }
  };
  fill(best);
  if (need == 0) {
    for (int i = 0; i < 4; ++i) reg[i] = best[i];
    return;
  }
  while (true) {
    fill(cand);
    if (conflict_free(cand)) {
      for (int i = 0; i < 4; ++i) best[i] = cand[i];
      break;
    }
    int k = need - 1;
    while (k >= 0 && idx[k] == nf - need + k) --k;
    if (k < 0) break;
    ++idx[k];
    for (int i = k + 1; i < need; ++i) idx[i] = idx[i - 1] + 1;
  }
  for (int i = 0; i < 4; ++i) reg[i] = best[i];
}

// Affine relabelling of the register pattern by a network of X / CNOT blocks (targets and controls as register-bit
// indices, ctl < 0: plain X), written as the 20-bit map of the header.  forward_order = true: the image of pattern j
// after the blocks ran in the given order (STORE map: register j goes to tile pattern net(j)); false: the pre-image
// (LOAD map: register j comes from the tile pattern that the network sends to j).
inline uint32_t rg_affine_map(const int* tgt, const int* ctl, int n_ops, bool forward_order) {
  uint32_t T[16];
  for (uint32_t j = 0; j < 16; ++j) {
    uint32_t x = j;
    // every block is an involution, so the pre-image applies them last to first
    for (int s = 0; s < n_ops; ++s) {
      const int i = forward_order ? s : n_ops - 1 - s;
      if (ctl[i] < 0 || ((x >> ctl[i]) & 1u)) x ^= 1u << tgt[i];
    }
    T[j] = x;
  }
  uint32_t map = T[0] << 16;
  for (int b = 0; b < 4; ++b) map |= (T[1u << b] ^ T[0]) << (4 * b);
  return map;
}

// the header of a group: register bits, number of sub-ops, load / store relabelling as swizzled word offsets
inline void rg_pack_header(const int* reg, int n_sub, uint32_t load_map, uint32_t store_map, OpDesc& h) {
  memset(&h, 0, sizeof(h));
  h.path = P_RG;
  h.k = RG_BITS;
  h.nins = (uint8_t)n_sub;
  uint32_t regbits = 0, o[4];
  for (int i = 0; i < RG_BITS; ++i) {
    h.tpos[i] = (uint8_t)reg[i];
    regbits |= (uint32_t)reg[i] << (8 * i);
  }
  rg_bit_offsets(regbits, o);
  const RgAddr L = rg_addr(o, load_map), S = rg_addr(o, store_map);
  const uint32_t ins = L.b1 | (L.b2 << 16);
  memcpy(h.ins, &ins, 4);
  h.cmask = L.b4 | (L.b8 << 16);
  h.pay_off = L.c | (S.c << 16);
  h.dslot = S.b1 | (S.b2 << 16);
  h.count = S.b4 | (S.b8 << 16);
}

// ---- group formation (host) -----------------------------------------------------------------------------------------
// One block of a sweep as the grouper sees it: its tile bits, whether it is an X / CNOT (foldable into the load / store
// addresses), its payload entries.
struct RgItem {
  int bits[4];
  int nbits;
  bool foldable;
  int pay;
};
// Takes the next group out of `rest` (indices into items, execution order).  pre: foldable blocks before the first
// other block; post: foldable blocks after the last one; mid: everything else, in order.  A block joins the open group
// when the group's tile bits plus its own stay within four and none of its bits belongs to a block that was passed over
// (it commutes with every passed-over block then); a block that is not foldable also must not share a bit with an
// accepted post block (it runs before them).  A foldable block in the pre phase has to touch a bit the group already
// has: otherwise unrelated CNOTs fill the four bits before the blocks they belong with arrive.
inline void rg_next_group(const std::vector<RgItem>& items, std::vector<int>& rest, int m_t, int pay_cap,
                          std::vector<int>& pre, std::vector<int>& mid, std::vector<int>& post, std::vector<char>& inb) {
  pre.clear();
  mid.clear();
  post.clear();
  inb.assign(m_t, 0);
  std::vector<char> blocked(m_t, 0), postb(m_t, 0);
  std::vector<int> keep;
  int nbits = 0, pay = 0;
  for (int bi : rest) {
    const RgItem& it = items[bi];
    bool ok = (int)mid.size() < RG_MAX_SUB;
    int need = 0;
    for (int k = 0; k < it.nbits; ++k) {
      const int b = it.bits[k];
      if (blocked[b]) ok = false;
      if (!it.foldable && postb[b]) ok = false;
      if (!inb[b]) ++need;
    }
    if (it.foldable && mid.empty() && nbits > 0 && need == it.nbits) ok = false;  // pre phase: connected growth only
    const int pe = it.foldable ? 0 : it.pay;
    if (ok && nbits + need <= RG_BITS && pay + pe <= pay_cap) {
      for (int k = 0; k < it.nbits; ++k)
        if (!inb[it.bits[k]]) {
          inb[it.bits[k]] = 1;
          ++nbits;
        }
      pay += pe;
      if (!it.foldable) {
        mid.push_back(bi);
      } else if (mid.empty()) {
        pre.push_back(bi);
      } else {
        post.push_back(bi);
        for (int k = 0; k < it.nbits; ++k) postb[it.bits[k]] = 1;
      }
    } else {
      for (int k = 0; k < it.nbits; ++k) blocked[it.bits[k]] = 1;
      keep.push_back(bi);
    }
  }
  rest.swap(keep);
}

#ifdef __CUDACC__
// ---- explicit shared-memory accesses (32-bit shared addresses: no generic-address loads on the dispatch path) --------
__device__ __forceinline__ uint32_t rg_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 rg_lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void rg_lds_c2(uint32_t a, cf& u, cf& v) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(u.x), "=f"(u.y), "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
}

// ---- one group on a shared-memory tile ------------------------------------------------------------------------------
// hdr: shared address of the header (its sub-op descriptors follow it); pay: shared address of the chunk's payload
// buffer, pay_gen the same as a pointer (RG_GEN indexes it), pay_begin the chunk's first entry.  The store map is read
// (volatile) after the sub-op loop so that the 16 store addresses are not kept alive across it.
__device__ __forceinline__ void rg_run_fwd(cf* s, uint32_t hdr, uint32_t pay, const cf* pay_gen, uint32_t pay_begin,
                                           int m) {
  const uint4 h0 = rg_lds_u4(hdr);
  const uint32_t sub = hdr + 32u;
  const uint32_t nsub = (h0.x >> 16) & 255u;
  const uint32_t ng = 1u << (m - RG_BITS);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t pb = rg_base(h0.z, g);
    cf a[16];
    {
      const RgAddr L = rg_header_load(h0, rg_lds_u4(hdr + 16u));
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = s[pb ^ RG_OFF(L, j)];
    }
    for (uint32_t i = 0; i < nsub; ++i) {
      const RgSub d = rg_decode(rg_lds_u4(sub + 32u * i), rg_lds_u4(sub + 32u * i + 16u));
      const uint32_t pa = pay + 8u * (d.pay_off - pay_begin);
      cf p[4], mm[4];
      rg_lds_c2(pa, p[0], p[1]);
      if (d.count != 2) rg_lds_c2(pa + 16u, p[2], p[3]);
      rg_ld2x2<false>(p, d.count, mm);
      rg_fwd_sub(a, d, mm, pay_gen + (d.pay_off - pay_begin));
    }
    {
      const RgAddr S = rg_header_store(rg_lds_u4(hdr + 16u));
#pragma unroll
      for (int j = 0; j < 16; ++j) s[pb ^ RG_OFF(S, j)] = a[j];
    }
  }
}

// Gradient contributions go to per-thread-group CELLS in shared memory, cells[slot][threadIdx >> rounds], after `rounds`
// butterfly steps inside the group (rounds = 0: every thread owns a cell: no shuffle, no atomic — a cell has one
// writer); the kernel adds the cells of a slot up once, at its end.  (One warp reduction + shared-memory atomic per
// slot and block was a fifth of the adjoint sweep: five dependent shuffles, and float atomics on shared memory are
// compare-and-swap loops.)  ng is a multiple of blockDim (host: threads = min(256, 2^(m-4)) >= 32).
constexpr int RG_GRAD_DIRECT = 31;  // rounds value: no cells (they would not fit): warp sum + atomic on the gradient itself
struct RgGrad {
  float* cells;             // shared memory: n_dslots x (blockDim >> rounds)
  float* direct;            // RG_GRAD_DIRECT: this set's gradient row in global memory
  const int32_t* slot_pidx;
  int rounds;
};
// DIRECT is a template parameter: the no-cells path lives in its own kernel instantiation so that the common kernel keeps
// its code (the run-time branch cost 5 % of the adjoint sweep)
template <bool DIRECT>
__device__ __forceinline__ void rg_grad_cell(const RgGrad& G, uint32_t slot, float v) {
  if constexpr (DIRECT) {
    v = warp_sum(v);
    if ((threadIdx.x & 31u) == 0) atomicAdd(&G.direct[G.slot_pidx[slot]], v);
  } else {
    for (int r = 0; r < G.rounds; ++r) v += __shfl_xor_sync(0xffffffffu, v, 1 << r);
    if ((threadIdx.x & ((1u << G.rounds) - 1u)) == 0)
      G.cells[slot * (blockDim.x >> G.rounds) + (threadIdx.x >> G.rounds)] += v;
  }
}
template <bool DIRECT>
__device__ __forceinline__ void rg_run_bwd(cf* sp, cf* sl, uint32_t hdr, uint32_t pay, const cf* pay_gen,
                                           uint32_t pay_begin, const RgGrad& grad, int m) {
  const uint4 h0 = rg_lds_u4(hdr);
  const uint32_t sub = hdr + 32u;
  const uint32_t nsub = (h0.x >> 16) & 255u;
  const uint32_t ng = 1u << (m - RG_BITS);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t pb = rg_base(h0.z, g);
    cf a[16], l[16];
    {
      const RgAddr L = rg_header_load(h0, rg_lds_u4(hdr + 16u));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        a[j] = sp[pb ^ RG_OFF(L, j)];
        l[j] = sl[pb ^ RG_OFF(L, j)];
      }
    }
    for (uint32_t i = 0; i < nsub; ++i) {
      const RgSub d = rg_decode(rg_lds_u4(sub + 32u * i), rg_lds_u4(sub + 32u * i + 16u));
      const uint32_t pa = pay + 8u * (d.pay_off - pay_begin);
      cf p[4], mh[4], W[4];
      rg_lds_c2(pa, p[0], p[1]);
      if (d.count != 2) rg_lds_c2(pa + 16u, p[2], p[3]);
      rg_ld2x2<true>(p, d.count, mh);
#pragma unroll
      for (int e = 0; e < 4; ++e) W[e] = mk<float>(0.f, 0.f);
      if (rg_bwd_sub(a, l, d, mh, pay_gen + (d.pay_off - pay_begin), W)) {
        if (d.gen) {  // generator sums, in slot order
          for (uint32_t e = 0; e < d.nderiv; ++e) rg_grad_cell<DIRECT>(grad, d.dslot + e, e == 0 ? W[0].x : W[0].y);
        } else
        for (uint32_t e = 0; e < d.nderiv; ++e) {
          cf De[4];
          if (d.count == 2) {
            rg_lds_c2(pa + 16u + 16u * e, De[0], De[1]);
          } else {
            rg_lds_c2(pa + 32u + 32u * e, De[0], De[1]);
            rg_lds_c2(pa + 48u + 32u * e, De[2], De[3]);
          }
          rg_grad_cell<DIRECT>(grad, d.dslot + e, rg_grad_term(W, De, d.count));
        }
      }
    }
    {
      const RgAddr S = rg_header_store(rg_lds_u4(hdr + 16u));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        sp[pb ^ RG_OFF(S, j)] = a[j];
        sl[pb ^ RG_OFF(S, j)] = l[j];
      }
    }
  }
}

// natural <-> swizzled order of a tile, in place: inside every aligned block of 16 words rg_phys is an XOR with a
// constant, hence an involution made of disjoint swaps
__device__ __forceinline__ void rg_swizzle_tile(cf* s, uint32_t tile_n) {
  for (uint32_t i = threadIdx.x; i < tile_n; i += blockDim.x) {
    const uint32_t p = rg_phys(i);
    if (i < p) {
      const cf t = s[i];
      s[i] = s[p];
      s[p] = t;
    }
  }
  __syncthreads();
}

// ---- tile <-> state vector: pairs of amplitudes (tile bit 0 is state bit 0: the host keeps >= 1 coalesce bit), the
// scatter of the tile index split into a per-thread part and a per-iteration table (blockDim is a power of two, so
// the two parts have disjoint bits and dep_local is an OR of the parts) -------------------------------------------------
constexpr int RG_IO_TAB = 64;
__device__ __forceinline__ void rg_io_table(const Geom& g, uint32_t tile_n, uint32_t* tab) {
  const uint32_t iters = tile_n / (2u * blockDim.x);
  for (uint32_t i = threadIdx.x; i < iters; i += blockDim.x) tab[i] = dep_local(g, 2u * i * blockDim.x);
}
__device__ __forceinline__ void rg_cp_async8(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src));
}
// STORE = false: asynchronous copies (LDGSTS, 8 bytes each: the two amplitudes of a pair may land swapped), every copy
// of the tile in flight at once; they are complete after the next cp_async_wait_all + barrier (rg_stream begins with
// one)
template <bool STORE>
__device__ __forceinline__ void rg_tile_io(cf* sm, cf* hbm, const uint32_t* tab, uint32_t base, uint32_t tile_n) {
  const uint32_t iters = tile_n / (2u * blockDim.x);
  for (uint32_t it = 0; it < iters; ++it) {
    const uint32_t q2 = 2u * (it * blockDim.x + threadIdx.x);
    const uint32_t p = rg_phys(q2);  // q2 + 1 sits at p ^ 1
    cf* gp = hbm + (base | tab[it]);
    if (STORE) {
      const float4 v = *reinterpret_cast<const float4*>(sm + (p & ~1u));
      *reinterpret_cast<float4*>(gp) = (p & 1u) ? make_float4(v.z, v.w, v.x, v.y) : v;
    } else {
      rg_cp_async8(sm + p, gp);
      rg_cp_async8(sm + (p ^ 1u), gp + 1);
    }
  }
  if (!STORE) cp_async_commit();
}

template <bool BWD, bool DIRECT>
__device__ __forceinline__ void rg_stream(cf* sp, cf* sl, const Ring<float>& ring, const StreamRef& st, const cf* pay_b,
                                          const RgGrad& grad, int m) {
  ring_start<float>(ring, st, pay_b);
  for (int c = 0; c < st.n_chunks; ++c) {
    cp_async_wait_all();
    __syncthreads();  // chunk c landed; everyone is done with the buffer chunk c+1 will overwrite
    if (c + 1 < st.n_chunks) ring_issue<float>(ring, st, pay_b, c + 1);
    const ChunkInfo ci = chunk_info<float>(ring, st, c);
    const cf* pp = (c & 1) ? ring.pay[1] : ring.pay[0];
    const uint32_t dd = rg_saddr((c & 1) ? ring.desc[1] : ring.desc[0]);
    const uint32_t pa = rg_saddr(pp);
    for (uint32_t o = 0; o < ci.op_count;) {
      const uint32_t hdr = dd + 32u * o;
      if (BWD)
        rg_run_bwd<DIRECT>(sp, sl, hdr, pa, pp, ci.pay_begin, grad, m);
      else
        rg_run_fwd(sp, hdr, pa, pp, ci.pay_begin, m);
      o += 1u + ((rg_lds_u4(hdr).x >> 16) & 255u);
      __syncthreads();
    }
  }
}

// ---- kernels (same arguments, flags and shared-memory layout as k_sweep_fwd / k_sweep_bwd) --------------------------
__global__ void __launch_bounds__(256, 3) k_rg_fwd(const __grid_constant__ FwdArgs<float> a) {
  __shared__ uint32_t io_tab[RG_IO_TAB];
  const int m = a.geom.m;
  const uint32_t tile_n = 1u << m;
  cf* sm = reinterpret_cast<cf*>(tq_smem);
  Ring<float> ring = ring_carve<float>(tq_smem + sizeof(cf) * tile_n);
  const int64_t b = (int64_t)(blockIdx.x >> a.tiles_log2);
  const uint32_t tile = blockIdx.x & ((1u << a.tiles_log2) - 1u);
  const uint32_t tbase = dep_tile(a.geom, tile) | dep_local(a.geom, 2u * threadIdx.x);
  const size_t sv = (size_t)1 << a.geom.n;
  cf* psi_b = a.psi ? a.psi + (size_t)b * sv : nullptr;
  rg_io_table(a.geom, tile_n, io_tab);
  __syncthreads();

  if (a.flags & SW_INIT) {
    if (a.init_state) {
      rg_tile_io<false>(sm, const_cast<cf*>(a.init_state), io_tab, tbase, tile_n);
    } else {
      for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sm[l] = mk<float>(0.f, 0.f);
      __syncthreads();
      if (threadIdx.x == 0 && dep_tile(a.geom, tile) == 0) sm[0] = mk<float>(1.f, 0.f);  // rg_phys(0) = 0
    }
  } else {
    rg_tile_io<false>(sm, psi_b, io_tab, tbase, tile_n);
  }
  const RgGrad no_grad = {nullptr, nullptr, nullptr, 0};
  rg_stream<false, false>(sm, nullptr, ring, a.st, a.stream + b * a.stride, no_grad, m);  // begins and ends with a barrier
  cp_async_wait_all();  // (a sweep without ops: the tile copies still have to land)
  __syncthreads();

  if (a.flags & SW_STORE) rg_tile_io<true>(sm, psi_b, io_tab, tbase, tile_n);
  if (a.flags & SW_MEASURE) {
    __syncthreads();
    rg_swizzle_tile(sm, tile_n);  // back to natural order for the measurement code
    float* s_acc = reinterpret_cast<float*>(tq_smem + sizeof(cf) * tile_n + RING_BYTES);
    for (int s = threadIdx.x; s < a.n_slots; s += blockDim.x) s_acc[s] = 0;
    __syncthreads();
    measure_block<float>(sm, 0, tile_n, a.meas, a.n_meas, a.fixed, a.out + (size_t)b * a.out_reals, s_acc, false);
  }
}

template <bool DIRECT>
__global__ void __launch_bounds__(256, 2) k_rg_bwd(const __grid_constant__ BwdArgs<float> a) {
  __shared__ uint32_t io_tab[RG_IO_TAB];
  const int m = a.geom.m;
  const uint32_t tile_n = 1u << m;
  cf* sp = reinterpret_cast<cf*>(tq_smem);
  cf* sl = sp + tile_n;
  Ring<float> ring = ring_carve<float>(tq_smem + 2 * sizeof(cf) * tile_n);
  float* s_grad = reinterpret_cast<float*>(tq_smem + 2 * sizeof(cf) * tile_n + RING_BYTES);
  const int64_t b = (int64_t)(blockIdx.x >> a.tiles_log2);
  const uint32_t tile = blockIdx.x & ((1u << a.tiles_log2) - 1u);
  const uint32_t tbase = dep_tile(a.geom, tile) | dep_local(a.geom, 2u * threadIdx.x);
  const size_t sv = (size_t)1 << a.geom.n;
  cf* psi_b = a.psi + (size_t)b * sv;  // the forward pass (with_backward) left the final state here
  cf* lam_b = a.lam ? a.lam + (size_t)b * sv : nullptr;

  const int rounds = a.grad_rounds;
  const uint32_t ncell = DIRECT ? 0u : blockDim.x >> (DIRECT ? 0 : rounds);
  float* grad_b = a.grad + b * a.n_params;
  const RgGrad grad = {s_grad, grad_b, a.slot_pidx, rounds};
  for (uint32_t s = threadIdx.x; s < (uint32_t)a.n_dslots * ncell; s += blockDim.x) s_grad[s] = 0;

  if (a.flags & SW_FULL) {
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sp[l] = psi_b[l];
    __syncthreads();
    const float* dy_b = a.dy + b * a.out_reals;
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x)
      sl[l] = seed_amp<float>(sp, l, a.meas, a.n_meas, a.fixed, dy_b);
    __syncthreads();
    rg_swizzle_tile(sp, tile_n);
    rg_swizzle_tile(sl, tile_n);
  } else {
    rg_io_table(a.geom, tile_n, io_tab);
    __syncthreads();
    rg_tile_io<false>(sp, psi_b, io_tab, tbase, tile_n);
    rg_tile_io<false>(sl, lam_b, io_tab, tbase, tile_n);
  }
  rg_stream<true, DIRECT>(sp, sl, ring, a.st_b, a.stream_b + b * a.stride_b, grad, m);  // begins and ends with a barrier
  cp_async_wait_all();
  __syncthreads();

  if (!(a.flags & SW_FULL) && (a.flags & SW_STORE)) {
    rg_tile_io<true>(sp, psi_b, io_tab, tbase, tile_n);
    rg_tile_io<true>(sl, lam_b, io_tab, tbase, tile_n);
  }
  // one warp per slot adds its cells up (tq_backward zeroed the gradient, so the whole-state mode may add as well)
  if constexpr (!DIRECT) {
    for (int s = threadIdx.x >> 5; s < a.n_dslots; s += blockDim.x >> 5) {
      float v = 0.f;
      for (uint32_t c = threadIdx.x & 31u; c < ncell; c += 32u) v += s_grad[(uint32_t)s * ncell + c];
      v = warp_sum(v);
      if ((threadIdx.x & 31u) == 0) {
        if (a.flags & SW_FULL)
          grad_b[a.slot_pidx[s]] = v;
        else
          atomicAdd(&grad_b[a.slot_pidx[s]], v);
      }
    }
  }
}
#endif  // __CUDACC__

}  // namespace tq
