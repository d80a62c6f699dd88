// tq_sv_rg.cuh — register-group sweeps of the state-vector engine (complex64, sm_100a).
//
// The default sweeps (tq_sv_kernels.cuh) make one pass over the shared-memory tile per fused block; on circuits made of
// one-qubit layers and sparse entanglers (the hardware-efficient ansatz, tedq/templates/layers.py:96-115) the blocks
// are dense 4x4 products and the pass overhead (descriptor, matrix load, index arithmetic, barrier) is a third of the
// instructions.  Here a thread owns the 16 amplitudes spanned by FOUR tile bits (the group's "register bits"), keeps
// them in registers, and applies a whole run of blocks that live inside those four bits before it writes them back:
//   * one shared-memory round trip per GROUP instead of per block,
//   * one-qubit runs stay 2x2 (8 FMAs per amplitude instead of 16 for the 4x4 product they would be fused into),
//   * CNOT / Toffoli-style controlled-X blocks are register swaps (no arithmetic),
//   * the adjoint pass accumulates the 2x2 W of every block in registers: one warp reduction per block and thread.
// The tile is stored XOR-swizzled (rg_phys) so that the 16 lanes of a half-warp always hit 16 different 8-byte bank
// pairs whichever four bits are register bits.  Replaces pytorch_backend.py:365-379 + autograd like the default sweeps.
#pragma once
#include "tq_sv_kernels.cuh"

namespace tq {

enum { P_RG = 14 };                                // header op of a register group (OpDesc.path in the op stream)
enum { RG_D1 = 0, RG_X1 = 1, RG_GEN = 2 };        // sub-op kinds (OpDesc.path of the descriptors after a header)
constexpr int RG_BITS = 4;                         // register bits of a group
constexpr int RG_MAX_SUB = CHUNK_OPS - 1;          // header + sub-ops travel in one prefetch chunk
constexpr int RG_MIN_TILE = 9;                     // 2^(m-4) >= 32 items: every lane of a warp owns a group

// Header:  path = P_RG, nins = number of sub-ops, tpos[0..3] = register bits (tile-local amplitude-bit positions,
//          ascending).
// Sub-op:  path = kind, k = register-bit INDEX (0..3) of the target (RG_D1, RG_X1), cmask = 16-bit "live" mask (bit j
//          set when register pattern j satisfies the block's controls), nderiv / dslot / pay_off / count as in the
//          default ops (count = 2: the payload is a diagonal, applied as a 2x2 with zero off-diagonals).
//          RG_GEN (multi-target diagonal, no trainable slot): ins[0..3] | tpos[0..3] hold sixteen 4-bit diagonal
//          indices, one per register pattern.

__host__ __device__ __forceinline__ uint32_t rg_phys(uint32_t i) {
  return i ^ ((i >> 4) & 15u) ^ ((i >> 8) & 15u) ^ ((i >> 12) & 15u);
}

struct RgGeom {
  int r0, r1, r2, r3;
  uint32_t off[16];  // swizzled word offset of register pattern j (XOR-linear: phys(base | off) = phys(base) ^ off[j])
};

__host__ __device__ __forceinline__ RgGeom rg_geom(const OpDesc& h) {
  RgGeom G;
  G.r0 = h.tpos[0];
  G.r1 = h.tpos[1];
  G.r2 = h.tpos[2];
  G.r3 = h.tpos[3];
  const uint32_t o1 = rg_phys(1u << G.r0), o2 = rg_phys(1u << G.r1), o4 = rg_phys(1u << G.r2), o8 = rg_phys(1u << G.r3);
#pragma unroll
  for (int j = 0; j < 16; ++j) G.off[j] = ((j & 1) ? o1 : 0u) ^ ((j & 2) ? o2 : 0u) ^ ((j & 4) ? o4 : 0u) ^ ((j & 8) ? o8 : 0u);
  return G;
}

__host__ __device__ __forceinline__ uint32_t rg_base(const RgGeom& G, uint32_t g) {
  uint32_t b = insert_zero_bit(g, G.r0);
  b = insert_zero_bit(b, G.r1);
  b = insert_zero_bit(b, G.r2);
  b = insert_zero_bit(b, G.r3);
  return rg_phys(b);
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

// the sixteen 4-bit diagonal indices of an RG_GEN sub-op: patterns 0..7 in ins[0..3], 8..15 in tpos[0..3]
__host__ __device__ __forceinline__ uint32_t rg_tab_lo(const OpDesc& d) {
  return (uint32_t)d.ins[0] | ((uint32_t)d.ins[1] << 8) | ((uint32_t)d.ins[2] << 16) | ((uint32_t)d.ins[3] << 24);
}
__host__ __device__ __forceinline__ uint32_t rg_tab_hi(const OpDesc& d) {
  return (uint32_t)d.tpos[0] | ((uint32_t)d.tpos[1] << 8) | ((uint32_t)d.tpos[2] << 16) | ((uint32_t)d.tpos[3] << 24);
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

// the 2x2 of a sub-op: dense payload (4 entries) or diagonal payload (2 entries); ADJ: conjugate transpose
template <bool ADJ>
__host__ __device__ __forceinline__ void rg_ld2x2(const cx<float>* pay, uint32_t count, cx<float>* m) {
  const cx<float> z = mk<float>(0.f, 0.f);
  if (count == 2) {
    m[0] = ADJ ? conj_(pay[0]) : pay[0];
    m[1] = z;
    m[2] = z;
    m[3] = ADJ ? conj_(pay[1]) : pay[1];
  } else if (ADJ) {
    m[0] = conj_(pay[0]); m[1] = conj_(pay[2]); m[2] = conj_(pay[1]); m[3] = conj_(pay[3]);
  } else {
    m[0] = pay[0]; m[1] = pay[1]; m[2] = pay[2]; m[3] = pay[3];
  }
}

__host__ __device__ __forceinline__ void rg_fwd_sub(cx<float> (&a)[16], const OpDesc& d, const cx<float>* pay) {
  const uint32_t live = d.cmask;
  switch (d.path) {
    case RG_D1: {
      cx<float> m[4];
      rg_ld2x2<false>(pay, d.count, m);
      switch (d.k) {
        case 0: rg_d1<0>(a, m, live); break;
        case 1: rg_d1<1>(a, m, live); break;
        case 2: rg_d1<2>(a, m, live); break;
        default: rg_d1<3>(a, m, live); break;
      }
    } break;
    case RG_X1:
      switch (d.k) {
        case 0: rg_x1<0>(a, live); break;
        case 1: rg_x1<1>(a, live); break;
        case 2: rg_x1<2>(a, live); break;
        default: rg_x1<3>(a, live); break;
      }
      break;
    default: {
      rg_gen<false>(a, rg_tab_lo(d), rg_tab_hi(d), live, pay);
    } break;
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

// what one thread adds to gradient slot e of a 2x2 sub-op: Re sum_rc dG_e[r][c] W[r][c]
__host__ __device__ __forceinline__ float rg_grad_term(const cx<float>* W, const cx<float>* pay, uint32_t count, int e) {
  if (count == 2) {
    const cx<float>* De = pay + 2 + 2 * e;
    return De[0].x * W[0].x - De[0].y * W[0].y + De[1].x * W[3].x - De[1].y * W[3].y;
  }
  const cx<float>* De = pay + 4 + 4 * e;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) v += De[i].x * W[i].x - De[i].y * W[i].y;
  return v;
}

// returns true when W holds gradient contributions the caller has to reduce (RG_D1 with trainable slots)
__host__ __device__ __forceinline__ bool rg_bwd_sub(cx<float> (&a)[16], cx<float> (&l)[16], const OpDesc& d,
                                                    const cx<float>* pay, cx<float>* W) {
  const uint32_t live = d.cmask;
  switch (d.path) {
    case RG_D1: {
      cx<float> mh[4];
      rg_ld2x2<true>(pay, d.count, mh);
      const bool has_d = d.nderiv > 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) W[i] = mk<float>(0.f, 0.f);
      switch (d.k) {
        case 0: rg_bwd_d1<0>(a, l, mh, live, has_d, W); break;
        case 1: rg_bwd_d1<1>(a, l, mh, live, has_d, W); break;
        case 2: rg_bwd_d1<2>(a, l, mh, live, has_d, W); break;
        default: rg_bwd_d1<3>(a, l, mh, live, has_d, W); break;
      }
      return has_d;
    }
    case RG_X1:  // a permutation is its own adjoint
      switch (d.k) {
        case 0: rg_x1<0>(a, live); rg_x1<0>(l, live); break;
        case 1: rg_x1<1>(a, live); rg_x1<1>(l, live); break;
        case 2: rg_x1<2>(a, live); rg_x1<2>(l, live); break;
        default: rg_x1<3>(a, live); rg_x1<3>(l, live); break;
      }
      return false;
    default: {
      rg_gen<true>(a, rg_tab_lo(d), rg_tab_hi(d), live, pay);
      rg_gen<true>(l, rg_tab_lo(d), rg_tab_hi(d), live, pay);
      return false;
    }
  }
}

// ---- descriptors (host) ------------------------------------------------------------------------------------------------
// treg / creg: register-bit INDEX (0..3) of every target (most significant bit of the matrix index first) / control
inline void rg_make_sub(int cls, int ntargets, const int* treg, int nctrl, const int* creg, bool is_x, int count,
                        int nderiv, OpDesc& d) {
  memset(&d, 0, sizeof(d));
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
  bool have = false;
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
      have = true;
      break;
    }
    int k = need - 1;
    while (k >= 0 && idx[k] == nf - need + k) --k;
    if (k < 0) break;
    ++idx[k];
    for (int i = k + 1; i < need; ++i) idx[i] = idx[i - 1] + 1;
  }
  (void)have;
  for (int i = 0; i < 4; ++i) reg[i] = best[i];
}

#ifdef __CUDACC__
// ---- one group on a shared-memory tile ------------------------------------------------------------------------------
__device__ __forceinline__ void rg_run_fwd(cf* s, const OpDesc& h, const OpDesc* sub, const cf* pp, uint32_t pay_begin,
                                           int m) {
  const RgGeom G = rg_geom(h);
  const int nsub = h.nins;
  const uint32_t ng = 1u << (m - RG_BITS);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t pb = rg_base(G, g);
    cf a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = s[pb ^ G.off[j]];
    for (int i = 0; i < nsub; ++i) {
      const OpDesc& d = sub[i];
      rg_fwd_sub(a, d, pp + (d.pay_off - pay_begin));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s[pb ^ G.off[j]] = a[j];
  }
}

// ng is a multiple of blockDim (host: threads = min(256, 2^(m-4)) >= 32), so every lane takes part in the reductions
__device__ __forceinline__ void rg_run_bwd(cf* sp, cf* sl, const OpDesc& h, const OpDesc* sub, const cf* pp,
                                           uint32_t pay_begin, float* s_grad, int m) {
  const RgGeom G = rg_geom(h);
  const int nsub = h.nins;
  const uint32_t ng = 1u << (m - RG_BITS);
  for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
    const uint32_t pb = rg_base(G, g);
    cf a[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      a[j] = sp[pb ^ G.off[j]];
      l[j] = sl[pb ^ G.off[j]];
    }
    for (int i = 0; i < nsub; ++i) {
      const OpDesc& d = sub[i];
      const cf* pay = pp + (d.pay_off - pay_begin);
      cf W[4];
      if (rg_bwd_sub(a, l, d, pay, W)) {
        for (int e = 0; e < d.nderiv; ++e) {
          const float v = warp_sum(rg_grad_term(W, pay, d.count, e));
          if ((threadIdx.x & 31) == 0) atomicAdd(&s_grad[d.dslot + e], v);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sp[pb ^ G.off[j]] = a[j];
      sl[pb ^ G.off[j]] = l[j];
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

template <bool BWD>
__device__ __forceinline__ void rg_stream(cf* sp, cf* sl, const Ring<float>& ring, const StreamRef& st, const cf* pay_b,
                                          float* s_grad, int m) {
  ring_start<float>(ring, st, pay_b);
  for (int c = 0; c < st.n_chunks; ++c) {
    cp_async_wait_all();
    __syncthreads();  // chunk c landed; everyone is done with the buffer chunk c+1 will overwrite
    if (c + 1 < st.n_chunks) ring_issue<float>(ring, st, pay_b, c + 1);
    const ChunkInfo ci = chunk_info<float>(ring, st, c);
    const OpDesc* dd = ring.desc[c & 1];
    const cf* pp = ring.pay[c & 1];
    for (uint32_t o = 0; o < ci.op_count;) {
      const OpDesc& h = dd[o];
      if (BWD)
        rg_run_bwd(sp, sl, h, dd + o + 1, pp, ci.pay_begin, s_grad, m);
      else
        rg_run_fwd(sp, h, dd + o + 1, pp, ci.pay_begin, m);
      o += 1u + h.nins;
      __syncthreads();
    }
  }
}

// ---- kernels (same arguments, flags and shared-memory layout as k_sweep_fwd / k_sweep_bwd) --------------------------
__global__ void __launch_bounds__(256, 3) k_rg_fwd(const __grid_constant__ FwdArgs<float> a) {
  const int m = a.geom.m;
  const uint32_t tile_n = 1u << m;
  cf* sm = reinterpret_cast<cf*>(tq_smem);
  Ring<float> ring = ring_carve<float>(tq_smem + sizeof(cf) * tile_n);
  const int64_t b = (int64_t)(blockIdx.x >> a.tiles_log2);
  const uint32_t tile = blockIdx.x & ((1u << a.tiles_log2) - 1u);
  const uint32_t tbase = dep_tile(a.geom, tile);
  const size_t sv = (size_t)1 << a.geom.n;
  cf* psi_b = a.psi ? a.psi + (size_t)b * sv : nullptr;

  if (a.flags & SW_INIT) {
    if (a.init_state) {
      for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sm[rg_phys(l)] = a.init_state[tbase | dep_local(a.geom, l)];
    } else {
      for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sm[l] = mk<float>(0.f, 0.f);
      __syncthreads();
      if (threadIdx.x == 0 && tbase == 0) sm[0] = mk<float>(1.f, 0.f);  // rg_phys(0) = 0
    }
  } else {
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) sm[rg_phys(l)] = psi_b[tbase | dep_local(a.geom, l)];
  }
  rg_stream<false>(sm, nullptr, ring, a.st, a.stream + b * a.stride, nullptr, m);  // begins and ends with a barrier

  if (a.flags & SW_STORE) {
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) psi_b[tbase | dep_local(a.geom, l)] = sm[rg_phys(l)];
  }
  if (a.flags & SW_MEASURE) {
    __syncthreads();
    rg_swizzle_tile(sm, tile_n);  // back to natural order for the measurement code
    float* s_acc = reinterpret_cast<float*>(tq_smem + sizeof(cf) * tile_n + RING_BYTES);
    for (int s = threadIdx.x; s < a.n_slots; s += blockDim.x) s_acc[s] = 0;
    __syncthreads();
    measure_block<float>(sm, 0, tile_n, a.meas, a.n_meas, a.fixed, a.out + (size_t)b * a.out_reals, s_acc, false);
  }
}

__global__ void __launch_bounds__(256, 2) k_rg_bwd(const __grid_constant__ BwdArgs<float> a) {
  const int m = a.geom.m;
  const uint32_t tile_n = 1u << m;
  cf* sp = reinterpret_cast<cf*>(tq_smem);
  cf* sl = sp + tile_n;
  Ring<float> ring = ring_carve<float>(tq_smem + 2 * sizeof(cf) * tile_n);
  float* s_grad = reinterpret_cast<float*>(tq_smem + 2 * sizeof(cf) * tile_n + RING_BYTES);
  const int64_t b = (int64_t)(blockIdx.x >> a.tiles_log2);
  const uint32_t tile = blockIdx.x & ((1u << a.tiles_log2) - 1u);
  const uint32_t tbase = dep_tile(a.geom, tile);
  const size_t sv = (size_t)1 << a.geom.n;
  cf* psi_b = a.psi + (size_t)b * sv;  // the forward pass (with_backward) left the final state here

  for (int s = threadIdx.x; s < a.n_dslots; s += blockDim.x) s_grad[s] = 0;

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
    const cf* lam_b = a.lam + (size_t)b * sv;
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) {
      const uint32_t gi = tbase | dep_local(a.geom, l);
      const uint32_t p = rg_phys(l);
      sp[p] = psi_b[gi];
      sl[p] = lam_b[gi];
    }
  }
  rg_stream<true>(sp, sl, ring, a.st_b, a.stream_b + b * a.stride_b, s_grad, m);  // begins and ends with a barrier

  if (!(a.flags & SW_FULL) && (a.flags & SW_STORE)) {
    cf* lam_b = a.lam + (size_t)b * sv;
    for (uint32_t l = threadIdx.x; l < tile_n; l += blockDim.x) {
      const uint32_t gi = tbase | dep_local(a.geom, l);
      const uint32_t p = rg_phys(l);
      psi_b[gi] = sp[p];
      lam_b[gi] = sl[p];
    }
  }
  float* grad_b = a.grad + b * a.n_params;
  for (int s = threadIdx.x; s < a.n_dslots; s += blockDim.x) {
    if (a.flags & SW_FULL)
      grad_b[a.slot_pidx[s]] = s_grad[s];
    else
      atomicAdd(&grad_b[a.slot_pidx[s]], s_grad[s]);
  }
}
#endif  // __CUDACC__

}  // namespace tq
