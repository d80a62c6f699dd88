// Host-side check of the register-group arithmetic and index maps (tq_sv_rg.cuh): the same inline functions the
// kernels call, run on the CPU over a whole swizzled tile and compared with a gate-by-gate reference on the natural
// tile.  Built and run by tests/test_rg_host_cpu.py (nvcc, host code only — no GPU needed).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <complex>
#include <vector>

#include "../../ted-q_b200/csrc/tq_sv_rg.cuh"

using namespace tq;
typedef std::complex<double> zc;
typedef cx<float> cf32;

static double urand() { return rand() / (double)RAND_MAX * 2.0 - 1.0; }

struct Blk {
  int cls;                      // OP_DENSE (1 target) or OP_DIAG (1..3 targets)
  std::vector<int> targets;     // tile bits, MSB of the matrix index first
  std::vector<int> controls;    // tile bits
  bool is_x;
  int nderiv;
  std::vector<cf32> pay;        // matrix (4) / diagonal (2^k), then nderiv derivative matrices of the same size
};

static zc Z(cf32 v) { return zc(v.x, v.y); }

// reference: apply block (or its conjugate transpose) to a natural-order tile
static void ref_apply(std::vector<zc>& s, int m, const Blk& b, bool adj) {
  const uint32_t n = 1u << m;
  uint32_t cm = 0;
  for (int c : b.controls) cm |= 1u << c;
  if (b.cls == OP_DIAG) {
    for (uint32_t i = 0; i < n; ++i) {
      if ((i & cm) != cm) continue;
      int idx = 0;
      for (int t : b.targets) idx = (idx << 1) | ((i >> t) & 1);
      zc v = Z(b.pay[idx]);
      s[i] *= adj ? std::conj(v) : v;
    }
    return;
  }
  const uint32_t tb = 1u << b.targets[0];
  zc M[4] = {Z(b.pay[0]), Z(b.pay[1]), Z(b.pay[2]), Z(b.pay[3])};
  if (b.is_x) { M[0] = 0; M[1] = 1; M[2] = 1; M[3] = 0; }
  if (adj) {
    zc T[4] = {std::conj(M[0]), std::conj(M[2]), std::conj(M[1]), std::conj(M[3])};
    for (int i = 0; i < 4; ++i) M[i] = T[i];
  }
  for (uint32_t i = 0; i < n; ++i) {
    if ((i & tb) || (i & cm) != cm) continue;
    zc a0 = s[i], a1 = s[i | tb];
    s[i] = M[0] * a0 + M[1] * a1;
    s[i | tb] = M[2] * a0 + M[3] * a1;
  }
}

// reference gradient term of slot e: Re <lambda | dG_e | psi_prev> over the block's live amplitudes
static double ref_grad(const std::vector<zc>& psi_prev, const std::vector<zc>& lam, int m, const Blk& b, int e) {
  const uint32_t n = 1u << m;
  uint32_t cm = 0;
  for (int c : b.controls) cm |= 1u << c;
  const uint32_t tb = 1u << b.targets[0];
  double g = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if ((i & tb) || (i & cm) != cm) continue;
    zc D[4];
    if (b.cls == OP_DIAG) {
      D[0] = Z(b.pay[2 + 2 * e]); D[1] = 0; D[2] = 0; D[3] = Z(b.pay[3 + 2 * e]);
    } else {
      for (int k = 0; k < 4; ++k) D[k] = Z(b.pay[4 + 4 * e + k]);
    }
    zc p0 = psi_prev[i], p1 = psi_prev[i | tb];
    g += std::real(std::conj(lam[i]) * (D[0] * p0 + D[1] * p1) + std::conj(lam[i | tb]) * (D[2] * p0 + D[3] * p1));
  }
  return g;
}

static RgSub decode(const OpDesc& d) {
  uint4 w[2];
  memcpy(w, &d, 32);
  return rg_decode(w[0], w[1]);
}

static int fails = 0;
#define CHECK(cond, ...)          \
  do {                            \
    if (!(cond)) {                \
      printf("FAIL: " __VA_ARGS__); \
      printf("\n");               \
      ++fails;                    \
    }                             \
  } while (0)

int main() {
  srand(1234);
  // 1. swizzle: bijection + involution inside aligned 16-word blocks
  for (uint32_t i = 0; i < (1u << 14); ++i) {
    const uint32_t p = rg_phys(i);
    CHECK((p >> 4) == (i >> 4), "rg_phys leaves the 16-word block");
    CHECK(rg_phys(p) == i, "rg_phys is not an involution at %u", i);
  }
  // 2. bank conflicts: contiguous register bits are conflict free for every tile size; padded picks are too
  for (int m = RG_MIN_TILE; m <= 14; ++m) {
    for (int b0 = 0; b0 + 4 <= m; ++b0) {
      uint32_t regbits = 0, o[4];
      for (int i = 0; i < 4; ++i) regbits |= (uint32_t)(b0 + i) << (8 * i);
      rg_bit_offsets(regbits, o);
      const RgAddr G = rg_addr(o, RG_MAP_ID);
      for (int j = 0; j < 16; ++j)
        for (uint32_t w = 0; w < (1u << (m - 4)); w += 16) {  // a half-warp: 16 consecutive groups
          int seen = 0;
          for (uint32_t l = 0; l < 16; ++l) {
            const uint32_t addr = rg_base(regbits, w + l) ^ RG_OFF(G, j);
            CHECK(addr < (1u << m), "address outside the tile");
            seen |= 1 << (addr & 15);
          }
          CHECK(seen == 0xffff, "bank conflict m=%d b0=%d j=%d", m, b0, j);
        }
    }
    for (int trial = 0; trial < 200; ++trial) {
      int nu = 1 + rand() % 4, used[4];
      for (int i = 0; i < nu;) {
        int b = rand() % m;
        bool dup = false;
        for (int k = 0; k < i; ++k) dup = dup || used[k] == b;
        if (!dup) used[i++] = b;
      }
      int reg[4];
      rg_pick_bits(m, used, nu, reg);
      for (int i = 0; i < 3; ++i) CHECK(reg[i] < reg[i + 1], "register bits not ascending");
      for (int i = 0; i < nu; ++i) {
        bool in = false;
        for (int k = 0; k < 4; ++k) in = in || reg[k] == used[i];
        CHECK(in, "used bit dropped");
      }
      CHECK(reg[3] < m, "register bit outside the tile");
    }
  }
  // 3. forward + adjoint of random groups on random tiles
  for (int trial = 0; trial < 300; ++trial) {
    const int m = RG_MIN_TILE + rand() % 3;
    const uint32_t n = 1u << m;
    // register bits: random 4 of m
    int reg[4];
    {
      int nu = 4, used[4];
      for (int i = 0; i < nu;) {
        int b = rand() % m;
        bool dup = false;
        for (int k = 0; k < i; ++k) dup = dup || used[k] == b;
        if (!dup) used[i++] = b;
      }
      rg_pick_bits(m, used, 4, reg);
    }
    const int nsub = 1 + rand() % 8;
    std::vector<Blk> blks;
    std::vector<OpDesc> descs;
    std::vector<cf32> payload;
    OpDesc h;
    memset(&h, 0, sizeof(h));
    h.path = P_RG;
    h.nins = (uint8_t)nsub;
    for (int i = 0; i < 4; ++i) h.tpos[i] = (uint8_t)reg[i];
    for (int sidx = 0; sidx < nsub; ++sidx) {
      Blk b;
      const int kind = rand() % 5;  // 0 dense, 1 diag-1, 2 controlled dense, 3 x (0..2 controls), 4 multi-target diag
      int perm[4] = {0, 1, 2, 3};
      for (int i = 3; i > 0; --i) {
        int j = rand() % (i + 1);
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
      }
      b.is_x = false;
      b.nderiv = 0;
      int nt = 1, nc = 0;
      b.cls = OP_DENSE;
      if (kind == 1) b.cls = OP_DIAG;
      if (kind == 2) nc = 1 + rand() % 2;
      if (kind == 3) { b.is_x = true; nc = rand() % 3; }
      if (kind == 4) { b.cls = OP_DIAG; nt = 2 + rand() % 2; nc = rand() % (5 - nt); }
      if (kind <= 2) b.nderiv = rand() % 3;
      if (kind == 1 && (rand() & 1)) nc = 1;
      int treg[4], creg[4];
      for (int i = 0; i < nt; ++i) { treg[i] = perm[i]; b.targets.push_back(reg[perm[i]]); }
      for (int i = 0; i < nc; ++i) { creg[i] = perm[nt + i]; b.controls.push_back(reg[perm[nt + i]]); }
      const int count = b.cls == OP_DIAG ? (1 << nt) : 4;
      for (int i = 0; i < count * (1 + b.nderiv); ++i) b.pay.push_back(mk<float>((float)urand(), (float)urand()));
      if (b.is_x) { b.pay[0] = mk<float>(0, 0); b.pay[1] = mk<float>(1, 0); b.pay[2] = mk<float>(1, 0); b.pay[3] = mk<float>(0, 0); }
      OpDesc d;
      rg_make_sub(b.cls, nt, treg, nc, creg, b.is_x, count, b.nderiv, d);
      d.pay_off = (uint32_t)payload.size();
      payload.insert(payload.end(), b.pay.begin(), b.pay.end());
      if (payload.size() & 1) payload.push_back(mk<float>(0, 0));
      blks.push_back(b);
      descs.push_back(d);
    }
    // tiles
    std::vector<zc> ref(n), lam_ref(n);
    std::vector<cf32> sp(n), sl(n);
    for (uint32_t i = 0; i < n; ++i) {
      cf32 v = mk<float>((float)urand(), (float)urand()), w = mk<float>((float)urand(), (float)urand());
      ref[i] = Z(v);
      lam_ref[i] = Z(w);
      sp[rg_phys(i)] = v;
      sl[rg_phys(i)] = w;
    }
    // X / CNOT networks folded into the load (pre) and store (post) addresses
    std::vector<Blk> pre, post;
    int pre_t[8], pre_c[8], post_t[8], post_c[8];
    const int npre = rand() % 5, npost = rand() % 5;
    for (int side = 0; side < 2; ++side)
      for (int i = 0; i < (side ? npost : npre); ++i) {
        const int t = rand() % 4;
        int c = rand() % 5 - 1;  // -1: plain X
        if (c == t) c = -1;
        (side ? post_t : pre_t)[i] = t;
        (side ? post_c : pre_c)[i] = c;
        Blk b;
        b.cls = OP_DENSE;
        b.is_x = true;
        b.nderiv = 0;
        b.targets.push_back(reg[t]);
        if (c >= 0) b.controls.push_back(reg[c]);
        b.pay.assign(4, mk<float>(0, 0));
        (side ? post : pre).push_back(b);
      }
    uint32_t regbits = 0, o[4];
    for (int i = 0; i < 4; ++i) regbits |= (uint32_t)reg[i] << (8 * i);
    rg_bit_offsets(regbits, o);
    const RgAddr LF = rg_addr(o, rg_affine_map(pre_t, pre_c, npre, false)), SF = rg_addr(o, rg_affine_map(post_t, post_c, npost, true));
    std::vector<RgSub> subs;
    for (int i = 0; i < nsub; ++i) subs.push_back(decode(descs[i]));
    // forward on sp
    for (uint32_t g = 0; g < (n >> 4); ++g) {
      const uint32_t pb = rg_base(regbits, g);
      cf32 a[16];
      for (int j = 0; j < 16; ++j) a[j] = sp[pb ^ RG_OFF(LF, j)];
      for (int i = 0; i < nsub; ++i) {
        cf32 mm[4];
        rg_ld2x2<false>(payload.data() + subs[i].pay_off, subs[i].count, mm);
        rg_fwd_sub(a, subs[i], mm, payload.data() + subs[i].pay_off);
      }
      for (int j = 0; j < 16; ++j) sp[pb ^ RG_OFF(SF, j)] = a[j];
    }
    for (const Blk& b : pre) ref_apply(ref, m, b, false);
    for (int i = 0; i < nsub; ++i) ref_apply(ref, m, blks[i], false);
    for (const Blk& b : post) ref_apply(ref, m, b, false);
    double err = 0, nrm = 0;
    for (uint32_t i = 0; i < n; ++i) {
      err = fmax(err, std::abs(Z(sp[rg_phys(i)]) - ref[i]));
      nrm = fmax(nrm, std::abs(ref[i]));
    }
    CHECK(err <= 2e-5 * nrm, "forward trial %d: err %g (max %g)", trial, err, nrm);
    // adjoint: the same descriptors in reverse order on (sp, sl); reference un-applies block by block
    std::vector<double> grad(64, 0.0), grad_ref(64, 0.0);
    std::vector<int> slot0(nsub);
    int ns = 0;
    for (int i = 0; i < nsub; ++i) { slot0[i] = ns; ns += blks[i].nderiv; }
    // the adjoint group: the blocks in reverse order; its pre network is the forward post network reversed
    int bpre_t[8], bpre_c[8], bpost_t[8], bpost_c[8];
    for (int i = 0; i < npost; ++i) { bpre_t[i] = post_t[npost - 1 - i]; bpre_c[i] = post_c[npost - 1 - i]; }
    for (int i = 0; i < npre; ++i) { bpost_t[i] = pre_t[npre - 1 - i]; bpost_c[i] = pre_c[npre - 1 - i]; }
    const RgAddr LB = rg_addr(o, rg_affine_map(bpre_t, bpre_c, npost, false)), SB = rg_addr(o, rg_affine_map(bpost_t, bpost_c, npre, true));
    for (uint32_t g = 0; g < (n >> 4); ++g) {
      const uint32_t pb = rg_base(regbits, g);
      cf32 a[16], l[16];
      for (int j = 0; j < 16; ++j) { a[j] = sp[pb ^ RG_OFF(LB, j)]; l[j] = sl[pb ^ RG_OFF(LB, j)]; }
      for (int i = nsub - 1; i >= 0; --i) {
        cf32 W[4] = {mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0)}, mh[4];
        const cf32* pay = payload.data() + subs[i].pay_off;
        rg_ld2x2<true>(pay, subs[i].count, mh);
        if (rg_bwd_sub(a, l, subs[i], mh, pay, W))
          for (uint32_t e = 0; e < subs[i].nderiv; ++e)
            grad[slot0[i] + e] += rg_grad_term(W, pay + (subs[i].count == 2 ? 2 + 2 * e : 4 + 4 * e), subs[i].count);
      }
      for (int j = 0; j < 16; ++j) { sp[pb ^ RG_OFF(SB, j)] = a[j]; sl[pb ^ RG_OFF(SB, j)] = l[j]; }
    }
    for (int i = npost - 1; i >= 0; --i) { ref_apply(ref, m, post[i], true); ref_apply(lam_ref, m, post[i], true); }
    for (int i = nsub - 1; i >= 0; --i) {
      ref_apply(ref, m, blks[i], true);  // psi_prev
      for (int e = 0; e < blks[i].nderiv; ++e) grad_ref[slot0[i] + e] = ref_grad(ref, lam_ref, m, blks[i], e);
      ref_apply(lam_ref, m, blks[i], true);
    }
    for (int i = npre - 1; i >= 0; --i) { ref_apply(ref, m, pre[i], true); ref_apply(lam_ref, m, pre[i], true); }
    double e1 = 0, e2 = 0, n1 = 0, n2 = 0;
    for (uint32_t i = 0; i < n; ++i) {
      e1 = fmax(e1, std::abs(Z(sp[rg_phys(i)]) - ref[i]));
      e2 = fmax(e2, std::abs(Z(sl[rg_phys(i)]) - lam_ref[i]));
      n1 = fmax(n1, std::abs(ref[i]));
      n2 = fmax(n2, std::abs(lam_ref[i]));
    }
    CHECK(e1 <= 1e-4 * n1 && e2 <= 1e-4 * n2, "adjoint trial %d: err %g %g (max %g %g)", trial, e1, e2, n1, n2);
    for (int sidx = 0; sidx < ns; ++sidx) {
      double scale = fmax(1.0, fabs(grad_ref[sidx]));
      CHECK(fabs(grad[sidx] - grad_ref[sidx]) <= 2e-3 * scale * sqrt((double)n) / 16.0 + 1e-3 * scale, "gradient trial %d slot %d: %g vs %g", trial, sidx, grad[sidx],
            grad_ref[sidx]);
    }
  }
  // 3b. generator gradients: blocks U = L * M * F with Pauli rotations as first / last member against the derivative
  //     matrices of the same block
  {
    auto rot = [](int P, double t, zc* M) {  // exp(-i t P / 2); P = 4: diag(1, e^{it})
      const double c = cos(t / 2), sn = sin(t / 2);
      const zc I(0, 1);
      if (P == RG_PX) { M[0] = c; M[1] = -I * sn; M[2] = -I * sn; M[3] = c; }
      if (P == RG_PY) { M[0] = c; M[1] = -sn; M[2] = sn; M[3] = c; }
      if (P == RG_PZ) { M[0] = std::exp(-I * (t / 2)); M[1] = 0; M[2] = 0; M[3] = std::exp(I * (t / 2)); }
      if (P == RG_PP) { M[0] = 1; M[1] = 0; M[2] = 0; M[3] = std::exp(I * t); }
    };
    auto gen = [](int P, zc* G) {  // dR/dt = G R
      const zc I(0, 1);
      for (int i = 0; i < 4; ++i) G[i] = 0;
      if (P == RG_PX) { G[1] = -I * 0.5; G[2] = -I * 0.5; }
      if (P == RG_PY) { G[1] = -0.5; G[2] = 0.5; }
      if (P == RG_PZ) { G[0] = -I * 0.5; G[3] = I * 0.5; }
      if (P == RG_PP) { G[3] = I; }
    };
    auto mul = [](const zc* A, const zc* B, zc* C) {
      zc T[4] = {A[0] * B[0] + A[1] * B[2], A[0] * B[1] + A[1] * B[3], A[2] * B[0] + A[3] * B[2], A[2] * B[1] + A[3] * B[3]};
      for (int i = 0; i < 4; ++i) C[i] = T[i];
    };
    for (int trial = 0; trial < 200; ++trial) {
      const int m = RG_MIN_TILE;
      const uint32_t n = 1u << m;
      int reg[4] = {1, 4, 6, 7};
      const int pf = rand() % 5, pl = rand() % 5;          // 0: that end is not trainable
      if (!pf && !pl) continue;
      const bool single = pf == 0 && (rand() & 1);         // one gate: first == last
      zc F[4], L[4], Mid[4], U[4], dF[4], dL[4], G[4], T[4];
      const double tf = urand() * 3, tl = urand() * 3;
      rot(pf ? pf : RG_PX, tf, F);
      rot(pl ? pl : RG_PY, tl, L);
      {  // a unitary middle part (the identity U U^dagger = 1 is what makes the output-side formula exact)
        zc A[4], B[4], C[4];
        rot(RG_PX, urand() * 3, A); rot(RG_PZ, urand() * 3, B); rot(RG_PY, urand() * 3, C);
        mul(B, A, Mid); mul(C, Mid, Mid);
        if (single) for (int i = 0; i < 4; ++i) Mid[i] = zc(i == 0 || i == 3, 0);
      }
      if (single) for (int i = 0; i < 4; ++i) F[i] = zc(i == 0 || i == 3, 0);
      mul(Mid, F, T);
      mul(L, T, U);
      Blk b;
      b.cls = OP_DENSE;
      b.is_x = false;
      const int t = rand() % 4;
      b.targets.push_back(reg[t]);
      int nc = rand() % 2, creg[1] = {(t + 1) % 4}, treg[1] = {t};
      if (nc) b.controls.push_back(reg[creg[0]]);
      b.nderiv = (pf && !single ? 1 : 0) + (pl ? 1 : 0);
      for (int i = 0; i < 4; ++i) b.pay.push_back(mk<float>((float)U[i].real(), (float)U[i].imag()));
      if (pf && !single) {  // dU = L Mid (G F)
        gen(pf, G); mul(G, F, dF); mul(Mid, dF, T); mul(L, T, T);
        for (int i = 0; i < 4; ++i) b.pay.push_back(mk<float>((float)T[i].real(), (float)T[i].imag()));
      }
      if (pl) {
        gen(pl, G); mul(G, U, dL);
        for (int i = 0; i < 4; ++i) b.pay.push_back(mk<float>((float)dL[i].real(), (float)dL[i].imag()));
      }
      const uint32_t code = (uint32_t)((pf && !single ? pf : 0) | ((pl ? pl : 0) << 4));
      if (!b.nderiv) continue;
      OpDesc d;
      rg_make_sub(b.cls, 1, treg, nc, creg, false, 4, b.nderiv, d, code);
      const RgSub sub = decode(d);
      std::vector<zc> psi(n), lam(n);
      std::vector<cf32> sp(n), sl(n);
      for (uint32_t i = 0; i < n; ++i) {
        cf32 v = mk<float>((float)urand(), (float)urand()), w = mk<float>((float)urand(), (float)urand());
        psi[i] = Z(v); lam[i] = Z(w);
        sp[rg_phys(i)] = v; sl[rg_phys(i)] = w;
      }
      uint32_t regbits = 0, o[4];
      for (int i = 0; i < 4; ++i) regbits |= (uint32_t)reg[i] << (8 * i);
      rg_bit_offsets(regbits, o);
      const RgAddr A = rg_addr(o, RG_MAP_ID);
      double g[2] = {0, 0};
      for (uint32_t gi = 0; gi < (n >> 4); ++gi) {
        const uint32_t pb = rg_base(regbits, gi);
        cf32 a[16], l[16], W[4] = {mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0)}, mh[4];
        for (int j = 0; j < 16; ++j) { a[j] = sp[pb ^ RG_OFF(A, j)]; l[j] = sl[pb ^ RG_OFF(A, j)]; }
        rg_ld2x2<true>(b.pay.data(), 4, mh);
        CHECK(rg_bwd_sub(a, l, sub, mh, b.pay.data(), W), "generator sub-op reports no gradient");
        g[0] += W[0].x;
        g[1] += W[0].y;
      }
      // reference: Re <lambda | dU_e | psi_prev>, psi_prev = U^dagger psi, with the block's own derivative matrices
      std::vector<zc> prev = psi;
      ref_apply(prev, m, b, true);
      for (int e = 0; e < b.nderiv; ++e) {
        const double r = ref_grad(prev, lam, m, b, e);
        CHECK(fabs(g[e] - r) <= 2e-3 * fmax(1.0, fabs(r)), "generator gradient trial %d slot %d: %g vs %g (pf %d pl %d single %d)", trial,
              e, g[e], r, pf, pl, (int)single);
      }
    }
  }
  // 4. grouping of a hardware-efficient ansatz inside one tile: every block lands in exactly one group, blocks that
  //    share a bit keep their order, and the groups hold close to four one-qubit blocks each
  for (int m = 9; m <= 14; ++m) {
    std::vector<RgItem> items;
    auto add = [&](int b0, int b1, bool fold) {
      RgItem it;
      it.nbits = b1 < 0 ? 1 : 2;
      it.bits[0] = b0;
      it.bits[1] = b1;
      it.foldable = fold;
      it.pay = fold ? 4 : 12;
      items.push_back(it);
    };
    const int depth = 10;
    for (int d = 0; d < depth; ++d) {
      for (int q = 0; q < m; ++q) add(m - 1 - q, -1, false);
      for (int first = 2; first >= 1; --first)
        for (int w = first; w < m; w += 2) add(m - 1 - (w - 1), m - 1 - w, true);
    }
    for (int q = 0; q < m; ++q) add(m - 1 - q, -1, false);
    std::vector<int> rest(items.size()), stamp(items.size(), -1);
    for (size_t i = 0; i < items.size(); ++i) rest[i] = (int)i;
    int groups = 0, d1 = 0, empty_groups = 0, clock = 0;
    while (!rest.empty()) {
      std::vector<int> pre, mid, post;
      std::vector<char> inb;
      const size_t before = rest.size();
      rg_next_group(items, rest, m, 256, pre, mid, post, inb);
      CHECK(pre.size() + mid.size() + post.size() + rest.size() == before, "grouping lost a block");
      CHECK(pre.size() + mid.size() + post.size() > 0, "grouping made no progress");
      if (pre.size() + mid.size() + post.size() == 0) break;
      int nb = 0;
      for (int b = 0; b < m; ++b) nb += inb[b];
      CHECK(nb <= 4, "group spans %d bits", nb);
      for (const std::vector<int>* part : {&pre, &mid, &post})
        for (int bi : *part) {
          stamp[bi] = clock++;
          for (int k = 0; k < items[bi].nbits; ++k) CHECK(inb[items[bi].bits[k]], "block bit outside the group");
        }
      ++groups;
      d1 += (int)mid.size();
      empty_groups += mid.empty();
    }
    for (size_t i = 0; i < items.size(); ++i)
      for (size_t j = i + 1; j < items.size(); ++j) {
        bool share = false;
        for (int a = 0; a < items[i].nbits; ++a)
          for (int b = 0; b < items[j].nbits; ++b) share = share || items[i].bits[a] == items[j].bits[b];
        if (share) CHECK(stamp[i] < stamp[j], "blocks %zu and %zu share a bit and were reordered", i, j);
      }
    printf("hea m=%d depth=%d: %d groups (%d without a one-qubit block), %.2f one-qubit blocks per group\n", m, depth, groups,
           empty_groups, d1 / (double)groups);
    CHECK(d1 / (double)groups >= 2.5, "groups too small");
  }
  // 5. grouping of random block lists (1-3 bits per block, X / CNOT-like blocks mixed in): the same invariants
  for (int trial = 0; trial < 300; ++trial) {
    const int m = RG_MIN_TILE + rand() % 6;
    const int nblk = 20 + rand() % 200;
    std::vector<RgItem> items(nblk);
    for (RgItem& it : items) {
      it.nbits = 1 + rand() % 3;
      for (int k = 0; k < it.nbits;) {
        const int b = rand() % m;
        bool dup = false;
        for (int j = 0; j < k; ++j) dup = dup || it.bits[j] == b;
        if (!dup) it.bits[k++] = b;
      }
      it.foldable = it.nbits <= 2 && rand() % 3 == 0;
      it.pay = 4 + 4 * (rand() % 4);
    }
    std::vector<int> rest(nblk), stamp(nblk, -1);
    for (int i = 0; i < nblk; ++i) rest[i] = i;
    int clock = 0;
    while (!rest.empty()) {
      std::vector<int> pre, mid, post;
      std::vector<char> inb;
      const size_t before = rest.size();
      rg_next_group(items, rest, m, 64, pre, mid, post, inb);
      if (pre.size() + mid.size() + post.size() + rest.size() != before || pre.size() + mid.size() + post.size() == 0) {
        CHECK(false, "random grouping trial %d lost a block or made no progress", trial);
        break;
      }
      int nb = 0, pay = 0;
      for (int b = 0; b < m; ++b) nb += inb[b];
      CHECK(nb <= 4 && (int)mid.size() <= RG_MAX_SUB, "random grouping: group spans %d bits, %zu sub-ops", nb, mid.size());
      for (int bi : mid) pay += items[bi].pay;
      CHECK(pay <= 64, "random grouping: payload %d over the cap", pay);
      for (int bi : pre) CHECK(items[bi].foldable, "non-foldable block in the pre phase");
      for (int bi : post) CHECK(items[bi].foldable, "non-foldable block in the post phase");
      for (const std::vector<int>* part : {&pre, &mid, &post})
        for (int bi : *part) stamp[bi] = clock++;
    }
    for (int i = 0; i < nblk; ++i)
      for (int j = i + 1; j < nblk; ++j) {
        bool share = false;
        for (int a = 0; a < items[i].nbits; ++a)
          for (int b = 0; b < items[j].nbits; ++b) share = share || items[i].bits[a] == items[j].bits[b];
        if (share) CHECK(stamp[i] >= 0 && stamp[i] < stamp[j], "random grouping trial %d: blocks %d and %d share a bit and were reordered", trial, i, j);
      }
  }
  if (fails) {
    printf("rg_check: %d failures\n", fails);
    return 1;
  }
  printf("rg_check: ok\n");
  return 0;
}
