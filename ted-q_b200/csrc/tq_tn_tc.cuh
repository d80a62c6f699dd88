// tq_tn_tc.cuh — tensor-core path of the contraction executor (complex64 steps, sm_100a only).
//
// Replaces the arithmetic of the third-party tree.contract call (pytorch_backend.py:276,:339 hand every
// pairwise step to torch.tensordot = cgemm) for the GEMM-shaped steps of a plan.
//
// One complex GEMM  C[r, c] = sum_k A[r, k] * B[c, k]  (A = the operand with more free indices) is run as ONE
// real TF32 GEMM on tcgen05 with fp32 accumulation in TMEM:
//
//   A'' [R][2K]   row r        = (Re A[r,0], Im A[r,0], Re A[r,1], Im A[r,1], ...)          (native complex layout)
//   B'' [2C][2K]  row c        = (Re B, -Im B, ...)   -> accumulator column c       = Re C[r, c]
//                 row C_t + c  = (Im B,  Re B, ...)   -> accumulator column C_t + c = Im C[r, c]       ("4M")
//
// and every fp32 operand is split  x = hi + lo  (hi = tf32(x), lo = tf32(x - hi)); the accumulator receives
// hi*hi + hi*lo + lo*hi  (error-compensated split-TF32: the dropped lo*lo term is 2^-22 relative).  A complex
// MAC therefore costs 3 x 8 = 24 TF32 flops for 8 algorithmic flops.
//
// Data movement: a pack kernel folds the step's bit permutation, the hi/lo split and the 128-byte shared-memory
// swizzle into "operand images" — every (row tile, k block) is one contiguous chunk in HBM that is exactly the
// shared-memory image the MMA wants, so the GEMM kernel's producer is two bulk-async (TMA engine) copies per
// stage, no tensor maps.  GEMM kernel: persistent, warp-specialised (bulk-copy producer / single-thread MMA
// issuer / 8 accumulation warps that drain TMEM into fp32 registers and write C), 2-4 smem stages,
// double-buffered TMEM accumulator (2 x 256 columns).  An opt-in variant (k_tc_gemm<C_T, true>) gathers and splits
// the row operand itself instead of reading its image (see the comment above the kernel).
#pragma once
#include "tq_common.h"

namespace tq {
namespace tc {

constexpr int ROWS = 128;                 // accumulator rows per tile (TMEM lanes)
constexpr int ACC_WARPS_MAX = 8;
#ifndef TQ_TC_KB_LOG
#define TQ_TC_KB_LOG 3
#endif
// A k-block is one swizzle row of an operand tile: 2^KB_LOG complex k.  KB_LOG 4 = 128-byte rows (SWIZZLE_128B,
// 96 KiB per stage at 128 columns: 2 stages); KB_LOG 3 = 64-byte rows (SWIZZLE_64B, 48 KiB per stage: 4 stages,
// which covers the fill latency of a stage with three others in flight).
constexpr int KB_LOG = TQ_TC_KB_LOG;
constexpr int KB_CPLX = 1 << KB_LOG;
constexpr int ROW_BYTES = 8 * KB_CPLX;
constexpr int K_STEPS = ROW_BYTES / 32;   // tcgen05 kind::tf32 consumes 8 reals (32 B) of K per instruction
constexpr int A_PLANE = ROWS * ROW_BYTES; // one plane (hi or lo) of an A tile
constexpr int A_CHUNK = 2 * A_PLANE;      // hi + lo
constexpr int ACC_WARPS = 8;              // accumulation / epilogue warps (two per TMEM lane quadrant)
constexpr int GEMM_THREADS = 64 + 32 * ACC_WARPS;  // warp 0 producer, warp 1 MMA issuer, warps 2..9 accumulate
// gather-A variant: 4 warpgroups.  WG0 = B bulk-copy producer (warp 0), MMA issuer (warp 1), A loaders (warps 2-3);
// WG1-2 = the 8 accumulation warps; WG3 = 4 converter warps.  Registers are re-balanced with setmaxnreg.
constexpr int GA_ACC0 = 4;                // first accumulation warp
constexpr int GA_CONV0 = 12;              // first converter warp
constexpr int CONV_WARPS = 4;
constexpr int CONV_ITERS = ROWS * KB_CPLX / (32 * CONV_WARPS);  // complex entries of a k-block per converter thread
constexpr int LOAD_ITERS = ROWS * KB_CPLX / 64;                 // ... per loader thread (2 loader warps)
constexpr int GEMM_THREADS_GA = 512;
constexpr int SMEM_BUDGET = 200 * 1024;    // operand stages
constexpr int EPI_PITCH = 80;             // bytes per staged row piece (64 data + 16 pad: conflict-free quarter-warps)
constexpr int EPI_WARP_BYTES = 32 * EPI_PITCH;
constexpr int EPI_BYTES = ACC_WARPS_MAX * EPI_WARP_BYTES;  // epilogue staging, after the operand stages
constexpr int GEMM_SMEM = SMEM_BUDGET + EPI_BYTES + 1024;

__host__ __device__ inline int b_plane_bytes(int c_t) { return 2 * c_t * ROW_BYTES; }  // Re rows + Im rows
__host__ __device__ inline int b_chunk_bytes(int c_t) { return 2 * b_plane_bytes(c_t); }
__host__ __device__ inline int stage_bytes(int c_t) { return A_CHUNK + b_chunk_bytes(c_t); }
inline int num_stages(int c_t) {
  int s = SMEM_BUDGET / stage_bytes(c_t);
  return s > 4 ? 4 : s;
}

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a launch error, never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (uint32_t spin = 1; !mbar_try_wait(bar, parity); ++spin)
    if ((spin & 1023u) == 0 && clock64() - t0 > 6000000000ll) __trap();  // ~3 s
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, rows of ROW_BYTES, matching swizzle: 8-row groups (SBO = 8 rows), descriptor version 1,
// layout type 2 = SWIZZLE_128B / 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  constexpr uint64_t sbo = (8 * ROW_BYTES) >> 4, layout = ROW_BYTES == 128 ? 2 : 4;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// 16-byte chunk c of row r sits at chunk c ^ swz(r): Swizzle<3,4,3> (128 B rows) / Swizzle<2,4,3> (64 B rows)
__host__ __device__ __forceinline__ uint32_t swz(uint32_t r) { return ROW_BYTES == 128 ? (r & 7u) : ((r >> 1) & 3u); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---------------------------------------------------------------------------------------------------------
// pack: bit-permuted complex operand -> operand image
// ---------------------------------------------------------------------------------------------------------
struct PackParams {
  const float2* src;
  uint8_t* img;
  int64_t src_set_stride;   // complex entries between parameter sets (0: shared)
  int64_t img_z_stride;     // bytes between consecutive z (= set * 2^n_b + bb) images
  int32_t n_row, n_k, n_b;  // log2 extents
  int32_t is_b;             // 0: A image (ROWS rows per tile), 1: B image (c_t rows -> 2 c_t image rows)
  int32_t rows_t_log2;      // log2 rows per tile (7 for A, log2(c_t) for B)
  int32_t kblocks;          // max(1, K / KB_CPLX)
  int32_t n_local;          // local bits handled inside one block = rows_t_log2 + min(n_k, KB_LOG)
  int8_t row_bits[32], k_bits[32], b_bits[32];  // source bit of row / k / kept-shared index bit j
  int8_t local_src[16];     // source bit of local element bit j (sorted: ascending source position)
  int16_t local_dst[16];    // its value in (r << 4 | kk)
};

// Both operands of a step are packed by ONE launch (blocks [0, blocks_a) take operand A, the rest operand B; either
// count may be zero when that image is pinned).  A block packs PACK_NCH consecutive k-blocks of one row tile, so
// that the per-block address set-up is paid once per 8 chunks; chunks alternate between two shared-memory buffers
// and leave through the bulk-copy engine while the next one is being gathered.
constexpr int PACK_NCH = 8;
struct PackPair {
  PackParams a, b;
  int32_t blocks_a, blocks_b;
  int32_t nch_a, nch_b;  // k-blocks per block (<= PACK_NCH): fewer when the operand is small, so that the grid
                         // still covers the GPU several times over
};
inline int pack_nch(int64_t chunks, int num_sms) {
  int nch = PACK_NCH;
  while (nch > 1 && chunks / nch < 6 * (int64_t)num_sms) nch >>= 1;
  return nch;
}
inline int pack_blocks(int tiles, int kblocks, int nch) { return tiles * ((kblocks + nch - 1) / nch); }

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_tc_pack(const __grid_constant__ PackPair pp) {
  static_assert(THREADS == 256, "element = tid + 256 * i");
  extern __shared__ __align__(128) uint8_t pk_smem[];
  const bool second = (int)blockIdx.x >= pp.blocks_a;
  const PackParams& p = second ? pp.b : pp.a;
  const int tid = threadIdx.x;
  const int rows_t = 1 << p.rows_t_log2;
  const int plane = p.is_b ? b_plane_bytes(rows_t) : A_PLANE;
  const int chunk = 2 * plane;
  const uint32_t blk = blockIdx.x - (second ? (uint32_t)pp.blocks_a : 0u);
  const uint32_t per = (uint32_t)(second ? pp.nch_b : pp.nch_a);
  const uint32_t groups = ((uint32_t)p.kblocks + per - 1) / per;
  const uint32_t kb0 = (blk % groups) * per;
  const uint32_t tile = blk / groups;
  const int nch = min((int)per, p.kblocks - (int)kb0);
  const uint32_t z = blockIdx.y;
  const uint32_t bb = z & ((1u << p.n_b) - 1u);
  const uint32_t set = z >> p.n_b;
  // base source offset of this (set, bb, tile)
  // bit offsets inside the tensor are OR-ed together; the set offset is ADDED (a per-set arena stride is not a
  // multiple of the tensor's size in general)
  int64_t base = 0;
  for (int j = 0; j < p.n_b; ++j) base |= (int64_t)((bb >> j) & 1u) << p.b_bits[j];
  for (int j = p.rows_t_log2; j < p.n_row; ++j) base |= (int64_t)((tile >> (j - p.rows_t_log2)) & 1u) << p.row_bits[j];
  base += (int64_t)set * p.src_set_stride;
  if (p.n_k < KB_LOG) {  // K padded to one k-block: the missing columns are zero (nch == 1)
    for (int i = tid; i < chunk / 16; i += THREADS) reinterpret_cast<float4*>(pk_smem)[i] = make_float4(0, 0, 0, 0);
    __syncthreads();
  }
  // element e = tid + 256 i: bits 0..7 come from tid (resolved once), bits 8.. from i.  Local bits are sorted by
  // source position, so consecutive threads read ascending source addresses.
  uint32_t so_t = 0, d_t = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < p.n_local && ((tid >> j) & 1)) {
      so_t |= 1u << p.local_src[j];
      d_t |= (uint32_t)p.local_dst[j];
    }
  const int iters = p.n_local > 8 ? 1 << (p.n_local - 8) : 1;
  const bool active = p.n_local >= 8 || tid < (1 << p.n_local);
  uint32_t so_i[8], d_i[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t so = so_t, d = d_t;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (8 + j < p.n_local && ((i >> j) & 1)) {
        so |= 1u << p.local_src[8 + j];
        d |= (uint32_t)p.local_dst[8 + j];
      }
    so_i[i] = so;
    const uint32_t r = d >> 4, kk = d & 15u;
    d_i[i] = r * ROW_BYTES + ((((kk >> 1) ^ swz(r)) << 4) | ((kk & 1u) << 3));  // byte offset inside a plane
  }
  uint8_t* img = p.img + (int64_t)z * p.img_z_stride + ((int64_t)tile * p.kblocks + kb0) * (int64_t)chunk;
  for (int c = 0; c < nch; ++c) {
    const uint32_t kb = kb0 + (uint32_t)c;
    int64_t kbase = 0;
    for (int j = KB_LOG; j < p.n_k; ++j) kbase |= (int64_t)((kb >> (j - KB_LOG)) & 1u) << p.k_bits[j];
    const float2* src = p.src + base + kbase;
    uint8_t* buf = pk_smem + (c & 1) * chunk;
    if (c >= 2) {  // the store issued two chunks ago must have finished reading this buffer
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
    }
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < iters && active) v[i] = __ldg(src + so_i[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < iters && active) {
        const float hr = to_tf32(v[i].x), hi = to_tf32(v[i].y);
        const float lr = to_tf32(v[i].x - hr), li = to_tf32(v[i].y - hi);
        if (!p.is_b) {
          *reinterpret_cast<float2*>(buf + d_i[i]) = make_float2(hr, hi);
          *reinterpret_cast<float2*>(buf + plane + d_i[i]) = make_float2(lr, li);
        } else {
          const uint32_t o_im = d_i[i] + (uint32_t)rows_t * ROW_BYTES;  // -> Im C  (swz(rows_t + r) == swz(r))
          *reinterpret_cast<float2*>(buf + d_i[i]) = make_float2(hr, -hi);
          *reinterpret_cast<float2*>(buf + o_im) = make_float2(hi, hr);
          *reinterpret_cast<float2*>(buf + plane + d_i[i]) = make_float2(lr, -li);
          *reinterpret_cast<float2*>(buf + plane + o_im) = make_float2(li, lr);
        }
      }
    }
    // the finished chunk leaves through the bulk-copy engine: one contiguous store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(img + (int64_t)c * chunk),
                   "r"(smem_u32(buf)), "r"((uint32_t)chunk)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------
// GEMM over operand images
// ---------------------------------------------------------------------------------------------------------
// Image output ("fused pack"): the epilogue writes the result straight into the operand image of the ONE step that
// consumes it (split into hi / lo TF32 planes, permuted, swizzled), so that the consumer needs no pack pass: the
// tensor is never stored in its plain layout and never re-read.  Every bit of the consumer's image address is a
// GF(2)-linear function of the producer's (row, column, kept-shared) index bits — the swizzle is an XOR of address
// bits 4-5 with row bits 1-2 — so the byte offset of element (row, col, bb) is the XOR of one table entry per set
// index bit.
struct ImgOut {
  uint8_t* img;          // nullptr: plain output
  int64_t set_stride;    // bytes between parameter sets
  uint32_t rmap[26];     // byte-offset contribution of accumulator row bit i
  uint32_t cmap[26];     // ... of accumulator column bit i
  uint32_t bmap[8];      // ... of kept-shared index bit i
  uint32_t plane;        // hi plane -> lo plane
  uint32_t im_off;       // B image: Re rows -> Im rows (0: the consumer reads this tensor as its row operand A)
  int32_t n_row, n_col, n_b;  // log2 rows / columns / kept-shared extent of THIS step
};

// hi / lo split of one complex result and its stores into an A image (one place) or a B image (Re row and Im row)
__device__ __forceinline__ void img_store1(uint8_t* dst, uint32_t plane, uint32_t im_off, float re, float im) {
  const float hr = to_tf32(re), hi = to_tf32(im);
  const float lr = to_tf32(re - hr), li = to_tf32(im - hi);
  if (im_off == 0) {
    *reinterpret_cast<float2*>(dst) = make_float2(hr, hi);
    *reinterpret_cast<float2*>(dst + plane) = make_float2(lr, li);
  } else {
    *reinterpret_cast<float2*>(dst) = make_float2(hr, -hi);
    *reinterpret_cast<float2*>(dst + im_off) = make_float2(hi, hr);
    *reinterpret_cast<float2*>(dst + plane) = make_float2(lr, -li);
    *reinterpret_cast<float2*>(dst + plane + im_off) = make_float2(li, lr);
  }
}
// two results that are neighbours along the image's k (16 contiguous bytes per plane)
__device__ __forceinline__ void img_store2(uint8_t* dst, uint32_t plane, uint32_t im_off, float re0, float im0, float re1,
                                           float im1) {
  const float hr0 = to_tf32(re0), hi0 = to_tf32(im0), hr1 = to_tf32(re1), hi1 = to_tf32(im1);
  const float lr0 = to_tf32(re0 - hr0), li0 = to_tf32(im0 - hi0), lr1 = to_tf32(re1 - hr1), li1 = to_tf32(im1 - hi1);
  if (im_off == 0) {
    *reinterpret_cast<float4*>(dst) = make_float4(hr0, hi0, hr1, hi1);
    *reinterpret_cast<float4*>(dst + plane) = make_float4(lr0, li0, lr1, li1);
  } else {
    *reinterpret_cast<float4*>(dst) = make_float4(hr0, -hi0, hr1, -hi1);
    *reinterpret_cast<float4*>(dst + im_off) = make_float4(hi0, hr0, hi1, hr1);
    *reinterpret_cast<float4*>(dst + plane) = make_float4(lr0, -li0, lr1, -li1);
    *reinterpret_cast<float4*>(dst + plane + im_off) = make_float4(li0, lr0, li1, lr1);
  }
}

struct GemmParams {
  const uint8_t* img_a;
  const uint8_t* img_b;
  float2* c;
  int64_t img_a_z, img_b_z;  // bytes between z images
  int64_t img_a_set, img_b_set;  // bytes between parameter sets (img_z << n_b for scratch / pinned images; the per-set
                                 // arena stride when the producing step wrote the image)
  ImgOut out;
  int64_t c_set_stride;      // complex entries between parameter sets of C
  int64_t c_rs, c_cs;        // complex-entry stride of an accumulator row / column inside C
  int32_t tiles_a, tiles_b, kblocks, n_z, n_b_log2, stages;
  // order of the tiles of one z over the persistent CTAs: consecutive work items run on different SMs at the same
  // time, so the tile index that varies fastest shares ITS PARTNER's tile through L2.  tb_fast = 1: column tile
  // fastest (the row operand's tile is fetched from HBM once, by two SMs at once) — chosen when the A image is the
  // larger one; 0: row tile fastest.
  int32_t tb_fast;
  int32_t chunk;             // k-blocks accumulated inside the tensor core before a drain (see below)
  // experiments only (TQ_TC_DEBUG; bits 0-3 give wrong results): bit 0 = drains skip their TMEM loads; gather-A
  // variant: bit 1 = no generic->async proxy fence, bit 2 = no conversion, bit 3 = no loads, bit 4 = every lane polls
  int32_t debug;
  // split-K: a step with few output tiles and a long K is cut into `splits` K ranges per tile so that every SM
  // has work; split s writes its partial sums to c + s * c_split_stride, k_tc_splitk_sum adds them in order
  int32_t splits, kb_per_split;
  int64_t c_split_stride;
  int64_t c_bb_stride;       // complex entries between kept-shared index values (2^(n_m + n_n))
  // gather-A variant (k_tc_gemm<C_T, true>): the A operand is read ONCE from its tensor (bit-permuted gather),
  // split and written to the swizzled stage by two extra warps; there is no A image.  ga = pack tables of A.
  PackParams ga;
};

// Accumulation is two-level.  tcgen05 adds into its fp32 accumulator with round-toward-zero, a bias of about
// 2e-8 per MMA that grows linearly with K (measured: 6e-5 at K = 4096).  So only `chunk` k-blocks (12 MMAs each)
// are accumulated in TMEM; the 8 accumulation warps then drain that partial sum and add it to fp32 registers with
// round-to-nearest while the MMA issuer fills the other TMEM buffer (Ootomo & Yokota's error-compensated scheme,
// moved from mma.sync registers to TMEM).
//
// GA = true: the "streamed" operand (the one whose tiles are read once: tiles_b <= 2) never takes the detour
// through an HBM image (write 2x its bytes, read them back).  Warps 10..11 gather a k-block of A straight from the
// tensor (bit-permuted 8-byte loads, ascending addresses across the warp), split it into hi / lo and store both
// planes at their swizzled places in the stage; a generic->async proxy fence and one arrive per warp on the
// stage's full barrier (which also counts the B image's bulk-copy bytes) hand it to the MMA issuer.  The loads of
// the next k-block are in flight while the current one waits for its stage.
template <int C_T, bool GA>
__global__ void __launch_bounds__(GA ? GEMM_THREADS_GA : GEMM_THREADS, 1)
k_tc_gemm(const __grid_constant__ GemmParams p) {
  constexpr int HALF = C_T / 2;  // complex columns owned by one accumulation warp
  extern __shared__ __align__(1024) uint8_t g_smem[];
  __shared__ __align__(8) uint64_t bars[4 + 4 + 2 + 2 + 4];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(g_smem) + 1023u) & ~1023u;
  uint8_t* epi_smem = g_smem + (smem0 - smem_u32(g_smem)) + SMEM_BUDGET;  // after the operand stages
  const int S = p.stages;
  constexpr uint32_t sbytes = (uint32_t)(A_CHUNK + 4 * C_T * ROW_BYTES);
  constexpr uint32_t b_plane = (uint32_t)(2 * C_T * ROW_BYTES);
  constexpr uint32_t b_chunk = 2 * b_plane;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (10 + s); };
  auto raw_full_bar = [&](int s) { return bar0 + 8u * (12 + s); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), GA ? 1 + CONV_WARPS : 1);  // GA: the B bulk copy's expect_tx arrive + the converter warps
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(raw_full_bar(s), 64);  // one cp.async-tracking arrive per loader lane
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), ACC_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int64_t tiles_per_z = (int64_t)p.tiles_a * p.tiles_b;
  const int64_t total = tiles_per_z * p.n_z * p.splits;  // work item = (tile, K range), K range fastest
  constexpr int ACC0 = GA ? GA_ACC0 : 2;
  auto role_producer = [&]() {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t t = w / p.splits;
        const int kbeg = (int)(w - t * p.splits) * p.kb_per_split;
        const int64_t z = t / tiles_per_z;
        const int64_t r = t - z * tiles_per_z;
        const int64_t tb = p.tb_fast ? r % p.tiles_b : r / p.tiles_a;
        const int64_t ta = p.tb_fast ? r / p.tiles_b : r - tb * p.tiles_a;
        const int64_t zset = z >> p.n_b_log2, zbb = z & (((int64_t)1 << p.n_b_log2) - 1);
        const uint8_t* ga = p.img_a + zset * p.img_a_set + zbb * p.img_a_z + ta * (int64_t)p.kblocks * A_CHUNK;
        const uint8_t* gb = p.img_b + zset * p.img_b_set + zbb * p.img_b_z + tb * (int64_t)p.kblocks * b_chunk;
        for (int kb = kbeg; kb < kbeg + p.kb_per_split; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), GA ? b_chunk : sbytes);
          const uint32_t sa = smem0 + (uint32_t)stage * sbytes;
          if (!GA) bulk_g2s(sa, ga + (int64_t)kb * A_CHUNK, A_CHUNK, full_bar(stage));
          bulk_g2s(sa + A_CHUNK, gb + (int64_t)kb * b_chunk, b_chunk, full_bar(stage));
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  };
  auto role_issuer = [&]() {
    if (lane == 0) {
      constexpr uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * C_T >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        for (int kb0 = 0; kb0 < p.kb_per_split; kb0 += p.chunk) {
          const int kb1 = min(kb0 + p.chunk, p.kb_per_split);
          mbar_wait(tempty_bar(as), aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem0 + (uint32_t)stage * sbytes;
            const uint64_t a_hi = smem_desc(sa), a_lo = smem_desc(sa + A_PLANE);
            const uint64_t b_hi = smem_desc(sa + A_CHUNK), b_lo = smem_desc(sa + A_CHUNK + b_plane);
#pragma unroll
            for (int ks = 0; ks < K_STEPS; ++ks) {  // 8 reals (32 B) of K per instruction
              const uint64_t adv = (uint64_t)(ks * 2);
              tc_mma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (uint32_t)((kb > kb0) | (ks > 0)));
              tc_mma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
              tc_mma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            tc_commit(empty_bar(stage));
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
          tc_commit(tfull_bar(as));
          if (++as == 2) {
            as = 0;
            aphase ^= 1u;
          }
        }
      }
    }
  };
  auto role_gather = [&]() {
    // Gather-A roles.  Warps 2..3 (loaders) request a k-block of A as 8-byte cp.async copies that land at their
    // swizzled places in the hi plane of a free stage (no registers held; up to `stages` k-blocks in flight per SM);
    // a cp.async-tracking arrive on the stage's raw barrier fires when they have landed.  Warps 12..15 (converters)
    // then split the k-block in place (hi rounded back to where it was, lo to the other plane), fence generic ->
    // async proxy and arrive on the stage's full barrier.  Loading and converting are separate warps because
    // that proxy fence waits for the executing thread's outstanding copies: a thread that both prefetches and
    // fences has nothing in flight across the fence.
    const PackParams& P = p.ga;
    const int64_t my_items = (int64_t)blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_blocks = my_items * p.kb_per_split;
    const bool loader = warp < 4;
    // loaders: element e = tl + 64 i (tl = 0..63); converters: e = tc + 128 i (tc = 0..127)
    const int tl = loader ? (int)threadIdx.x - 64 : (int)threadIdx.x - 32 * GA_CONV0;
    const int tbits = loader ? 6 : 7;
    uint32_t so_t = 0, d_t = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < tbits && ((tl >> j) & 1)) {
        so_t |= 1u << P.local_src[j];
        d_t |= (uint32_t)P.local_dst[j];
      }
    auto dst_of = [&](uint32_t d) {
      const uint32_t r = d >> 4, kk = d & 15u;
      return r * ROW_BYTES + ((((kk >> 1) ^ swz(r)) << 4) | ((kk & 1u) << 3));
    };
    int stage = 0;
    uint32_t phase = 0;
    // one polling lane per warp: 192 threads spinning on try_wait would compete with the MMA issuer's own waits
    auto warp_wait = [&](uint32_t bar, uint32_t parity) {
      if (p.debug & 16) {
        mbar_wait(bar, parity);
      } else {
        if (lane == 0) mbar_wait(bar, parity);
        __syncwarp();
      }
    };
    if (loader) {
      uint32_t so_i[LOAD_ITERS], d_i[LOAD_ITERS];
#pragma unroll
      for (int i = 0; i < LOAD_ITERS; ++i) {
        uint32_t so = so_t, d = d_t;
#pragma unroll
        for (int j = 6; j < 7 + KB_LOG; ++j)
          if ((i >> (j - 6)) & 1) {
            so |= 1u << P.local_src[j];
            d |= (uint32_t)P.local_dst[j];
          }
        so_i[i] = so;
        d_i[i] = dst_of(d);
      }
      const uint32_t bb_mask = (1u << P.n_b) - 1u;
      int64_t wi = blockIdx.x;
      int kbi = (int)(wi % p.splits) * p.kb_per_split, kend = kbi + p.kb_per_split;
      for (int64_t n = 0; n < n_blocks; ++n) {
        const int64_t t = wi / p.splits;
        const int64_t z = t / tiles_per_z;
        const int64_t r = t - z * tiles_per_z;
        const uint32_t ta = (uint32_t)(p.tb_fast ? r / p.tiles_b : r % p.tiles_a);
        const uint32_t bb = (uint32_t)z & bb_mask;
        int64_t base = 0;
        for (int j = 0; j < P.n_b; ++j) base |= (int64_t)((bb >> j) & 1u) << P.b_bits[j];
        for (int j = 7; j < P.n_row; ++j) base |= (int64_t)((ta >> (j - 7)) & 1u) << P.row_bits[j];
        for (int j = KB_LOG; j < P.n_k; ++j) base |= (int64_t)(((uint32_t)kbi >> (j - KB_LOG)) & 1u) << P.k_bits[j];
        const float2* src = P.src + base + (z >> P.n_b) * P.src_set_stride;
        warp_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t dst = smem0 + (uint32_t)stage * sbytes;
        if (!(p.debug & 8)) {
#pragma unroll
          for (int i = 0; i < LOAD_ITERS; ++i)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + d_i[i]), "l"(src + so_i[i]) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(raw_full_bar(stage)) : "memory");
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
        if (++kbi == kend) {
          wi += gridDim.x;
          kbi = (int)(wi % p.splits) * p.kb_per_split;
          kend = kbi + p.kb_per_split;
        }
      }
    } else {
      uint32_t d_i[CONV_ITERS];
#pragma unroll
      for (int i = 0; i < CONV_ITERS; ++i) {
        uint32_t d = d_t;
#pragma unroll
        for (int j = 7; j < 7 + KB_LOG; ++j)
          if ((i >> (j - 7)) & 1) d |= (uint32_t)P.local_dst[j];
        d_i[i] = dst_of(d);
      }
      for (int64_t n = 0; n < n_blocks; ++n) {
        warp_wait(raw_full_bar(stage), phase);
        uint8_t* buf = g_smem + (smem0 - smem_u32(g_smem)) + (size_t)stage * sbytes;
        if (!(p.debug & 4)) {
          float2 v[CONV_ITERS];
#pragma unroll
          for (int i = 0; i < CONV_ITERS; ++i) v[i] = *reinterpret_cast<const float2*>(buf + d_i[i]);
#pragma unroll
          for (int i = 0; i < CONV_ITERS; ++i) {
            const float hr = to_tf32(v[i].x), hi = to_tf32(v[i].y);
            const float lr = to_tf32(v[i].x - hr), li = to_tf32(v[i].y - hi);
            *reinterpret_cast<float2*>(buf + d_i[i]) = make_float2(hr, hi);
            *reinterpret_cast<float2*>(buf + A_PLANE + d_i[i]) = make_float2(lr, li);
          }
        }
        if (!(p.debug & 2)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(stage));
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  };
  auto role_acc = [&]() {
    const int q = warp & 3;           // TMEM lane quadrant this warp may read
    const int h = (warp - ACC0) >> 2;  // which half of the complex columns it accumulates
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t bb_mask = (1u << p.n_b_log2) - 1u;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
      const int64_t t = w / p.splits;
      const int64_t split = w - t * p.splits;
      float acc_re[HALF], acc_im[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) acc_re[j] = acc_im[j] = 0.f;
      for (int kb0 = 0; kb0 < p.kb_per_split; kb0 += p.chunk) {
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)as * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * HALF);
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
          if (p.debug & 1) break;
          uint32_t re[16], im[16];
          tc_ld16(taddr + (uint32_t)c0, re);
          tc_ld16(taddr + (uint32_t)(C_T + c0), im);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c0 + j < HALF) {
              acc_re[c0 + j] += __uint_as_float(re[j]);
              acc_im[c0 + j] += __uint_as_float(im[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
        if (++as == 2) {
          as = 0;
          aphase ^= 1u;
        }
      }
      const int64_t z = t / tiles_per_z;
      const int64_t r = t - z * tiles_per_z;
      const int64_t tb = p.tb_fast ? r % p.tiles_b : r / p.tiles_a;
      const int64_t ta = p.tb_fast ? r / p.tiles_b : r - tb * p.tiles_a;
      const int64_t set = z >> p.n_b_log2, bb = z & bb_mask;
      const int64_t row = ta * ROWS + q * 32 + lane;
      const int64_t col0 = tb * C_T + h * HALF;
      float2* dst = p.c + split * p.c_split_stride + set * p.c_set_stride + bb * p.c_bb_stride + row * p.c_rs +
                    col0 * p.c_cs;
      if (p.out.img != nullptr) {
        // image output: hi / lo split and scatter into the consumer's operand image (see ImgOut)
        constexpr int LOG2HALF = HALF == 64 ? 6 : HALF == 32 ? 5 : HALF == 16 ? 4 : 3;
        const uint32_t urow = (uint32_t)row, ucol0 = (uint32_t)col0, ubb = (uint32_t)bb;
        uint32_t off = 0;
        for (int i = 0; i < p.out.n_row; ++i)
          if ((urow >> i) & 1u) off ^= p.out.rmap[i];
        for (int i = LOG2HALF; i < p.out.n_col; ++i)
          if ((ucol0 >> i) & 1u) off ^= p.out.cmap[i];
        for (int i = 0; i < p.n_b_log2; ++i)
          if ((ubb >> i) & 1u) off ^= p.out.bmap[i];
        uint32_t cm[LOG2HALF];
#pragma unroll
        for (int i = 0; i < LOG2HALF; ++i) cm[i] = p.out.cmap[i];
        uint8_t* base = p.out.img + set * p.out.set_stride;
        const uint32_t plane = p.out.plane, im_off = p.out.im_off;
        if (cm[0] == 8u) {  // column bit 0 is the image's lowest k bit: neighbouring columns share a 16-byte word
#pragma unroll
          for (int j = 0; j < HALF; j += 2) {
            uint32_t o = off;
#pragma unroll
            for (int i = 1; i < LOG2HALF; ++i)
              if ((j >> i) & 1) o ^= cm[i];
            img_store2(base + o, plane, im_off, acc_re[j], acc_im[j], acc_re[j + 1], acc_im[j + 1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < HALF; ++j) {
            uint32_t o = off;
#pragma unroll
            for (int i = 0; i < LOG2HALF; ++i)
              if ((j >> i) & 1) o ^= cm[i];
            img_store1(base + o, plane, im_off, acc_re[j], acc_im[j]);
          }
        }
      } else if (p.c_cs == 1) {
        // Row-major result: a lane owns a row, so direct stores would touch 32 rows with 16 bytes each per
        // instruction.  Stage 64-byte row pieces in shared memory and let 4 lanes write one row's piece: every
        // store instruction covers 8 rows x 64 contiguous bytes (full sectors).
        uint8_t* stg = epi_smem + (warp - ACC0) * EPI_WARP_BYTES;
        float2* tile0 = p.c + split * p.c_split_stride + set * p.c_set_stride + bb * p.c_bb_stride +
                        (ta * ROWS + q * 32) * p.c_rs + col0;
#pragma unroll
        for (int j0 = 0; j0 < HALF; j0 += 8) {
          float4* mine = reinterpret_cast<float4*>(stg + lane * EPI_PITCH);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            mine[c4] = make_float4(acc_re[j0 + 2 * c4], acc_im[j0 + 2 * c4], acc_re[j0 + 2 * c4 + 1],
                                   acc_im[j0 + 2 * c4 + 1]);
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int row_l = rr * 8 + (lane >> 2), c4 = lane & 3;
            const float4 v = *reinterpret_cast<const float4*>(stg + row_l * EPI_PITCH + c4 * 16);
            *reinterpret_cast<float4*>(tile0 + (int64_t)row_l * p.c_rs + j0 + 2 * c4) = v;
          }
          __syncwarp();
        }
      } else {
#pragma unroll
        for (int j = 0; j < HALF; ++j) dst[j * p.c_cs] = make_float2(acc_re[j], acc_im[j]);
      }
    }
  
  };
  if constexpr (GA) {
    // 512 threads start with 128 registers each; setmaxnreg moves registers from the light warpgroups to the two
    // accumulation warpgroups (88 * 128 + 72 * 128 + 176 * 256 = 65536).  One setmaxnreg per warpgroup, at the top
    // of that warpgroup's branch, so that the register allocator sees the limit of the code that follows it.
    const int wg = warp >> 2;
    if (wg == 0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
      if (warp == 0) role_producer();
      else if (warp == 1) role_issuer();
      else role_gather();
    } else if (wg == 3) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
      role_gather();
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
      role_acc();
    }
  } else {
    if (warp == 0) role_producer();
    else if (warp == 1) role_issuer();
    else role_acc();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// C[set][i] = sum over splits of part[split][set][i], fixed order (deterministic); float4 = 2 complex
__global__ void k_tc_splitk_sum(const float4* __restrict__ part, int splits, int64_t split_stride4,
                                int64_t part_set_stride4, float4* __restrict__ c, int64_t c_set_stride4, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  part += blockIdx.y * part_set_stride4;
  float4 acc = part[i];
  for (int s = 1; s < splits; ++s) {
    const float4 v = part[s * split_stride4 + i];
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  c[blockIdx.y * c_set_stride4 + i] = acc;
}

// Split-K step whose result feeds a fused-pack consumer: adds the partial sums (stored [split][set][bb][row][col] in
// accumulator order) in a fixed order and writes the consumer's operand image (see ImgOut).
__global__ void k_tc_splitk_sum_img(const float2* __restrict__ part, int splits, int64_t split_stride,
                                    int64_t part_set_stride, const __grid_constant__ ImgOut out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t set = blockIdx.y;
  part += set * part_set_stride;
  float2 acc = part[i];
  for (int s = 1; s < splits; ++s) {
    const float2 v = part[s * split_stride + i];
    acc.x += v.x;
    acc.y += v.y;
  }
  const uint32_t col = (uint32_t)i & ((1u << out.n_col) - 1u);
  const uint32_t row = (uint32_t)(i >> out.n_col) & ((1u << out.n_row) - 1u);
  const uint32_t bb = (uint32_t)(i >> (out.n_col + out.n_row));
  uint32_t off = 0;
  for (int b = 0; b < out.n_row; ++b)
    if ((row >> b) & 1u) off ^= out.rmap[b];
  for (int b = 0; b < out.n_col; ++b)
    if ((col >> b) & 1u) off ^= out.cmap[b];
  for (int b = 0; b < out.n_b; ++b)
    if ((bb >> b) & 1u) off ^= out.bmap[b];
  img_store1(out.img + set * out.set_stride + off, out.plane, out.im_off, acc.x, acc.y);
}

}  // namespace tc
}  // namespace tq
