// Gram matrix G += A^T A (A = DH, n x 704 fp32; solver.py:167 `DH.t() @ DH`) on the INT8 tensor cores with
// EXACT products and exact accumulation (an Ozaki-style error-free split), replacing the fp64-pipe DMMA path
// for large n.  The normal matrix has cond ~1e15: the Gram must be positive semi-definite to
// fp64 rounding or the Cholesky breaks down, and it must be the Gram of exactly the matrix that the RHS
// projection uses (measured on the CPU oracle, G = 32: a Gram that is off by 1e-11 costs 1e-4 in u).  The fp64
// pipe gives 35 TFLOP/s, the int8 tensor pipe 4.5 POP/s.
//
//   1. xop.cu               per-column max |a|  ->  exponent e_f with |a[:, f]| < 2^e_f (or the calibrated grid of
//                           the fused feature epilogue, layer_tc.cu)
//   2. gi_split_kernel      X = rint(a * 2^(30 - e_f)), a 31-bit integer.  For |a| >= 2^(e_f - 7) this is a itself;
//                           smaller elements move by < 2^-31 of the column maximum (< 1 % of one fp32 ulp of that
//                           maximum) and are SNAPPED IN PLACE: A is overwritten by X 2^(e_f - 30), which is again
//                           an fp32 number, so everything downstream (RHS projection, f_pred, the saved state)
//                           sees exactly the matrix whose Gram is computed.  X is written as four balanced
//                           radix-256 digits  X = d0 2^24 + d1 2^16 + d2 2^8 + d3,  d in [-128, 127], into int8
//                           planes laid out K-major and blocked by 128-point slab for the tensor core:
//                           planes[slab][digit][feature][128 points] - every TMA box is one contiguous 16 KiB
//                           (a feature-major layout with 1 MiB between rows ran at half the speed)
//   3. gram_i8_kernel       per (256 x 64) tile and K range:  Acc_o = sum_{a+b=o} D_a^T D_b  for ALL 16 digit pairs,
//                           orders o = 0..6, with tcgen05.mma.kind::i8 into seven int32 TMEM accumulators
//                           (exact; drained before 2^31 can be reached), cta_group::2 pairs, TMA-fed
//   4. gi_reduce_kernel     G[r][c] += 2^(e_r + e_c - 12) * sum_splits sum_o 2^(-8 o) Acc_o   (fixed order, fp64)
//
// The same engine computes the batched RHS projection R += DH^T F (solver.py:197 + :301, launch_project_i8):
// A = the digit planes of DH (the SAME planes the Gram used when the caller keeps them, launch_project_x; rebuilt per
// call otherwise), B = digit planes of F (a copy on the grid of each right-hand side; F itself is not modified),
// rectangular tiling 3 x ceil(B / 64).
//
// Every product and every partial sum is an exact integer; the only rounding is the final fp64 combination
// (2^-53 relative), so G is the Gram of the stored A to fp64 rounding, bit-reproducible and independent of the
// K split.  (Dropping the orders 5 and 6, 2^-37 of the product of the column maxima, was tried first: the
// Cholesky then needs a diagonal shift and f_pred moves by 1.5e-4.)
#include "fps_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <algorithm>

namespace fps {

constexpr int GI_SL = 4;                    // int8 digits per value
constexpr int GI_TM = 128;                  // feature rows per CTA (UMMA M = 256 per pair)
constexpr int GI_TN = 64;                   // feature columns per tile (UMMA N); each CTA stages 32 of them
constexpr int GI_NACC = 7;                  // digit orders 0..6, one int32 accumulator each (448 TMEM columns)
constexpr int GI_SLAB = 128;                // points per digit-plane row chunk (split kernel granularity)
constexpr int GI_ROWS = FPW;                // 768 feature rows per digit plane
// K slab of the tensor-core kernel = RB points = one RB-byte swizzled row (template parameter):
//   RB = 128: 80 KiB stages x 2 (production);  RB = 64: 40 KiB stages x 5 (FPS_GRAM_ROWB=64, measured slower)
constexpr int GI_RING = 220 * 1024;
constexpr int GI_SMEM = GI_RING + 1024 + 256;
constexpr int GI_THREADS = 320;             // TMA warp, MMA warp, 8 drain warps
constexpr int GI_NTILES = 23;               // (256 x 64) tiles that touch the lower triangle of 704 x 704 (padded to 768 rows)
// points per TMEM drain: an order holds at most 4 digit pairs of magnitude <= 2^14 -> 2^16 per point;
// 28 672 points keep every accumulator below 2^31
constexpr int GI_SEG_POINTS = 28672;
constexpr int GI_TILE_ELEMS = 2 * GI_TM * GI_TN;  // fp64 partial tile of a pair

// row block I (256 features) needs the column blocks J (64 features) with 64 J <= 256 I + 255 and 64 J < 704: 4 + 8 + 11
__host__ __device__ inline void gi_tile(int t, int& I, int& J) {
  if (t < 4) { I = 0; J = t; }
  else if (t < 12) { I = 1; J = t - 4; }
  else { I = 2; J = t - 12; }
}

// ---- 1. column maxima and exponents: xop.cu (shared with recon_i8.cu) --------------------------------

// ---- 2. digit split + transpose -------------------------------------------------------------------
// CTA = 128 points x 64 features.  Reads are coalesced along the feature axis, the digits go through shared
// memory (row pitch 33 words: conflict-free both ways) and leave as 128-byte rows along the point axis.
// SNAP: write X 2^(e - 30) back into A (the Gram / projection operand DH); otherwise A is read only (F).
// rows = feature rows per digit plane (a multiple of 64 >= ncols); columns >= ncols produce zero digits.
template <bool SNAP>
__global__ void __launch_bounds__(256)
gi_split_kernel(float* __restrict__ A, int64_t ld, int64_t n, int ncols, int rows, int slabs,
                const float2* __restrict__ scale, int8_t* __restrict__ planes, const int* __restrict__ cond) {
  __shared__ uint32_t sm[GI_SL][64][33];
  if (cond && *cond == 0) return;   // overflow fallback of the fused feature epilogue: nothing to redo (whole CTA)
  const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int fh = w & 1, pq = w >> 1;
  const int fl = fh * 32 + lane;
  // feature block = blockIdx.x (fastest varying): the CTAs that share a slab's rows run together, so each
  // 128-row band of A is read as whole rows by neighbouring CTAs (DRAM page locality); slab = blockIdx.y + 65535 z
  const int fb = blockIdx.x;
  const int64_t slab_idx = (int64_t)blockIdx.z * 65535 + blockIdx.y;
  const int f = fb * 64 + fl;
  const bool col_ok = f < ncols;
  if (slab_idx >= slabs) return;   // (whole CTA: before the barrier)
  const float sc = scale[f].x, inv_sc = scale[f].y;   // powers of two: both products below are exact
  const int64_t pbase = slab_idx * GI_SLAB + pq * 32;
  // all 32 loads first: the in-place stores below would otherwise serialise them (same array)
  float vals[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const int64_t p = pbase + q;
    vals[q] = (col_ok && p < n) ? A[p * ld + f] : 0.f;
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    uint32_t word[GI_SL], z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float v = vals[it * 4 + u];
      const int X = __float2int_rn(v * sc);
      if (SNAP) {
        const float snapped = (float)X * inv_sc;   // exact: X has at most 24 significant bits
        if (snapped != v) A[(pbase + it * 4 + u) * ld + f] = snapped;
      }
      z[u] = balanced_digits4(X);                  // |X| <= 2^30: [d0 | d1 | d2 | d3]
    }
    digits_transpose4(z[0], z[1], z[2], z[3], word);
#pragma unroll
    for (int s = 0; s < GI_SL; ++s) sm[s][fl][pq * 8 + it] = word[s];
  }
  __syncthreads();
  // planes[slab][digit][feature][128]: a (digit, 64-feature) group is 8 KiB of consecutive bytes
  int8_t* slab = planes + slab_idx * ((int64_t)GI_SL * rows * GI_SLAB);
#pragma unroll 4
  for (int rr = 0; rr < 32; ++rr) {
    const int row = w * 32 + rr;
    const int s = row / 64, rfl = row % 64;
    *reinterpret_cast<uint32_t*>(slab + ((int64_t)s * rows + fb * 64 + rfl) * GI_SLAB + lane * 4) = sm[s][rfl][lane];
  }
}

// ---- 3. the tensor-core kernel --------------------------------------------------------------------
// One instruction covers one A digit against ALL FOUR B digits: the B operand of a tile is the stack
//   B' = [D_0; D_1; D_2; D_3]  (4 x 64 feature rows = UMMA N 256; CTA 0 of the pair stages digits 0, 1, CTA 1 digits 2, 3)
// and the MMA with A = D_a writes the 256-column TMEM window that starts at column 64 a, so that column group o
// (64 columns) accumulates every digit pair with a + b = o:
//   window(a) = [64 a, 64 a + 256)  =  orders a, a + 1, a + 2, a + 3.
// Four N = 256 instructions per 32 points instead of sixteen N = 64 ones: the narrow form re-read the 4 KiB A tile
// from shared memory sixteen times and was bound by shared-memory bandwidth at 47 % tensor-pipe activity.
// All MMAs accumulate; the drain warps zero the accumulators after reading them (tcgen05.st).
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t da, uint64_t db) {
  // kind::i8: D s32 (c_format 2), A and B signed 8-bit (format 1), K-major, M = 256 (pair), N = 256, K = 32
  constexpr uint32_t idesc =
      (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((GI_SL * GI_TN) >> 3) << 17) | ((uint32_t)((2 * GI_TM) >> 4) << 24);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
      : "memory");
}

// 32 lanes x 16 consecutive columns of raw 32-bit words
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// zero 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const int z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}

struct GiParams {
  double* partial;       // nitems x (256 x 64) fp64
  int slabs_total;       // ceil(n / RB)
  int slabs_per_split;
  int ntiles;            // symmetric: GI_NTILES (lower triangle); else 3 x tiles_j
  int nitems;            // ntiles * nsplit; item = split * ntiles + tile
  int sym;               // 1: Gram (B planes = A planes), 0: rectangular A^T B
  int tiles_j;           // column blocks of 64 (rectangular case)
  int rows_a, rows_b;    // feature rows per digit plane of the A / B operand
  int slab_a0;           // first 128-point slab of this pass inside the A planes (B planes start at 0)
};

__device__ __forceinline__ void gi_item_tile(const GiParams& P, int item, int& I, int& J) {
  const int t = item % P.ntiles;
  if (P.sym) gi_tile(t, I, J);
  else { I = t / P.tiles_j; J = t % P.tiles_j; }
}

template <int RB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GI_THREADS, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GiParams P) {
  constexpr int GI_A_BYTES = GI_TM * RB;            // per digit: 128 feature rows of this CTA
  constexpr int GI_B_BYTES = GI_TN * RB;            // per digit: all 64 feature rows of the column block
  constexpr int GI_STAGE = GI_SL * GI_A_BYTES + (GI_SL / 2) * GI_B_BYTES;   // this CTA stages 2 of the 4 B digits
  constexpr int GI_STAGES = GI_RING / GI_STAGE;
  constexpr int GI_SEG_SLABS = GI_SEG_POINTS / RB;
  static_assert(GI_STAGES >= 2 && GI_STAGES <= 8, "ring depth");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + GI_RING;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 64, bar_tfull = bar_base + 128, bar_tempty = bar_base + 136;
  const uint32_t tmem_slot = bar_base + 144;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int rank = (int)cluster_ctarank();
  const int group = blockIdx.x / 2, ngroups = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < GI_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 16);  // one arrive per drain warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer (both CTAs; bytes are counted on the leader's barrier) =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = group; item < P.nitems; item += ngroups) {
        int I, J;
        gi_item_tile(P, item, I, J);
        const int kb0 = (item / P.ntiles) * P.slabs_per_split;
        const int kb1 = min(P.slabs_total, kb0 + P.slabs_per_split);
        const int arow = I * 2 * GI_TM + rank * GI_TM, brow = J * GI_TN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * GI_STAGE;
          const uint32_t full = bar_full + 8 * stage;
          if (rank == 0) mbar_expect_tx(full, 2 * GI_STAGE);
          const int kcol = (kb * RB) % GI_SLAB, slab = (kb * RB) / GI_SLAB;
#pragma unroll
          for (int s = 0; s < GI_SL; ++s)
            tma_load_2d<2>(sa + s * GI_A_BYTES, &tmA, full, kcol, ((P.slab_a0 + slab) * GI_SL + s) * P.rows_a + arow);
#pragma unroll
          for (int s = 0; s < GI_SL / 2; ++s)   // B' rows [128 rank, 128 rank + 128) = digits 2 rank, 2 rank + 1
            tma_load_2d<2>(sa + GI_SL * GI_A_BYTES + s * GI_B_BYTES, &tmB, full, kcol,
                           ((P.sym ? P.slab_a0 + slab : slab) * GI_SL + 2 * rank + s) * P.rows_b + brow);
          if (++stage == GI_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (pair leader) =================
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, tphase = 0;
      for (int item = group; item < P.nitems; item += ngroups) {
        const int kb0 = (item / P.ntiles) * P.slabs_per_split;
        const int kb1 = min(P.slabs_total, kb0 + P.slabs_per_split);
        for (int ks = kb0; ks < kb1; ks += GI_SEG_SLABS) {
          const int ke = min(kb1, ks + GI_SEG_SLABS);
          mbar_wait(bar_tempty, tphase);  // the drain warps have read and zeroed the accumulators
          tphase ^= 1;
          tc_fence_after();
          for (int kb = ks; kb < ke; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * GI_STAGE;
            const uint64_t db = umma_desc<RB>(sa + GI_SL * GI_A_BYTES);
#pragma unroll
            for (int j = 0; j < RB / 32; ++j) {  // 32 bytes (= 32 points) of K per instruction
              const uint64_t adv = (uint64_t)((j * 32) >> 4);
#pragma unroll
              for (int a = 0; a < GI_SL; ++a)
                umma_i8_pair(tmem_base + a * GI_TN, umma_desc<RB>(sa + a * GI_A_BYTES) + adv, db + adv);
            }
            umma_commit<2>(bar_empty + 8 * stage);
            if (++stage == GI_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit<2>(bar_tfull);
        }
      }
    }
  } else {
    // ================= drain warps 2..9: int32 accumulators -> fp64 partial tile =================
    const int quad = warp % 4;          // TMEM lanes [32 quad, 32 quad + 32)
    const int half = (warp - 2) / 4;    // 32 of the 64 columns of every order
    const uint32_t taddr = tmem_base + half * (GI_TN / 2) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
    for (int o = 0; o < GI_NACC; ++o)
#pragma unroll
      for (int c = 0; c < GI_TN / 32; ++c) tmem_zero16(taddr + o * GI_TN + c * 16);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_leader(bar_tempty);
    uint32_t tphase = 0;
    for (int item = group; item < P.nitems; item += ngroups) {
      const int kb0 = (item / P.ntiles) * P.slabs_per_split;
      const int kb1 = min(P.slabs_total, kb0 + P.slabs_per_split);
      double* out = P.partial + (size_t)item * GI_TILE_ELEMS + (size_t)(rank * GI_TM + quad * 32 + lane) * GI_TN + half * (GI_TN / 2);
      bool first_seg = true;
      for (int ks = kb0; ks < kb1; ks += GI_SEG_SLABS) {
        mbar_wait(bar_tfull, tphase);
        tphase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < GI_TN / 32; ++c) {
          int v[GI_NACC][16];
#pragma unroll
          for (int o = 0; o < GI_NACC; ++o) tmem_ld16(taddr + o * GI_TN + c * 16, v[o]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int o = 0; o < GI_NACC; ++o) tmem_zero16(taddr + o * GI_TN + c * 16);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            double g = (double)v[GI_NACC - 1][i];
#pragma unroll
            for (int o = GI_NACC - 2; o >= 0; --o) g = fma(g, 0.00390625, (double)v[o][i]);  // sum_o 2^(-8 o) Acc_o
            if (!first_seg) g += out[c * 16 + i];
            out[c * 16 + i] = g;
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar_tempty);
        first_seg = false;
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- 4. fixed-order reduction of the K splits, scaling, write-back ----------------------------------
// C[r][c] += 2^(ea[r] + eb[c] - 12) * sum_splits partial;  symmetric: lower triangle computed, mirrored.
__global__ void __launch_bounds__(256)
gi_reduce_kernel(const double* __restrict__ partial, int nsplit, int ntiles, int sym, int tiles_j, int M, int Nc,
                 const int* __restrict__ ea, const int* __restrict__ eb, double* __restrict__ C, int64_t ldc) {
  int I, J;
  if (sym) gi_tile(blockIdx.x, I, J);
  else { I = blockIdx.x / tiles_j; J = blockIdx.x % tiles_j; }
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < GI_TILE_ELEMS; e += gridDim.y * blockDim.x) {
    const int r = I * 2 * GI_TM + e / GI_TN, c = J * GI_TN + e % GI_TN;
    if (r >= M || c >= Nc || (sym && c > r)) continue;
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[((size_t)sp * ntiles + blockIdx.x) * GI_TILE_ELEMS + e];
    s = ldexp(s, ea[r] + eb[c] - 12);
    C[(int64_t)r * ldc + c] += s;
    if (sym && c < r) C[(int64_t)c * ldc + r] += s;
  }
}

// ---- host -----------------------------------------------------------------------------------------
struct GiOperand {           // digit planes of one operand, planes[slab][digit][rows][128], and its per-row grid
  int8_t* data;
  const int* exps;
  const float2* scale;
  int rows;                  // feature rows (A: 768) or right-hand sides rounded up to 64 (B) per digit plane
};

struct GiState {
  int8_t* a_data = nullptr;  // context-owned A planes (uncached entry points only)
  size_t a_cap = 0;
  int8_t* b_data = nullptr;  // digit planes of F (one projection pass)
  size_t b_cap = 0;
  unsigned* b_meta = nullptr;  // b_rows_cap x {u32 cmax, i32 exps, float2 scale}
  int b_rows_cap = 0;
  double* partial = nullptr;
  size_t partial_bytes = 0;
  void* encode = nullptr;
};

typedef CUresult (*GiEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int64_t GI_MAX_CHUNK = 1 << 20;          // points per pass through context-owned planes: 3 GiB
constexpr size_t GI_B_PLANE_BYTES = (size_t)5 << 30;  // budget for the digit planes of F per projection pass

void gram_i8_destroy(fps_ctx* ctx) {
  GiState* s = static_cast<GiState*>(ctx->gi_state);
  if (!s) return;
  cudaFree(s->a_data);
  cudaFree(s->b_data);
  cudaFree(s->b_meta);
  cudaFree(s->partial);
  delete s;
  ctx->gi_state = nullptr;
}

static int gi_state_get(fps_ctx* ctx, GiState** out) {
  GiState* s = static_cast<GiState*>(ctx->gi_state);
  if (!s) {
    s = new GiState();
    ctx->gi_state = s;
    cudaDriverEntryPointQueryResult qres;
    FPS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &s->encode, cudaEnableDefault, &qres));
    FPS_REQUIRE(s->encode && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    FPS_CUDA(cudaFuncSetAttribute(gram_i8_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GI_SMEM));
    FPS_CUDA(cudaFuncSetAttribute(gram_i8_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GI_SMEM));
  }
  *out = s;
  return 0;
}

static size_t gi_plane_bytes(int rows, int64_t points) {
  return (size_t)GI_SL * rows * ((points + GI_SLAB - 1) / GI_SLAB * GI_SLAB);
}

// grow-only context-owned buffers (workspace_bytes follows).  Growing frees and reallocates, which synchronises the
// device: fps_reserve_workspace sizes them up front so that nothing allocates mid-stream.
static int gi_grow(fps_ctx* ctx, int8_t** buf, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*buf) { FPS_CUDA(cudaFree(*buf)); ctx->workspace_bytes -= *cap; *buf = nullptr; *cap = 0; }
  FPS_CUDA(cudaMalloc(buf, need));
  *cap = need;
  ctx->workspace_bytes += need;
  return 0;
}

static int gi_reserve_b(fps_ctx* ctx, GiState* s, int rows_b, int64_t points) {
  if (int rc = gi_grow(ctx, &s->b_data, &s->b_cap, gi_plane_bytes(rows_b, points))) return rc;
  if (rows_b > s->b_rows_cap) {
    if (s->b_meta) FPS_CUDA(cudaFree(s->b_meta));
    FPS_CUDA(cudaMalloc(&s->b_meta, (size_t)rows_b * (sizeof(unsigned) + sizeof(int) + sizeof(float2))));
    s->b_rows_cap = rows_b;
  }
  return 0;
}

static int gi_reserve_partial(fps_ctx* ctx, GiState* s, int nitems) {
  const size_t need = (size_t)nitems * GI_TILE_ELEMS * sizeof(double);
  if (need > s->partial_bytes) {
    if (s->partial) { FPS_CUDA(cudaFree(s->partial)); ctx->workspace_bytes -= s->partial_bytes; }
    FPS_CUDA(cudaMalloc(&s->partial, need));
    s->partial_bytes = need;
    ctx->workspace_bytes += need;
  }
  return 0;
}

// rows of one projection pass for B right-hand sides: max_chunk / 2^k points, the F planes within their budget
static int64_t gi_project_chunk(int rows_b, int64_t n, int64_t max_chunk) {
  int64_t chunk = max_chunk;
  while ((size_t)GI_SL * rows_b * chunk > GI_B_PLANE_BYTES && chunk / 2 >= GI_SLAB) chunk /= 2;
  return chunk > n ? n : chunk;
}

static int64_t gi_env_chunk() {
  // FPS_GRAM_I8_CHUNK: points per pass (tests exercise the multi-pass loops with a small value)
  static const int64_t v = getenv("FPS_GRAM_I8_CHUNK") ? atoll(getenv("FPS_GRAM_I8_CHUNK")) : GI_MAX_CHUNK;
  return v;
}

// digit planes of `np` rows of A (starting at slab 0 of `planes`) on the grid `scale`; SNAP writes the grid values back
template <bool SNAP>
static int gi_split(float* A, int64_t lda, int64_t np, int ncols, int rows, const float2* scale, int8_t* planes,
                    const int* cond, cudaStream_t st) {
  const int64_t slabs = (np + GI_SLAB - 1) / GI_SLAB;
  FPS_REQUIRE(slabs <= (int64_t)65535 * 65535, "gram_i8: too many slabs");
  const unsigned gy = (unsigned)(slabs < 65535 ? slabs : 65535), gz = (unsigned)((slabs + 65534) / 65535);
  gi_split_kernel<SNAP><<<dim3((unsigned)(rows / 64), gy, gz), 256, 0, st>>>(A, lda, np, ncols, rows, (int)slabs, scale,
                                                                              planes, cond);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

static int gi_make_map(GiState* s, CUtensorMap* m, const int8_t* base, int64_t slabs, int rows, int box_rows, int rb) {
  cuuint64_t dims[2] = {(cuuint64_t)GI_SLAB, (cuuint64_t)slabs * GI_SL * rows};
  cuuint64_t strides[1] = {(cuuint64_t)GI_SLAB};
  cuuint32_t box[2] = {(cuuint32_t)rb, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((GiEncodeFn)s->encode)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FPS_REQUIRE(r == CUDA_SUCCESS, "gram_i8: cuTensorMapEncodeTiled failed (%d), slabs=%lld rows=%d", (int)r,
              (long long)slabs, rows);
  return 0;
}

// K split: items = ntiles x nsplit should fill whole waves of CTA pairs
static int gi_pick_split(int ntiles, int slabs, int pairs) {
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= 64 && ns <= slabs; ++ns) {
    if (ns > 1 && slabs / ns < 8) break;            // keep the TMEM drain amortised
    const int items = ntiles * ns;
    const int waves = (items + pairs - 1) / pairs;
    const double eff = (double)items / ((double)waves * pairs);
    if (eff > best_eff + 0.03) { best_eff = eff; best = ns; }
    if (waves >= 8) break;
  }
  return best;
}

static int gi_nitems(int ntiles, int64_t slabs128, int pairs, int rb, int* slabs_total, int* slabs_per_split) {
  *slabs_total = (int)(slabs128 * (GI_SLAB / rb));
  int nsplit = gi_pick_split(ntiles, *slabs_total, pairs);
  *slabs_per_split = (*slabs_total + nsplit - 1) / nsplit;
  nsplit = (*slabs_total + *slabs_per_split - 1) / *slabs_per_split;
  return ntiles * nsplit;
}

// one pass: C (M x Nc, ldc) += A^T B over the `slabs` 128-point slabs that start at slab_a0 of the A planes (which hold
// slabs_in_a slabs) and at 0 of the B planes
static int gi_run(fps_ctx* ctx, GiState* s, const GiOperand& pa, int64_t slabs_in_a, int64_t slab_a0, const GiOperand& pb,
                  int64_t slabs_in_b, int64_t slabs, bool sym, int M, int Nc, double* C, int64_t ldc, cudaStream_t st) {
  static const int rb = getenv("FPS_GRAM_ROWB") ? atoi(getenv("FPS_GRAM_ROWB")) : 128;
  FPS_REQUIRE(rb == 64 || rb == 128, "FPS_GRAM_ROWB must be 64 or 128");
  FPS_REQUIRE(slabs_in_a * GI_SL * pa.rows < ((int64_t)1 << 31) && slabs * (GI_SLAB / rb) < ((int64_t)1 << 31),
              "gram_i8: %lld slabs exceed the 32-bit tile coordinates", (long long)slabs_in_a);
  CUtensorMap tmA, tmB;
  if (gi_make_map(s, &tmA, pa.data, slabs_in_a, pa.rows, GI_TM, rb)) return 2;
  if (gi_make_map(s, &tmB, pb.data, slabs_in_b, pb.rows, GI_TN, rb)) return 2;
  GiParams P;
  P.sym = sym ? 1 : 0;
  P.tiles_j = (Nc + GI_TN - 1) / GI_TN;
  P.ntiles = sym ? GI_NTILES : 3 * P.tiles_j;
  P.rows_a = pa.rows;
  P.rows_b = pb.rows;
  P.slab_a0 = (int)slab_a0;
  const int pairs = ctx->sm_count / 2;
  P.nitems = gi_nitems(P.ntiles, slabs, pairs, rb, &P.slabs_total, &P.slabs_per_split);
  const int nsplit = P.nitems / P.ntiles;
  if (int rc = gi_reserve_partial(ctx, s, P.nitems)) return rc;
  P.partial = s->partial;
  const int groups = P.nitems < pairs ? P.nitems : pairs;
  if (rb == 128) gram_i8_kernel<128><<<groups * 2, GI_THREADS, GI_SMEM, st>>>(tmA, tmB, P);
  else gram_i8_kernel<64><<<groups * 2, GI_THREADS, GI_SMEM, st>>>(tmA, tmB, P);
  FPS_CUDA(cudaGetLastError());
  gi_reduce_kernel<<<dim3(P.ntiles, 8), 256, 0, st>>>(s->partial, nsplit, P.ntiles, P.sym, P.tiles_j, M, Nc, pa.exps,
                                                      pb.exps, C, ldc);
  FPS_CUDA(cudaGetLastError());
  count_launch(2);
  return 0;
}

size_t gram_planes_bytes(int64_t n) { return n > 0 ? gi_plane_bytes(GI_ROWS, n) : 0; }

// ---- cached form: the caller owns the descriptor `x` and the digit planes of ALL n rows ------------------------------
// G (FP x FP, ldg) += A^T A.  fused: the feature kernel has already written planes, snapped A and the calibrated grid
// into x (layer_tc.cu epilogue); only if it raised x->flags[0] (an element beyond the calibrated grid) are the
// exponents rebuilt from the recorded maxima and A re-split - both kernels are launched either way and return at once
// when there is nothing to do, so no host round trip is needed.  !fused: column maxima, exponents and split run here.
int launch_gram_x(fps_ctx* ctx, float* A, int64_t lda, int64_t n, Xop* x, int8_t* planes, bool fused, double* G,
                  int64_t ldg, cudaStream_t st) {
  if (n <= 0) return 0;
  GiState* s;
  if (int rc = gi_state_get(ctx, &s)) return rc;
  const int* cond = nullptr;
  if (fused) {
    cond = &x->flags[0];
  } else {
    if (int rc = xop_colmax(ctx, A, lda, n, FP, x->cmax, st)) return rc;
  }
  if (int rc = xop_exps(x->cmax, FP, GI_ROWS, 0, x->exps, x->scale, cond, st)) return rc;
  if (int rc = gi_split<true>(A, lda, n, FP, GI_ROWS, x->scale, planes, cond, st)) return rc;
  const int64_t slabs = (n + GI_SLAB - 1) / GI_SLAB;
  const GiOperand op{planes, x->exps, x->scale, GI_ROWS};
  return gi_run(ctx, s, op, slabs, 0, op, slabs, slabs, true, FP, FP, G, ldg, st);
}

// R (FP x B, ldr) += DH^T F from the cached planes of DH; F (n x B, ldf) is read only (its digit planes, on the grid of
// each right-hand side per pass, live in the context workspace)
int launch_project_x(fps_ctx* ctx, const Xop* x, const int8_t* planes, int64_t n, const float* F, int64_t ldf, int B,
                     double* R, int64_t ldr, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  GiState* s;
  if (int rc = gi_state_get(ctx, &s)) return rc;
  const int rows_b = (B + 63) / 64 * 64;
  const int64_t chunk = gi_project_chunk(rows_b, n, GI_MAX_CHUNK);
  if (int rc = gi_reserve_b(ctx, s, rows_b, chunk)) return rc;
  unsigned* b_cmax = s->b_meta;
  int* b_exps = reinterpret_cast<int*>(s->b_meta + s->b_rows_cap);
  float2* b_scale = reinterpret_cast<float2*>(s->b_meta + 2 * (size_t)s->b_rows_cap);   // 8-byte aligned: rows % 64 == 0
  const int64_t slabs_all = (n + GI_SLAB - 1) / GI_SLAB;
  const GiOperand pa{const_cast<int8_t*>(planes), x->exps, x->scale, GI_ROWS};
  const GiOperand pb{s->b_data, b_exps, b_scale, rows_b};
  for (int64_t p0 = 0; p0 < n; p0 += chunk) {
    const int64_t np = (n - p0 < chunk) ? n - p0 : chunk;
    const int64_t slabs = (np + GI_SLAB - 1) / GI_SLAB;
    float* Fc = const_cast<float*>(F) + p0 * ldf;
    if (int rc = xop_colmax(ctx, Fc, ldf, np, B, b_cmax, st)) return rc;
    if (int rc = xop_exps(b_cmax, B, rows_b, 0, b_exps, b_scale, nullptr, st)) return rc;
    if (int rc = gi_split<false>(Fc, ldf, np, B, rows_b, b_scale, s->b_data, nullptr, st)) return rc;
    if (int rc = gi_run(ctx, s, pa, slabs_all, p0 / GI_SLAB, pb, slabs, slabs, false, FP, B, R, ldr, st)) return rc;
  }
  return 0;
}

// ---- uncached forms (fps_gram_accum / fps_project_rhs on any matrix): planes in the context workspace, passes of at
// most 2^20 points, ONE grid for all n rows (column maxima over the whole matrix first) --------------------------------
int launch_gram_i8(fps_ctx* ctx, float* A, int64_t lda, int64_t n, double* G, int64_t ldg, cudaStream_t st) {
  if (n <= 0) return 0;
  const int64_t max_chunk = gi_env_chunk();
  FPS_REQUIRE(max_chunk >= GI_SLAB && (max_chunk & (max_chunk - 1)) == 0, "FPS_GRAM_I8_CHUNK must be a power of two >= 128");
  const int64_t chunk = n < max_chunk ? n : max_chunk;
  GiState* s;
  if (int rc = gi_state_get(ctx, &s)) return rc;
  if (int rc = gi_grow(ctx, &s->a_data, &s->a_cap, gi_plane_bytes(GI_ROWS, chunk))) return rc;
  Xop* x = ctx->xop_tmp;
  if (int rc = xop_colmax(ctx, A, lda, n, FP, x->cmax, st)) return rc;
  if (int rc = xop_exps(x->cmax, FP, GI_ROWS, 0, x->exps, x->scale, nullptr, st)) return rc;
  const GiOperand op{s->a_data, x->exps, x->scale, GI_ROWS};
  for (int64_t p0 = 0; p0 < n; p0 += chunk) {
    const int64_t np = (n - p0 < chunk) ? n - p0 : chunk;
    if (int rc = gi_split<true>(A + p0 * lda, lda, np, FP, GI_ROWS, x->scale, s->a_data, nullptr, st)) return rc;
    const int64_t slabs = (np + GI_SLAB - 1) / GI_SLAB;
    if (int rc = gi_run(ctx, s, op, slabs, 0, op, slabs, slabs, true, FP, FP, G, ldg, st)) return rc;
  }
  return 0;
}

int launch_project_i8(fps_ctx* ctx, float* DH, int64_t lda, const float* F, int64_t ldf, int B, int64_t n, double* R,
                      int64_t ldr, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  GiState* s;
  if (int rc = gi_state_get(ctx, &s)) return rc;
  const int rows_b = (B + 63) / 64 * 64;
  const int64_t max_chunk = gi_env_chunk();
  FPS_REQUIRE(max_chunk >= GI_SLAB && (max_chunk & (max_chunk - 1)) == 0, "FPS_GRAM_I8_CHUNK must be a power of two >= 128");
  const int64_t chunk = gi_project_chunk(rows_b, n, max_chunk);
  if (int rc = gi_grow(ctx, &s->a_data, &s->a_cap, gi_plane_bytes(GI_ROWS, chunk))) return rc;
  if (int rc = gi_reserve_b(ctx, s, rows_b, chunk)) return rc;
  unsigned* b_cmax = s->b_meta;
  int* b_exps = reinterpret_cast<int*>(s->b_meta + s->b_rows_cap);
  float2* b_scale = reinterpret_cast<float2*>(s->b_meta + 2 * (size_t)s->b_rows_cap);
  Xop* x = ctx->xop_tmp;
  if (int rc = xop_colmax(ctx, DH, lda, n, FP, x->cmax, st)) return rc;
  if (int rc = xop_exps(x->cmax, FP, GI_ROWS, 0, x->exps, x->scale, nullptr, st)) return rc;
  const GiOperand pa{s->a_data, x->exps, x->scale, GI_ROWS};
  const GiOperand pb{s->b_data, b_exps, b_scale, rows_b};
  for (int64_t p0 = 0; p0 < n; p0 += chunk) {
    const int64_t np = (n - p0 < chunk) ? n - p0 : chunk;
    const int64_t slabs = (np + GI_SLAB - 1) / GI_SLAB;
    float* Fc = const_cast<float*>(F) + p0 * ldf;
    if (int rc = gi_split<true>(DH + p0 * lda, lda, np, FP, GI_ROWS, x->scale, s->a_data, nullptr, st)) return rc;
    if (int rc = xop_colmax(ctx, Fc, ldf, np, B, b_cmax, st)) return rc;
    if (int rc = xop_exps(b_cmax, B, rows_b, 0, b_exps, b_scale, nullptr, st)) return rc;
    if (int rc = gi_split<false>(Fc, ldf, np, B, rows_b, b_scale, s->b_data, nullptr, st)) return rc;
    if (int rc = gi_run(ctx, s, pa, slabs, 0, pb, slabs, slabs, false, FP, B, R, ldr, st)) return rc;
  }
  return 0;
}

// size the context-owned buffers of the projection for n rows and B right-hand sides (fps_reserve_workspace)
int gram_i8_reserve(fps_ctx* ctx, int64_t n, int B) {
  GiState* s;
  if (int rc = gi_state_get(ctx, &s)) return rc;
  const int pairs = ctx->sm_count / 2;
  int st_, sp_;
  int nitems = gi_nitems(GI_NTILES, (n + GI_SLAB - 1) / GI_SLAB, pairs, 128, &st_, &sp_);
  if (B > 0) {
    const int rows_b = (B + 63) / 64 * 64;
    const int64_t chunk = gi_project_chunk(rows_b, n, GI_MAX_CHUNK);
    if (int rc = gi_reserve_b(ctx, s, rows_b, chunk)) return rc;
    nitems = std::max(nitems, gi_nitems(3 * (rows_b / 64), (chunk + GI_SLAB - 1) / GI_SLAB, pairs, 128, &st_, &sp_));
  }
  return gi_reserve_partial(ctx, s, nitems);
}

}  // namespace fps
