// Batched reconstruction  out[i, j] = sum_k A[i, k] W[k, j] + bias[j]   (solver.py:327-330 for B right-hand sides:
// u = H w + b, f_pred = DH w) on the INT8 tensor cores with exact products.
//
// Why not fp32-grade tensor arithmetic: for rough sources |w| reaches 1e5 while u = O(1), the 700 terms cancel to
// 1e-5 of their size, and the fp64-pipe kernel this replaces (28 TFLOP/s) exists for that reason.  Here:
//   A (n x 704 fp32; H or DH)   X[i, k] = rint(A[i, k] 2^(30 - e_k)), e_k from the column maximum, four balanced
//                               radix-256 digits; A is snapped in place to X 2^(e_k - 30) exactly like in
//                               gram_i8.cu (a no-op for a DH that went through the Gram); digit planes
//                               pa[d][i][k], feature axis contiguous = K-major, no transpose
//   W (704 x B fp64)            W'[k, j] = W[k, j] 2^(e_k - 30) (exact), Y[k, j] = rint(W' 2^(38 - g_j)) with g_j from the
//                               column maximum of W': a 39-bit integer, five digits; planes pw[d][j][k]
//   out[i, j] = 2^(g_j - 38) sum_k X[i, k] Y[k, j] + bias[j],   X Y = sum_{a, b} 2^(8 (7 - a - b)) (Wd_a . Ad_b)
// One tcgen05.mma.kind::i8 instruction multiplies one W digit (pair kernel: M = 256 right-hand sides per CTA pair; solo
// kernel, the default: M = 128 per CTA) with a stack of A digits of 64 points (pair: all four, N = 256; solo: the digits
// b <= 4 - a that are read back) into the TMEM window that starts at column 64 a, so that column group
// o collects the digit pairs with a + b = o (see gram_i8.cu).  K = 704 = 22 instructions of 32 per digit; the int32
// accumulators cannot overflow (5 pairs x 704 x 2^14 < 2^26).  Orders 0..4 are combined exactly in a 64-bit integer
// by the drain warps (orders 5..7 weigh < 2^-40 of the largest term and are not read), scaled, biased, stored as fp32.
// The rounding of W to 39 bits moves out by ~ sqrt(700) |w|_max 2^-39 |a|: 1e-9 of the largest term.
#include "fps_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <algorithm>

namespace fps {

constexpr int RI_TM = 128;                    // right-hand sides per CTA (UMMA M = 256 per pair)
constexpr int RI_TN = 64;                     // points per tile
constexpr int RI_DA = 4;                      // digits of A (stacked along UMMA N: 4 x 64 = 256)
constexpr int RI_DW = 5;                      // digits of W
constexpr int RI_NORD = 5;                    // digit orders read back (0..4)
constexpr int RI_W_BYTES = RI_TM * 128;       // 16 KiB per W digit per 128-byte K slab
constexpr int RI_A_BYTES = RI_TN * 128;       // 8 KiB per A digit per slab; a CTA stages two of the four
constexpr int RI_STAGE = RI_DW * RI_W_BYTES + (RI_DA / 2) * RI_A_BYTES;  // 96 KiB
constexpr int RI_STAGES = 2;
constexpr int RI_SMEM = RI_STAGES * RI_STAGE + 1024 + 256;
constexpr int RI_THREADS = 320;
constexpr int RI_KSLABS = (FP + 127) / 128;   // 6 slabs, the last one holds 64 bytes (TMA zero-fills the rest)

// ---- operand preparation ----------------------------------------------------------------------------
// digits of A, K-major (no transpose): thread = 4 consecutive features of one point
__global__ void __launch_bounds__(256)
ri_split_a_kernel(float* __restrict__ A, int64_t ld, int64_t n, int64_t n_pad, const float2* __restrict__ scale,
                  int8_t* __restrict__ planes) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = idx / (FP / 4);
  const int k = (int)(idx % (FP / 4)) * 4;
  if (i >= n_pad) return;
  uint32_t word[RI_DA] = {0u, 0u, 0u, 0u};
  if (i < n) {
    float4 v = *reinterpret_cast<const float4*>(A + i * ld + k);
    float vv[4] = {v.x, v.y, v.z, v.w};
    bool changed = false;
    uint32_t z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 sc = scale[k + u];
      const int X = (k + u < F) ? __float2int_rn(vv[u] * sc.x) : 0;
      const float snapped = (float)X * sc.y;
      if (snapped != vv[u]) { vv[u] = snapped; changed = true; }
      z[u] = balanced_digits4(X);
    }
    digits_transpose4(z[0], z[1], z[2], z[3], word);
    if (changed) *reinterpret_cast<float4*>(A + i * ld + k) = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
#pragma unroll
  for (int s = 0; s < RI_DA; ++s)
    *reinterpret_cast<uint32_t*>(planes + ((int64_t)s * n_pad + i) * FP + k) = word[s];
}

// column maxima of W'[k, j] = W[k, j] 2^(e_k - 30) -> exponent g_j with |W'[:, j]| < 2^g_j.
// Block = 32 right-hand sides x 8 feature groups (coalesced along j), reduced through shared memory.
__global__ void __launch_bounds__(256)
ri_wexp_kernel(const double* __restrict__ W, int64_t ldw, int B, int b_pad, const int* __restrict__ ea,
               int* __restrict__ gw) {
  __shared__ double sm[8][32];
  const int jl = threadIdx.x % 32, kg = threadIdx.x / 32;
  const int j = blockIdx.x * 32 + jl;
  double m = 0.0;
  if (j < B)
    for (int k = kg; k < F; k += 8) m = fmax(m, fabs(W[(int64_t)k * ldw + j]) * __longlong_as_double((long long)(1023 + ea[k] - 30) << 52));
  sm[kg][jl] = m;
  __syncthreads();
  if (kg == 0 && j < b_pad) {
#pragma unroll
    for (int q = 1; q < 8; ++q) m = fmax(m, sm[q][jl]);
    int g = 0;
    if (m > 0.0) {
      int ex;
      frexp(m, &ex);   // m = f 2^ex, 0.5 <= f < 1  ->  m < 2^ex
      g = ex;
    }
    gw[j] = g;
  }
}

// digits of W', transposed to K-major: thread = (right-hand side j, 4 consecutive features)
__global__ void __launch_bounds__(256)
ri_split_w_kernel(const double* __restrict__ W, int64_t ldw, int B, int b_pad, const int* __restrict__ ea,
                  const int* __restrict__ gw, int8_t* __restrict__ planes) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y * 4;
  if (j >= b_pad) return;
  uint32_t word[RI_DW] = {0u, 0u, 0u, 0u, 0u};
  if (j < B) {
    const int g = gw[j];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      long long Y = 0;
      if (k + u < F) Y = __double2ll_rn(ldexp(W[(int64_t)(k + u) * ldw + j], ea[k + u] - 30 + 38 - g));
#pragma unroll
      for (int s = RI_DW - 1; s >= 1; --s) {
        const long long d = ((Y + 128) & 255) - 128;
        Y = (Y - d) >> 8;
        word[s] |= (uint32_t)(d & 255) << (8 * u);
      }
      word[0] |= (uint32_t)(Y & 255) << (8 * u);
    }
  }
#pragma unroll
  for (int s = 0; s < RI_DW; ++s)
    *reinterpret_cast<uint32_t*>(planes + ((int64_t)s * b_pad + j) * FP + k) = word[s];
}

// ---- the tensor-core kernel ---------------------------------------------------------------------------
__device__ __forceinline__ void umma_i8_256(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ri_tmem_ld16(uint32_t taddr, int* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
struct RiParams {
  float* out;            // (n, ldo) fp32
  int64_t ldo;
  int64_t n;             // valid points of this pass
  int B;                 // valid right-hand sides
  const int* gw;         // exponent per right-hand side
  const double* bias;    // per right-hand side or null
  int rhs_blocks;        // b_pad / 256
  int point_blocks;      // n_pad / 64
  int n_pad, b_pad;      // rows per digit plane of A / W
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RI_THREADS, 1)
recon_i8_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const RiParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + RI_STAGES * RI_STAGE;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 32, bar_tfull = bar_base + 64, bar_tempty = bar_base + 72;
  const uint32_t tmem_slot = bar_base + 80;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int rank = (int)cluster_ctarank();
  const int group = blockIdx.x / 2, ngroups = gridDim.x / 2;
  const int ntiles = P.rhs_blocks * P.point_blocks;   // tile t: rhs block t % rhs_blocks, point block t / rhs_blocks

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    for (int s = 0; s < RI_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 16);
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
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = group; t < ntiles; t += ngroups) {
        const int jb = t % P.rhs_blocks, ib = t / P.rhs_blocks;
        const int wrow = jb * 2 * RI_TM + rank * RI_TM, arow = ib * RI_TN;
        for (int kb = 0; kb < RI_KSLABS; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * RI_STAGE;
          const uint32_t full = bar_full + 8 * stage;
          if (rank == 0) mbar_expect_tx(full, 2 * RI_STAGE);
#pragma unroll
          for (int s = 0; s < RI_DW; ++s) tma_load_2d<2>(sa + s * RI_W_BYTES, &tmW, full, kb * 128, s * P.b_pad + wrow);
#pragma unroll
          for (int s = 0; s < RI_DA / 2; ++s)   // stack rows [128 rank, 128 rank + 128) = A digits 2 rank, 2 rank + 1
            tma_load_2d<2>(sa + RI_DW * RI_W_BYTES + s * RI_A_BYTES, &tmA, full, kb * 128, (2 * rank + s) * P.n_pad + arow);
          if (++stage == RI_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (pair leader) =================
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, tphase = 0;
      for (int t = group; t < ntiles; t += ngroups) {
        mbar_wait(bar_tempty, tphase);   // accumulators read and zeroed
        tphase ^= 1;
        tc_fence_after();
        for (int kb = 0; kb < RI_KSLABS; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * RI_STAGE;
          const uint64_t db = umma_desc<128>(sa + RI_DW * RI_W_BYTES);
          const int ksteps = (kb == RI_KSLABS - 1) ? (FP - 128 * (RI_KSLABS - 1)) / 32 : 4;
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t adv = (uint64_t)((j * 32) >> 4);
            // The very first step of a tile overwrites: W digit 4 initialises columns [256, 512), digit 0 columns
            // [0, 256); the windows of digits 1..3 then only touch initialised columns.  No zeroing pass needed.
            const uint32_t keep = (kb > 0 || j > 0) ? 1u : 0u;
            umma_i8_256(tmem_base + 4 * RI_TN, umma_desc<128>(sa + 4 * RI_W_BYTES) + adv, db + adv, keep);
            umma_i8_256(tmem_base, umma_desc<128>(sa) + adv, db + adv, keep);
#pragma unroll
            for (int a = 1; a < RI_DW - 1; ++a)
              umma_i8_256(tmem_base + a * RI_TN, umma_desc<128>(sa + a * RI_W_BYTES) + adv, db + adv, 1u);
          }
          umma_commit<2>(bar_empty + 8 * stage);
          if (++stage == RI_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<2>(bar_tfull);
      }
    }
  } else {
    // ================= drain warps 2..9 =================
    const int quad = warp % 4;          // TMEM lanes [32 quad, 32 quad + 32) = right-hand sides
    const int half = (warp - 2) / 4;    // 32 of the 64 points
    const uint32_t taddr = tmem_base + half * (RI_TN / 2) + ((uint32_t)(quad * 32) << 16);
    if (lane == 0) mbar_arrive_leader(bar_tempty);   // the accumulators are free from the start
    uint32_t tphase = 0;
    for (int t = group; t < ntiles; t += ngroups) {
      const int jb = t % P.rhs_blocks, ib = t / P.rhs_blocks;
      const int j = jb * 2 * RI_TM + rank * RI_TM + quad * 32 + lane;
      const bool j_ok = j < P.B;
      // out = 2^g sum_o 2^(-8 o) Acc_o,  g = g_j - 38 + 56.  The orders 0..4 are combined exactly in a 64-bit integer
      // S = sum_o 2^(8 (4 - o)) Acc_o  (|Acc_o| < 2^26, so |S| < 2^59), converted once: out = S 2^(g - 32) + bias.
      // (Five int->double conversions and a fp64 Horner chain per output made the drain fp64-pipe bound.)
      const int g = j_ok ? P.gw[j] - 38 + 8 * (RI_DA + RI_DW - 2) : 0;
      const double scale = __longlong_as_double((long long)(1023 + g - 32) << 52);
      const double bj = (j_ok && P.bias) ? P.bias[j] : 0.0;
      const int64_t i0 = (int64_t)ib * RI_TN + half * (RI_TN / 2);
      mbar_wait(bar_tfull, tphase);
      tphase ^= 1;
      tc_fence_after();
      // Phase 1 (TMEM-read bound): pull the five orders out and fold them into one 64-bit integer per output, then hand
      // the accumulators back to the MMA issuer at once; phase 2 (conversion, scaling, stores) overlaps the next tile.
      long long S[RI_TN / 2];
#pragma unroll
      for (int c = 0; c < RI_TN / 32; ++c) {
        int v[RI_NORD][16];
#pragma unroll
        for (int o = 0; o < RI_NORD; ++o) ri_tmem_ld16(taddr + o * RI_TN + c * 16, v[o]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          long long acc = v[0][q];
#pragma unroll
          for (int o = 1; o < RI_NORD; ++o) acc = acc * 256 + v[o][q];
          S[c * 16 + q] = acc;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_tempty);
#pragma unroll
      for (int q = 0; q < RI_TN / 2; ++q) {
        const int64_t i = i0 + q;
        if (j_ok && i < P.n) P.out[i * P.ldo + j] = (float)fma((double)S[q], scale, bj);
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

// ---- single-CTA form: only the digit pairs that are read back ---------------------------------------------------------
// Orders a + b <= 4 are all the drain reads, yet the pair kernel above multiplies every W digit with the full stack of four
// A digits (20 products, 6 of them - orders 5..7 - for nothing): under the 1 kW power cap the kernel runs at the chip's
// sustained int8 rate, so wasted products are wasted time.  Trimming the stack per W digit (N = 256, 256, 192, 128, 64:
// 14 products) needs the stack's PREFIX at one shared-memory address, which a cta_group::2 operand cannot offer (its N
// halves come from the two CTAs and move with N).  Here every CTA works alone (cta_group::1, M = 128 right-hand sides)
// and stages all four A digits itself: [d0 | d1 | d2 | d3], 32 KiB per K slab next to the five W digits (80 KiB).
constexpr int RS_STAGE = RI_DW * RI_W_BYTES + RI_DA * RI_A_BYTES;   // 112 KiB
constexpr int RS_SMEM = RI_STAGES * RS_STAGE + 1024 + 256;

template <int N>
__device__ __forceinline__ void umma_i8_128(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(RI_THREADS, 1)
recon_i8_solo_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const RiParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + RI_STAGES * RS_STAGE;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 32, bar_tfull = bar_base + 64, bar_tempty = bar_base + 72;
  const uint32_t tmem_slot = bar_base + 80;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int rhs_blocks = 2 * P.rhs_blocks;            // blocks of 128 right-hand sides
  const int ntiles = rhs_blocks * P.point_blocks;     // tile t: rhs block t % rhs_blocks, point block t / rhs_blocks

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    for (int s = 0; s < RI_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int jb = t % rhs_blocks, ib = t / rhs_blocks;
        const int wrow = jb * RI_TM, arow = ib * RI_TN;
        for (int kb = 0; kb < RI_KSLABS; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * RS_STAGE;
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, RS_STAGE);
#pragma unroll
          for (int s = 0; s < RI_DW; ++s) tma_load_2d<1>(sa + s * RI_W_BYTES, &tmW, full, kb * 128, s * P.b_pad + wrow);
#pragma unroll
          for (int s = 0; s < RI_DA; ++s)
            tma_load_2d<1>(sa + RI_DW * RI_W_BYTES + s * RI_A_BYTES, &tmA, full, kb * 128, s * P.n_pad + arow);
          if (++stage == RI_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, tphase = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(bar_tempty, tphase);   // accumulators read
        tphase ^= 1;
        tc_fence_after();
        for (int kb = 0; kb < RI_KSLABS; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * RS_STAGE;
          const uint64_t db = umma_desc<128>(sa + RI_DW * RI_W_BYTES);    // stack [d0 | d1 | d2 | d3], 64 rows each
          const int ksteps = (kb == RI_KSLABS - 1) ? (FP - 128 * (RI_KSLABS - 1)) / 32 : 4;
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t adv = (uint64_t)((j * 32) >> 4);
            // W digit a times the A digits b <= 4 - a, into the window that starts at column 64 a: column group o
            // collects a + b = o.  The first step of a tile overwrites: digit 0 initialises columns [0, 256), digit 4
            // columns [256, 320); the windows of digits 1..3 then only touch initialised columns.
            const uint32_t keep = (kb > 0 || j > 0) ? 1u : 0u;
            umma_i8_128<256>(tmem_base, umma_desc<128>(sa) + adv, db + adv, keep);
            umma_i8_128<64>(tmem_base + 4 * RI_TN, umma_desc<128>(sa + 4 * RI_W_BYTES) + adv, db + adv, keep);
            umma_i8_128<256>(tmem_base + 1 * RI_TN, umma_desc<128>(sa + 1 * RI_W_BYTES) + adv, db + adv, 1u);
            umma_i8_128<192>(tmem_base + 2 * RI_TN, umma_desc<128>(sa + 2 * RI_W_BYTES) + adv, db + adv, 1u);
            umma_i8_128<128>(tmem_base + 3 * RI_TN, umma_desc<128>(sa + 3 * RI_W_BYTES) + adv, db + adv, 1u);
          }
          umma_commit<1>(bar_empty + 8 * stage);
          if (++stage == RI_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<1>(bar_tfull);
      }
    }
  } else {
    // ================= drain warps 2..9 =================
    const int quad = warp % 4;          // TMEM lanes [32 quad, 32 quad + 32) = right-hand sides
    const int half = (warp - 2) / 4;    // 32 of the 64 points
    const uint32_t taddr = tmem_base + half * (RI_TN / 2) + ((uint32_t)(quad * 32) << 16);
    if (lane == 0) mbar_arrive(bar_tempty);   // the accumulators are free from the start
    uint32_t tphase = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int jb = t % rhs_blocks, ib = t / rhs_blocks;
      const int j = jb * RI_TM + quad * 32 + lane;
      const bool j_ok = j < P.B;
      const int g = j_ok ? P.gw[j] - 38 + 8 * (RI_DA + RI_DW - 2) : 0;
      const double scale = __longlong_as_double((long long)(1023 + g - 32) << 52);
      const double bj = (j_ok && P.bias) ? P.bias[j] : 0.0;
      const int64_t i0 = (int64_t)ib * RI_TN + half * (RI_TN / 2);
      mbar_wait(bar_tfull, tphase);
      tphase ^= 1;
      tc_fence_after();
      long long S[RI_TN / 2];
#pragma unroll
      for (int c = 0; c < RI_TN / 32; ++c) {
        int v[RI_NORD][16];
#pragma unroll
        for (int o = 0; o < RI_NORD; ++o) ri_tmem_ld16(taddr + o * RI_TN + c * 16, v[o]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          long long acc = v[0][q];
#pragma unroll
          for (int o = 1; o < RI_NORD; ++o) acc = acc * 256 + v[o][q];
          S[c * 16 + q] = acc;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
#pragma unroll
      for (int q = 0; q < RI_TN / 2; ++q) {
        const int64_t i = i0 + q;
        if (j_ok && i < P.n) P.out[i * P.ldo + j] = (float)fma((double)S[q], scale, bj);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host -----------------------------------------------------------------------------------------------
struct RiState {
  int8_t* pa = nullptr;      // context-owned digit planes of A (uncached entry point): RI_DA x n_pad x FP
  size_t pa_bytes = 0;
  int8_t* pw = nullptr;      // digit planes of W': RI_DW x b_pad x FP
  size_t pw_bytes = 0;
  int* gw = nullptr;         // b_pad exponents
  int gw_cap = 0;
  void* encode = nullptr;
};

typedef CUresult (*RiEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void recon_i8_destroy(fps_ctx* ctx) {
  RiState* s = static_cast<RiState*>(ctx->ri_state);
  if (!s) return;
  cudaFree(s->pa);
  cudaFree(s->pw);
  cudaFree(s->gw);
  delete s;
  ctx->ri_state = nullptr;
}

static int ri_state_get(fps_ctx* ctx, RiState** out) {
  RiState* s = static_cast<RiState*>(ctx->ri_state);
  if (!s) {
    s = new RiState();
    ctx->ri_state = s;
    cudaDriverEntryPointQueryResult qres;
    FPS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &s->encode, cudaEnableDefault, &qres));
    FPS_REQUIRE(s->encode && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    FPS_CUDA(cudaFuncSetAttribute(recon_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RI_SMEM));
    FPS_CUDA(cudaFuncSetAttribute(recon_i8_solo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
  }
  *out = s;
  return 0;
}

static int ri_map(RiState* s, CUtensorMap* m, const int8_t* base, int64_t rows, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)FP, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)FP};
  cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((RiEncodeFn)s->encode)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FPS_REQUIRE(r == CUDA_SUCCESS, "recon_i8: cuTensorMapEncodeTiled failed (%d), rows=%lld", (int)r, (long long)rows);
  return 0;
}

constexpr int64_t RI_MAX_CHUNK = 1 << 20;   // points per pass through context-owned planes: 2.75 GiB

static int64_t ri_pad(int64_t n) { return (n + RI_TN - 1) / RI_TN * RI_TN; }
size_t recon_planes_bytes(int64_t n) { return n > 0 ? (size_t)RI_DA * ri_pad(n) * FP : 0; }

static int ri_reserve_w(fps_ctx* ctx, RiState* s, int b_pad) {
  if ((size_t)RI_DW * b_pad * FP > s->pw_bytes) {
    if (s->pw) { FPS_CUDA(cudaFree(s->pw)); ctx->workspace_bytes -= s->pw_bytes; s->pw = nullptr; s->pw_bytes = 0; }
    FPS_CUDA(cudaMalloc(&s->pw, (size_t)RI_DW * b_pad * FP));
    s->pw_bytes = (size_t)RI_DW * b_pad * FP;
    ctx->workspace_bytes += s->pw_bytes;
  }
  if (b_pad > s->gw_cap) {
    if (s->gw) FPS_CUDA(cudaFree(s->gw));
    FPS_CUDA(cudaMalloc(&s->gw, (size_t)b_pad * sizeof(int)));
    s->gw_cap = b_pad;
  }
  return 0;
}

int recon_i8_reserve(fps_ctx* ctx, int B) {
  RiState* s;
  if (int rc = ri_state_get(ctx, &s)) return rc;
  if (B <= 0) return 0;
  return ri_reserve_w(ctx, s, (B + 2 * RI_TM - 1) / (2 * RI_TM) * (2 * RI_TM));
}

// digit planes of np rows of A on the grid of x (written to `planes` = RI_DA x ri_pad(np) x FP); A is snapped in place
// (nothing is written for an A that already lies on the grid)
static int ri_split_a(float* A, int64_t lda, int64_t np, const Xop* x, int8_t* planes, cudaStream_t st) {
  const int64_t n_pad = ri_pad(np);
  const int64_t threads = n_pad * (FP / 4);
  ri_split_a_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(A, lda, np, n_pad, x->scale, planes);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

// out (np x B, ldo) = X W + bias from digit planes of np rows
static int ri_run(fps_ctx* ctx, RiState* s, const Xop* x, const int8_t* planes, int64_t np, const double* W, int64_t ldw,
                  const double* bias, int B, float* out, int64_t ldo, cudaStream_t st) {
  const int b_pad = (B + 2 * RI_TM - 1) / (2 * RI_TM) * (2 * RI_TM);
  if (int rc = ri_reserve_w(ctx, s, b_pad)) return rc;
  const int64_t n_pad = ri_pad(np);
  FPS_REQUIRE((int64_t)RI_DA * n_pad < ((int64_t)1 << 31), "recon_i8: %lld points exceed the 32-bit tile coordinates",
              (long long)np);
  ri_wexp_kernel<<<(b_pad + 31) / 32, 256, 0, st>>>(W, ldw, B, b_pad, x->exps, s->gw);
  ri_split_w_kernel<<<dim3((b_pad + 255) / 256, FP / 4), 256, 0, st>>>(W, ldw, B, b_pad, x->exps, s->gw, s->pw);
  FPS_CUDA(cudaGetLastError());
  CUtensorMap tmW, tmA;
  if (ri_map(s, &tmW, s->pw, (int64_t)RI_DW * b_pad, RI_TM)) return 2;
  if (ri_map(s, &tmA, planes, (int64_t)RI_DA * n_pad, RI_TN)) return 2;
  RiParams P;
  P.out = out;
  P.ldo = ldo;
  P.n = np;
  P.B = B;
  P.gw = s->gw;
  P.bias = bias;
  P.rhs_blocks = b_pad / (2 * RI_TM);
  P.point_blocks = (int)(n_pad / RI_TN);
  P.n_pad = (int)n_pad;
  P.b_pad = b_pad;
  const int pairs = ctx->sm_count / 2;
  const int64_t ntiles = (int64_t)P.rhs_blocks * P.point_blocks;
  FPS_REQUIRE(ntiles < ((int64_t)1 << 31), "recon_i8: too many tiles");
  const int groups = (int)std::min<int64_t>(ntiles, pairs);
  // default: the single-CTA kernel that issues only the 14 digit pairs that are read back (see above; 4 096 RHS x 262 144
  // points, H and DH: 21.4 -> 17.0 ms); FPS_RECON_SOLO=0 selects the pair kernel with all 20 (A/B measurements)
  static const int solo = getenv("FPS_RECON_SOLO") ? atoi(getenv("FPS_RECON_SOLO")) : 1;
  if (solo) {
    const int ctas = (int)std::min<int64_t>(2 * ntiles, ctx->sm_count);
    recon_i8_solo_kernel<<<ctas, RI_THREADS, RS_SMEM, st>>>(tmW, tmA, P);
  } else {
    recon_i8_kernel<<<groups * 2, RI_THREADS, RI_SMEM, st>>>(tmW, tmA, P);
  }
  FPS_CUDA(cudaGetLastError());
  count_launch(3);
  return 0;
}

// ---- cached form: the caller owns the descriptor and the planes of all n rows -------------------------------------------
// need_grid: derive the grid of x from the column maxima of A first (false: x already holds it, e.g. from the feature
// kernel or from the Gram of the same matrix)
int launch_recon_planes_build(fps_ctx* ctx, float* A, int64_t lda, int64_t n, Xop* x, int8_t* planes, bool need_grid,
                              cudaStream_t st) {
  if (n <= 0) return 0;
  RiState* s;
  if (int rc = ri_state_get(ctx, &s)) return rc;
  if (need_grid) {
    if (int rc = xop_colmax(ctx, A, lda, n, FP, x->cmax, st)) return rc;
    if (int rc = xop_exps(x->cmax, FP, XOP_ROWS, 0, x->exps, x->scale, nullptr, st)) return rc;
  }
  return ri_split_a(A, lda, n, x, planes, st);
}

int launch_recon_x(fps_ctx* ctx, const Xop* x, const int8_t* planes, int64_t n, const double* W, int64_t ldw,
                   const double* bias, int B, float* out, int64_t ldo, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  RiState* s;
  if (int rc = ri_state_get(ctx, &s)) return rc;
  return ri_run(ctx, s, x, planes, n, W, ldw, bias, B, out, ldo, st);
}

// ---- uncached form (fps_reconstruct on any matrix): planes in the context workspace, passes of <= 2^20 points, one grid
// for all n rows; A is snapped in place (see the header) ------------------------------------------------------------------
int launch_recon_i8(fps_ctx* ctx, float* A, int64_t lda, int64_t n, const double* W, int64_t ldw, const double* bias,
                    int B, float* out, int64_t ldo, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  RiState* s;
  if (int rc = ri_state_get(ctx, &s)) return rc;
  static const int64_t max_chunk = getenv("FPS_GRAM_I8_CHUNK") ? atoll(getenv("FPS_GRAM_I8_CHUNK")) : RI_MAX_CHUNK;
  const int64_t chunk = n < max_chunk ? n : max_chunk;
  const size_t need = recon_planes_bytes(chunk);
  if (need > s->pa_bytes) {
    if (s->pa) { FPS_CUDA(cudaFree(s->pa)); ctx->workspace_bytes -= s->pa_bytes; s->pa = nullptr; s->pa_bytes = 0; }
    FPS_CUDA(cudaMalloc(&s->pa, need));
    s->pa_bytes = need;
    ctx->workspace_bytes += need;
  }
  Xop* x = ctx->xop_tmp;
  if (int rc = xop_colmax(ctx, A, lda, n, FP, x->cmax, st)) return rc;
  if (int rc = xop_exps(x->cmax, FP, XOP_ROWS, 0, x->exps, x->scale, nullptr, st)) return rc;
  for (int64_t p0 = 0; p0 < n; p0 += chunk) {
    const int64_t np = (n - p0 < chunk) ? n - p0 : chunk;
    if (int rc = ri_split_a(A + p0 * lda, lda, np, x, s->pa, st)) return rc;
    if (int rc = ri_run(ctx, s, x, s->pa, np, W, ldw, bias, B, out + p0 * ldo, ldo, st)) return rc;
  }
  return 0;
}

}  // namespace fps
