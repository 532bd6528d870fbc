// Dense layer + tanh with forward-mode derivative streams on the 5th-gen tensor cores
// (FPS_ENGINE_TCGEN05): TMA -> 128B-swizzled shared memory -> tcgen05.mma.kind::tf32 (3xTF32 split)
// -> fp32 accumulators in TMEM -> tcgen05.ld -> fused tanh/derivative epilogue -> next layer's
// hi/lo planes (or H / DH for the last layer).
//
// Reference maths: nn.Linear + nn.Tanh (model.py:53-60, :144) and the Laplacian that
// utils.py:54-68 extracts by 2 800 autograd sweeps, here as closed-form streams.
//
// GEMM mapping (per CTA tile):
//   D[f, r] = sum_k W[f, k] * Act[r, k]      f: 128 output features  -> UMMA M  (TMEM lanes)
//                                            r: 256 activation rows  -> UMMA N  (TMEM columns)
//   A operand = W tile   (128 x 32 fp32, K-major, SWIZZLE_128B)  hi and lo planes
//   B operand = Act tile (256 x 32 fp32, K-major, SWIZZLE_128B)  hi and lo planes
//   Activation row r = 4*point + stream, so one thread (= one TMEM lane = one feature) finds the
//   four streams (v, v_x, v_y, v_lap) of a point in four adjacent TMEM columns and can apply the
//   tanh recurrences without any cross-thread exchange.
//   3xTF32:  D += Alo*Bhi + Ahi*Blo + Ahi*Bhi per 8-wide k step (hi = fp32 with the low 13 mantissa
//   bits cleared, lo = exact remainder; both planes are produced by the previous layer's epilogue).
//
//
// Accumulation accuracy.  The tensor core truncates (rounds toward zero) when it writes the fp32
// accumulator, so a single 704-long chain of 264 MMAs biases every output by ~3e-6 per layer
// (measured: H off by 1.6e-5 after six layers).  The K loop is therefore cut into segments of
// SEG_CHUNKS x 32: each segment accumulates in TMEM from zero, then the accumulate warps pull it
// out with tcgen05.ld and add it to a round-to-nearest fp32 running sum held in registers while the
// tensor core already works on the next segment in the other TMEM buffer.
//
// Warp roles (320 threads, persistent CTAs, one per SM):
//   warp 0     TMA producer            (2-stage ring of {Ahi, Alo, Bhi, Blo} = 96 KiB per stage)
//   warp 1     TMEM allocator + MMA issuer (one elected lane)
//   warps 2-9  accumulate + epilogue; TMEM lane quadrant = warp_id % 4, column half = (warp_id-2)/4,
//              128 running sums per thread.  The 512 TMEM columns hold two 256-column segment
//              accumulators (double buffer between MMA issuer and accumulate warps).
#include "fps_common.cuh"
#include <cuda.h>

namespace fps {

constexpr int TM = 128;                 // features per tile  (UMMA M)
constexpr int TN = 256;                 // activation rows per tile (UMMA N)
#ifndef FPS_TK
#define FPS_TK 16
#endif
constexpr int TK = FPS_TK;              // fp32 per K chunk: 16 -> 64-byte swizzle rows, 32 -> 128-byte
constexpr int UK = 8;                   // K per tcgen05.mma.kind::tf32
constexpr int TC_STAGES = 64 / TK;      // ring depth: 4 x 48 KiB (TK = 16) or 2 x 96 KiB (TK = 32)
constexpr int ROW_BYTES = TK * 4;       // shared-memory row = swizzle span
constexpr int A_BYTES = TM * TK * 4;    // 16 KiB
constexpr int B_BYTES = TN * TK * 4;    // 32 KiB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 96 KiB
constexpr int TC_THREADS = 320;       // TMA warp, MMA warp, 8 accumulate/epilogue warps
constexpr int SEG_CHUNKS = 32 / TK;   // K chunks accumulated inside the tensor core before promotion (K = 32)
constexpr int TMEM_COLS = 512;
constexpr int N_FEAT_TILES = FPW / TM;  // 6
constexpr int TC_SMEM = TC_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// ---- PTX wrappers -------------------------------------------------------------------------------
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
// Bounded wait: a protocol bug must trap, not hang the GPU (a hung box is a lost lease).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 200000000u) {
      printf("fps layer_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major swizzled shared-memory matrix descriptor: rows of ROW_BYTES (= swizzle span), 8-row groups
// 8*ROW_BYTES apart.  layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);         // start address, 16-byte units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;     // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61; // swizzle mode
  return d;
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC_TF32), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcParams {
  const float* bias;
  float* out_hi;
  float* out_lo;
  float* H;
  float* DH;
  int64_t ld;
  int64_t rows;   // valid activation rows = points * S
  int nk;         // K chunks of 32
  int row_tiles;  // ceil(rows / 256)
};

template <int S, bool LAST>
__global__ void __launch_bounds__(TC_THREADS, 1)
layer_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC_STAGES * STAGE_BYTES;
  // barrier slots (8 bytes each): full[4], empty[4], tmem_full[2], tmem_empty[2], then the TMEM base word
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 32, bar_tfull = bar_base + 64, bar_tempty = bar_base + 80;
  const uint32_t tmem_slot = bar_base + 96;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int total_tiles = P.row_tiles * N_FEAT_TILES;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWlo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAlo)) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 8);  // one arrive per accumulate warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
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
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ft = t % N_FEAT_TILES, rt = t / N_FEAT_TILES;
        for (int kc = 0; kc < P.nk; ++kc) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, STAGE_BYTES);
          tma_load_2d(sa, &tmWhi, full, kc * TK, ft * TM);
          tma_load_2d(sa + A_BYTES, &tmWlo, full, kc * TK, ft * TM);
          tma_load_2d(sa + 2 * A_BYTES, &tmAhi, full, kc * TK, rt * TN);
          tma_load_2d(sa + 2 * A_BYTES + B_BYTES, &tmAlo, full, kc * TK, rt * TN);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        for (int kc0 = 0; kc0 < P.nk; kc0 += SEG_CHUNKS) {
          const int kc1 = (kc0 + SEG_CHUNKS < P.nk) ? kc0 + SEG_CHUNKS : P.nk;
          mbar_wait(bar_tempty + 8 * abuf, aphase ^ 1);  // accumulate warps have drained this buffer
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + abuf * TN;
          for (int kc = kc0; kc < kc1; ++kc) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + A_BYTES);
            const uint64_t b_hi = umma_desc_sw128(sa + 2 * A_BYTES),
                           b_lo = umma_desc_sw128(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int j = 0; j < TK / UK; ++j) {
              const uint64_t adv = (uint64_t)((j * UK * 4) >> 4);  // +32 bytes per k step inside the swizzle atom
              umma_tf32(d_tmem, a_lo + adv, b_hi + adv, (kc > kc0 || j > 0) ? 1u : 0u);
              umma_tf32(d_tmem, a_hi + adv, b_lo + adv, 1u);
              umma_tf32(d_tmem, a_hi + adv, b_hi + adv, 1u);
            }
            umma_commit(bar_empty + 8 * stage);  // smem slot is free once these MMAs retire
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_tfull + 8 * abuf);  // segment accumulator complete
          if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ================= accumulate + epilogue (warps 2..9) =================
    const int quad = warp % 4;          // TMEM lanes [32*quad, 32*quad+32) are the only ones this warp may read
    const int half = (warp - 2) / 4;    // which 128 of the 256 accumulator columns
    uint32_t abuf = 0, aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int ft = t % N_FEAT_TILES, rt = t / N_FEAT_TILES;
      const int f = ft * TM + quad * 32 + lane;  // this thread's output feature
      const bool f_ok = f < FP;                  // features 704..767 are tile padding
      const float bias = P.bias[f];
      float acc[TN / 2];
#pragma unroll
      for (int i = 0; i < TN / 2; ++i) acc[i] = 0.f;
      for (int kc0 = 0; kc0 < P.nk; kc0 += SEG_CHUNKS) {
        mbar_wait(bar_tfull + 8 * abuf, aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + abuf * TN + half * (TN / 2) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c = 0; c < TN / 64; ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c * 32 + i] += v[i];  // round-to-nearest promotion
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * abuf);
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
      }
      const int64_t r0 = (int64_t)rt * TN + half * (TN / 2);
      if (S == 4) {
#pragma unroll
        for (int q = 0; q < TN / 8; ++q) {
          const int64_t r = r0 + q * 4;
          float h, hx, hy, hl;
          tanh_streams(acc[4 * q] + bias, acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3], h, hx, hy, hl);
          if (f_ok && r < P.rows) {
            if (LAST) {
              const int64_t p = r >> 2;
              P.H[p * P.ld + f] = h;
              P.DH[p * P.ld + f] = hl;
            } else {
              const float o[4] = {h, hx, hy, hl};
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const float hi = tf32_hi(o[s]);
                P.out_hi[(r + s) * FP + f] = hi;
                P.out_lo[(r + s) * FP + f] = o[s] - hi;
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < TN / 2; ++q) {
          const int64_t r = r0 + q;
          const float h = tanhf(acc[q] + bias);
          if (f_ok && r < P.rows) {
            if (LAST) {
              P.H[r * P.ld + f] = h;
            } else {
              const float hi = tf32_hi(h);
              P.out_hi[r * FP + f] = hi;
              P.out_lo[r * FP + f] = h - hi;
            }
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side: tensor maps -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  CUtensorMap w_hi[NLAYER], w_lo[NLAYER];
  CUtensorMap act_hi[2][2], act_lo[2][2];  // [buffer][0: K = K0P, 1: K = FP]
};

static int make_map(EncodeTiledFn enc, CUtensorMap* m, const float* base, uint64_t K, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for K=%llu rows=%llu", (int)r, (unsigned long long)K,
              (unsigned long long)rows);
    return 1;
  }
  return 0;
}

int tc_init(fps_ctx* ctx) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FPS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  FPS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  TcState* s = new TcState();
  ctx->tc_state = s;
  for (int l = 0; l < NLAYER; ++l) {
    const Layer& L = ctx->layer[l];
    if (make_map(enc, &s->w_hi[l], L.w_hi, L.Kp, FPW, TM)) return 1;
    if (make_map(enc, &s->w_lo[l], L.w_lo, L.Kp, FPW, TM)) return 1;
  }
  const uint64_t cap_rows = (uint64_t)ctx->chunk * NSTREAM;
  for (int b = 0; b < 2; ++b) {
    // a plane holds cap_rows x FP floats; viewed with K = K0P it has room for more rows, but the same
    // row capacity is all a wave ever uses
    if (make_map(enc, &s->act_hi[b][0], ctx->act[b].hi, K0P, cap_rows, TN)) return 1;
    if (make_map(enc, &s->act_lo[b][0], ctx->act[b].lo, K0P, cap_rows, TN)) return 1;
    if (make_map(enc, &s->act_hi[b][1], ctx->act[b].hi, FP, cap_rows, TN)) return 1;
    if (make_map(enc, &s->act_lo[b][1], ctx->act[b].lo, FP, cap_rows, TN)) return 1;
  }
  FPS_CUDA(cudaFuncSetAttribute(layer_tc_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  FPS_CUDA(cudaFuncSetAttribute(layer_tc_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  FPS_CUDA(cudaFuncSetAttribute(layer_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  FPS_CUDA(cudaFuncSetAttribute(layer_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  return 0;
}

void tc_destroy(fps_ctx* ctx) {
  delete static_cast<TcState*>(ctx->tc_state);
  ctx->tc_state = nullptr;
}

int launch_layer_tc(const fps_ctx* ctx, int l, ActPlanes in, ActPlanes out, int64_t points, int streams, bool last,
                    float* H, float* DH, int64_t ld, cudaStream_t st) {
  if (points <= 0) return 0;
  const TcState* s = static_cast<const TcState*>(ctx->tc_state);
  FPS_REQUIRE(s, "tcgen05 engine not initialised");
  const Layer& L = ctx->layer[l];
  const int b = (in.hi == ctx->act[0].hi) ? 0 : 1;
  FPS_REQUIRE(in.hi == ctx->act[b].hi, "layer_tc: input planes must be the context's workspace");
  const int kv = (L.Kp == K0P) ? 0 : 1;
  TcParams P;
  P.bias = L.bias;
  P.out_hi = out.hi;
  P.out_lo = out.lo;
  P.H = H;
  P.DH = DH;
  P.ld = ld;
  P.rows = points * streams;
  P.nk = L.Kp / TK;
  P.row_tiles = (int)((P.rows + TN - 1) / TN);
  const int total = P.row_tiles * N_FEAT_TILES;
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
#define FPS_TC_LAUNCH(S_, LAST_)                                                                              \
  layer_tc_kernel<S_, LAST_><<<grid, TC_THREADS, TC_SMEM, st>>>(s->w_hi[l], s->w_lo[l], s->act_hi[b][kv],      \
                                                                 s->act_lo[b][kv], P)
  if (streams == 4) {
    if (last) FPS_TC_LAUNCH(4, true); else FPS_TC_LAUNCH(4, false);
  } else {
    if (last) FPS_TC_LAUNCH(1, true); else FPS_TC_LAUNCH(1, false);
  }
#undef FPS_TC_LAUNCH
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
