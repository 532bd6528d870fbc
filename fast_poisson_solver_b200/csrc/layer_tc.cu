// Dense layer + tanh with forward-mode derivative streams on the 5th-gen tensor cores (FPS_ENGINE_TCGEN05):
// TMA -> 128B-swizzled shared memory -> tcgen05.mma.kind::f16 (FP16x2 split, three MMAs per product)
// -> fp32 segment accumulators in TMEM -> tcgen05.ld -> round-to-nearest running sums in registers
// -> fused tanh/derivative epilogue -> next layer's hi/lo planes (or H / DH for the last layer).
//
// Reference maths: nn.Linear + nn.Tanh (model.py:53-60, :144) and the Laplacian that
// utils.py:54-68 extracts by 2 800 autograd sweeps, here as closed-form streams.
//
// GEMM mapping (per CTA pair, cta_group::2):
//   D[f, r] = sum_k W[f, k] * Act[r, k]      f: 256 output features  -> UMMA M  (128 TMEM lanes per CTA)
//                                            r: 256 activation rows  -> UMMA N  (TMEM columns)
//   A operand = W tile   (128 x 64 halves per CTA, K-major, SWIZZLE_128B)  hi and lo planes
//   B operand = Act tile (256 x 64 halves, 128 rows staged by each CTA)    hi and lo planes
//   Activation row r = 4*point + stream, so one thread (= one TMEM lane = one feature) finds the
//   four streams (v, v_x, v_y, v_lap) of a point in four adjacent TMEM columns and can apply the
//   tanh recurrences without any cross-thread exchange.
//   Products:  D += Alo*Bhi + Ahi*Blo + Ahi*Bhi per 16-wide k step on fp16 planes of power-of-two scaled
//   values (hi = fp16(v S), lo = fp16(v S - hi); fps_common.cuh).  Template parameter F16 = false selects
//   the earlier 3xTF32 format (fp32 planes, kind::tf32, 8-wide k steps); it also serves the scale calibration,
//   which needs a scale-free format.  RB = bytes of K per shared-memory row (64 or 128; 128 in production).
//
// The pair shares the 256-row activation tile: each CTA loads only 128 of its rows and the tensor cores of
// both SMs read both halves.  That cut the L2->SM operand traffic per MMA by a third, which bounds this kernel.
// NCTA = 1: single-CTA variant kept for A/B measurements (FPS_TC_PAIR=0).
//
// Accumulation accuracy.  The tensor core truncates (rounds toward zero) when it writes the fp32
// accumulator, so a single 704-long chain biases every output by ~3e-6 per layer (measured: H off by 1.6e-5
// after six layers).  The K loop is therefore cut into segments (K = 192 or 256): each segment accumulates in TMEM
// from zero, then the accumulate warps pull it out with tcgen05.ld and add it, times a gain 1 + kappa(K_seg)
// that compensates the first-order shrink, to a round-to-nearest fp32 running sum held in registers while the
// tensor core already works on the next segment in the other TMEM buffer.
//
// Warp roles (384 threads = 3 warpgroups, persistent CTAs, one per SM):
//   warp 0     TMA producer (ring of {Ahi, Alo, Bhi, Blo} slabs: 3 stages x 64 KiB); in the pair leader also the
//              tile scheduler (dynamic: a global ticket counter, see "tile sequence" below)
//   warp 1     TMEM allocator + MMA issuer (one elected lane; in a pair only the rank-0 CTA issues)
//   warps 2-3  idle (they complete the first warpgroup so that setmaxnreg can hand its registers to the others)
//   warps 4-11 accumulate + epilogue; TMEM lane quadrant = warp_id % 4, column half = (warp_id-4)/4,
//              128 running sums per thread.  The 512 TMEM columns hold two 256-column segment
//              accumulators (double buffer between MMA issuer and accumulate warps).
#include "fps_common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

namespace fps {

constexpr int TM = 128;                 // features per CTA tile (UMMA M per CTA)
constexpr int TN = 256;                 // activation rows per tile (UMMA N)
constexpr int BOX_ROWS = 128;           // every TMA box is 128 rows x one swizzle span of K
// The shared-memory row (= swizzle span = K bytes per slab) is a template parameter RB: 64 (SWIZZLE_64B,
// 8 KiB boxes, 6-stage ring) or 128 (SWIZZLE_128B, 16 KiB boxes, full 128-byte L2 lines per TMA row).
constexpr int TC_THREADS = 384;         // three warpgroups: {TMA warp, MMA warp, 2 idle warps}, 2 x 4 accumulate/epilogue warps
// Register budget per warpgroup (setmaxnreg): the kernel launches with 168 registers per thread (65536 / 384); the
// first warpgroup - two single-thread roles - shrinks to 64 and the 13312 registers it gives up let the eight
// accumulate/epilogue warps grow to 216: their 128 running sums plus the epilogue's temporaries no longer spill
// (ptxas: 0 bytes in the hidden-layer instance, 28 in the last-layer one; 40 / 232 made the control roles spill).
constexpr int TC_REGS_CTRL = 64;
constexpr int TC_REGS_EPI = 216;
constexpr int TMEM_COLS = 512;
constexpr int N_FEAT_TILES = FPW / TM;  // 6
constexpr int RING_BYTES = 224 * 1024;
constexpr int TC_SMEM = RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int MAX_STAGES = 8;

// kind::tf32, fp32 accumulate, A and B K-major, M = 128 * NCTA, N = 256
template <int NCTA>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  constexpr uint32_t idesc =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((TM * NCTA) >> 4) << 24);
  if (NCTA == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// kind::f16 (fp16 x fp16 -> fp32 accumulate), K = 16 per instruction: the two small cross terms
template <int NCTA>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((TM * NCTA) >> 4) << 24);
  if (NCTA == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// FPS_TC_STATS=1: in-kernel cycle accounting (developer instrumentation, read with fps_debug_tc_stats)
__device__ unsigned long long g_tc_stats[16];
#define TC_STAT_ADD(i, v) do { if (P.stats && blockIdx.x == 0) atomicAdd(&g_tc_stats[i], (unsigned long long)(v)); } while (0)

struct TcParams {
  const float* bias;
  ActPlanes out;
  float* H;
  float* DH;
  unsigned* dh_cmax;  // per-feature max |DH| (float bits, atomicMax), feeds the exact int8 Gram; may be null
  // exact-engine outputs of the last layer (struct FusedOut): DH snapped to the calibrated grid dh_scale and written as
  // point-major digit planes (gram_i8.cu layout) for the points [point0, ...); dh_flag is raised on overflow
  int8_t* gplanes;
  const float2* dh_scale;
  int* dh_flag;
  int64_t point0;
  int snap_h;         // snap H to the uniform 2^-30 grid
  int64_t ld;
  int64_t rows;   // valid activation rows = points * S
  int nk;         // K slabs of TK
  int tail_steps; // 32-byte MMA steps of the last slab that hold real K (the rest is TMA zero fill: skipped)
  int row_tiles;  // ceil(rows / 256)
  int stats;      // FPS_TC_STATS=1: cycle accounting of CTA 0 into g_tc_stats
  int seg;        // K slabs per tensor-core accumulation segment
  float in_inv[4]; // per stream: 1 / (W16_SCALE * input scale) (FP16x2 format), else 1
  float out_sc[4]; // per stream: scale of the next layer's input planes
  float gain;      // 1 + kappa(K_seg): first-order compensation of the accumulator truncation bias
  float gain_tail; // same for the shorter last segment of a tile
  int stages;      // ring depth actually used (<= RING_BYTES / stage bytes)
  int* sched;      // pair kernel: {next tile, finished pairs} for the dynamic tile scheduler (both zero between launches)
};

// F16 = false: 3xTF32, planes hi32 / lo32, K slab = 16 floats.  F16 = true: FP16x2 split, planes hi16 / lo16,
// K slab = 32 halves.  Either way a slab row is 64 bytes (SWIZZLE_64B), a box is 128 rows = 8 KiB, one MMA
// covers 32 bytes of K, and the stage is {Ahi, Alo, Bhi.., Blo..}; only the tensor maps' element type, the
// MMA kind and the K extent per slab differ.
template <int S, bool LAST, int NCTA, bool F16, int RB>
__device__ __forceinline__ void layer_tc_body(const CUtensorMap& tmWhi, const CUtensorMap& tmWlo,
                                              const CUtensorMap& tmAhi, const CUtensorMap& tmAlo, const TcParams& P) {
  constexpr int ROW_BYTES = RB;
  constexpr int BOX_BYTES = BOX_ROWS * RB;
  constexpr int B_BOXES = TN / BOX_ROWS / NCTA;                   // activation boxes this CTA loads per plane
  constexpr int STAGE_BYTES = 2 * BOX_BYTES + 2 * B_BOXES * BOX_BYTES;  // {Ahi, Alo, Bhi.., Blo..}
  constexpr int TKE = RB / (F16 ? 2 : 4);                         // K elements per slab
  static_assert(RING_BYTES / STAGE_BYTES <= MAX_STAGES, "ring too deep for the barrier block");
  const int STAGES = P.stages;                                    // <= RING_BYTES / STAGE_BYTES

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + RING_BYTES;
  // barrier slots (8 bytes each): full[8], empty[8], tmem_full[2], tmem_empty[2], then the TMEM base word
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 64, bar_tfull = bar_base + 128, bar_tempty = bar_base + 144;
  const uint32_t tmem_slot = bar_base + 160;
  // dynamic tile scheduler (pairs): sched_full[4] | sched_empty[4] (leader's copy is used) | tile ids[4]
  constexpr int SQ = 4;
  const uint32_t sched_full = bar_base + 176, sched_empty = bar_base + 208, sched_tile = bar_base + 240;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int rank = (NCTA == 2) ? (int)cluster_ctarank() : 0;
  const int group = blockIdx.x / NCTA, ngroups = gridDim.x / NCTA;
  constexpr int FT_GROUPS = N_FEAT_TILES / NCTA;  // feature tiles (pairs of tiles) per row tile
  const int total_tiles = P.row_tiles * FT_GROUPS;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWlo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAlo)) : "memory");
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 8 * NCTA);  // one arrive per accumulate warp of every CTA in the pair
    }
    for (int s = 0; s < SQ; ++s) {
      mbar_init(sched_full + 8 * s, 1);
      mbar_init(sched_empty + 8 * s, 18);       // leader: MMA issuer + 8 warps; peer: producer + 8 warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (NCTA == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // ---- tile sequence of this CTA (pair) --------------------------------------------------------------------
  // NCTA == 1: static round robin.  NCTA == 2: DYNAMIC.  SMs differ by ~10 % in operand-feed rate (L2 die
  // locality), so equal static shares left the fast pairs idle for the last ~9 % of every launch.  The leader's
  // producer thread draws tile numbers from a global counter and publishes them through a 4-deep queue in the
  // shared memory of both CTAs; every other role (peer producer, MMA issuer, the 16 accumulate warps) reads the
  // queue in order.  -1 ends the sequence.  The last pair to finish resets the counter for the next launch.
  uint32_t sq_slot = 0, sq_phase = 0;
  int t_static = group - ngroups;
  auto sq_advance = [&]() { if (++sq_slot == SQ) { sq_slot = 0; sq_phase ^= 1; } };
  // leader producer thread only: `ticket` = a number drawn from the global counter (drawn one tile ahead, so that
  // the latency of the atomic hides behind the loads of the current tile)
  auto draw_ticket = [&]() -> int { return (NCTA == 1) ? 0 : atomicAdd(P.sched, 1); };
  auto publish_tile = [&](int ticket) -> int {
    if (NCTA == 1) { t_static += ngroups; return t_static < total_tiles ? t_static : -1; }
    mbar_wait(sched_empty + 8 * sq_slot, sq_phase ^ 1);
    const int t = ticket < total_tiles ? ticket : -1;
#pragma unroll
    for (uint32_t r = 0; r < 2; ++r) st_cluster_u32(mapa_u32(sched_tile + 4 * sq_slot, r), (uint32_t)t);
#pragma unroll
    for (uint32_t r = 0; r < 2; ++r) mbar_arrive_release_cluster(mapa_u32(sched_full + 8 * sq_slot, r));
    sq_advance();
    if (t < 0 && atomicAdd(P.sched + 1, 1) == ngroups - 1) {   // every pair has drawn its last ticket
      P.sched[0] = 0;
      P.sched[1] = 0;
    }
    return t;
  };
  auto next_tile_thread = [&]() -> int {       // a single consumer thread (peer producer, MMA issuer)
    if (NCTA == 1) { t_static += ngroups; return t_static < total_tiles ? t_static : -1; }
    mbar_wait_cluster(sched_full + 8 * sq_slot, sq_phase);
    const int t = (int)ld_shared_u32(sched_tile + 4 * sq_slot);
    if (t >= -1) mbar_arrive_leader(sched_empty + 8 * sq_slot);   // (always true: orders the read before the arrive)
    sq_advance();
    return t;
  };
  auto next_tile_warp = [&]() -> int {         // a whole consumer warp
    if (NCTA == 1) { t_static += ngroups; return t_static < total_tiles ? t_static : -1; }
    mbar_wait_cluster(sched_full + 8 * sq_slot, sq_phase);
    const int t = (int)ld_shared_u32(sched_tile + 4 * sq_slot);
    __syncwarp();
    if (lane == 0 && t >= -1) mbar_arrive_leader(sched_empty + 8 * sq_slot);
    sq_advance();
    return t;
  };

  if (warp < 4) {
  // ---- first warpgroup: the two single-thread roles (every role's code is dominated by its own setmaxnreg, so that
  // ptxas allocates each region with that region's budget)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int t = (rank == 0) ? publish_tile(draw_ticket()) : next_tile_thread();
      while (t >= 0) {
        const int ticket = (rank == 0) ? draw_ticket() : 0;
        const int ft = (t % FT_GROUPS) * NCTA + rank, rt = t / FT_GROUPS;
        const int brow = rt * TN + rank * (TN / NCTA);
        for (int kc = 0; kc < P.nk; ++kc) {
          const long long t_p = clock64();
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          TC_STAT_ADD(7, clock64() - t_p);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          if (rank == 0) mbar_expect_tx(full, STAGE_BYTES * NCTA);
          tma_load_2d<NCTA>(sa, &tmWhi, full, kc * TKE, ft * TM);
          tma_load_2d<NCTA>(sa + BOX_BYTES, &tmWlo, full, kc * TKE, ft * TM);
#pragma unroll
          for (int b = 0; b < B_BOXES; ++b) {
            tma_load_2d<NCTA>(sa + (2 + b) * BOX_BYTES, &tmAhi, full, kc * TKE, brow + b * BOX_ROWS);
            tma_load_2d<NCTA>(sa + (2 + B_BOXES + b) * BOX_BYTES, &tmAlo, full, kc * TKE, brow + b * BOX_ROWS);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        t = (rank == 0) ? publish_tile(ticket) : next_tile_thread();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (pair leader only) =================
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0;
      const long long t_loop = clock64();
      for (;;) {
        const int t = next_tile_thread();
        if (t < 0) break;
        TC_STAT_ADD(6, 1);
        for (int kc0 = 0; kc0 < P.nk; kc0 += P.seg) {
          const int kc1 = (kc0 + P.seg < P.nk) ? kc0 + P.seg : P.nk;
          const long long t_a = clock64();
          mbar_wait(bar_tempty + 8 * abuf, aphase ^ 1);  // accumulate warps have drained this buffer
          TC_STAT_ADD(1, clock64() - t_a);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + abuf * TN;
          for (int kc = kc0; kc < kc1; ++kc) {
            const long long t_b = clock64();
            mbar_wait(bar_full + 8 * stage, phase);
            TC_STAT_ADD(0, clock64() - t_b);
            if (kc == 0) TC_STAT_ADD(8, clock64() - t_b);   // first slab of a tile
            tc_fence_after();
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            const uint64_t a_hi = umma_desc<ROW_BYTES>(sa), a_lo = umma_desc<ROW_BYTES>(sa + BOX_BYTES);
            const uint64_t b_hi = umma_desc<ROW_BYTES>(sa + 2 * BOX_BYTES),
                           b_lo = umma_desc<ROW_BYTES>(sa + (2 + B_BOXES) * BOX_BYTES);
            const int jn = (kc == P.nk - 1) ? P.tail_steps : ROW_BYTES / 32;
#pragma unroll
            for (int j = 0; j < ROW_BYTES / 32; ++j) {
              if (j >= jn) break;
              const uint64_t adv = (uint64_t)((j * 32) >> 4);  // +32 bytes of K per MMA inside the swizzle row
              const uint32_t first = (kc > kc0 || j > 0) ? 1u : 0u;
              if (F16) {
                umma_f16<NCTA>(d_tmem, a_lo + adv, b_hi + adv, first);
                umma_f16<NCTA>(d_tmem, a_hi + adv, b_lo + adv, 1u);
                umma_f16<NCTA>(d_tmem, a_hi + adv, b_hi + adv, 1u);
              } else {
                umma_tf32<NCTA>(d_tmem, a_lo + adv, b_hi + adv, first);
                umma_tf32<NCTA>(d_tmem, a_hi + adv, b_lo + adv, 1u);
                umma_tf32<NCTA>(d_tmem, a_hi + adv, b_hi + adv, 1u);
              }
            }
            umma_commit<NCTA>(bar_empty + 8 * stage);  // smem slot is free once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit<NCTA>(bar_tfull + 8 * abuf);  // segment accumulator complete
          if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
      }
      TC_STAT_ADD(2, clock64() - t_loop);
      if (P.stats) {  // all pair leaders: sum / max / count of the per-launch MMA loop time
        const unsigned long long dt = (unsigned long long)(clock64() - t_loop);
        atomicAdd(&g_tc_stats[9], dt);
        atomicMax(&g_tc_stats[10], dt);
        atomicAdd(&g_tc_stats[11], 1ull);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_EPI));
    // ================= accumulate + epilogue (warps 4..11) =================
    const int quad = warp % 4;          // TMEM lanes [32*quad, 32*quad+32) are the only ones this warp may read
    const int half = (warp - 4) / 4;    // which 128 of the 256 accumulator columns
    uint32_t abuf = 0, aphase = 0;
    for (;;) {
      const int t = next_tile_warp();
      if (t < 0) break;
      const int ft = (t % FT_GROUPS) * NCTA + rank, rt = t / FT_GROUPS;
      const int f = ft * TM + quad * 32 + lane;  // this thread's output feature

      const float bias = P.bias[f];
      float acc[TN / 2];
#pragma unroll
      for (int i = 0; i < TN / 2; ++i) acc[i] = 0.f;
      for (int kc0 = 0; kc0 < P.nk; kc0 += P.seg) {
        const long long t_c = clock64();
        mbar_wait(bar_tfull + 8 * abuf, aphase);
        const long long t_d = clock64();
        if (warp == 4 && lane == 0) TC_STAT_ADD(3, t_d - t_c);
        tc_fence_after();
        const uint32_t taddr = tmem_base + abuf * TN + half * (TN / 2) + ((uint32_t)(quad * 32) << 16);
        const float g = (kc0 + P.seg <= P.nk) ? P.gain : P.gain_tail;
#pragma unroll
        for (int c = 0; c < TN / 64; ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(v[i], g, acc[c * 32 + i]);  // RN promotion + bias gain
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 1) mbar_arrive(bar_tempty + 8 * abuf); else mbar_arrive_leader(bar_tempty + 8 * abuf);
        }
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        if (warp == 4 && lane == 0) TC_STAT_ADD(4, clock64() - t_d);
      }
      const long long t_e = clock64();
      // ---- epilogue: tanh recurrences, split into the next layer's operand planes, 128-byte coalesced stores
      // (a shared-memory staged variant with cp.async.bulk stores was measured SLOWER: 29.5 vs 25.1 ms)
      const int64_t r0 = (int64_t)rt * TN + half * (TN / 2);
      const bool f_ok = f < FP;  // features 704..767 are tile padding
      // rows of a point are valid together (rows = points * S), so one test per point suffices; tiles that lie
      // entirely inside the wave (all but the last row tile) skip the test
      const bool tile_full = r0 + TN / 2 <= P.rows;
      const size_t o0 = (size_t)r0 * FP + f;   // element offset of (row r0, feature f) in the output planes
      if (S == 4 && LAST) {
        // H, DH and (fused with the exact int8 engine, gram_i8.cu) the digit planes of DH.  This thread owns feature f
        // of 32 consecutive points: per digit that is one 32-byte run of planes[slab][digit][f][point % 128].  Lanes with
        // f in [704, 768) are tile padding but still own (all-zero) plane rows.
        const bool planes = P.gplanes != nullptr;
        const float2 sd = (planes && f_ok) ? P.dh_scale[f] : make_float2(0.f, 0.f);
        uint32_t dg[4][8];
        uint32_t zq[4] = {0u, 0u, 0u, 0u};
        float dh_max = 0.f;
        bool ovf = false;
#pragma unroll
        for (int q = 0; q < TN / 8; ++q) {
          float o[4];
          tanh_streams(fmaf(acc[4 * q], P.in_inv[0], bias), acc[4 * q + 1] * P.in_inv[1],
                       acc[4 * q + 2] * P.in_inv[2], acc[4 * q + 3] * P.in_inv[3], o[0], o[1], o[2], o[3]);
          const bool valid = tile_full || r0 + q * 4 < P.rows;
          float hv = o[0], dv = o[3];
          if (P.snap_h) hv = (float)__float2int_rn(hv * 1073741824.0f) * 9.313225746154785e-10f;   // |h| <= 1: exact
          if (planes) {
            const float t = dv * sd.x;                       // sd.x = 2^(30 - e_f): exact
            int X = 0;
            if (fabsf(t) <= 1073741824.0f) {
              X = __float2int_rn(t);
              dv = (float)X * sd.y;                          // exact: X has at most 24 significant bits
            } else if (valid) {
              ovf = true;                                    // beyond the calibrated grid: keep dv, redo the split later
            }
            zq[q & 3] = valid ? balanced_digits4(X) : 0u;
            if ((q & 3) == 3) {
              uint32_t w4[4];
              digits_transpose4(zq[0], zq[1], zq[2], zq[3], w4);
#pragma unroll
              for (int s4 = 0; s4 < 4; ++s4) dg[s4][q >> 2] = w4[s4];
            }
          }
          if (valid && f_ok) {
            const int64_t p = (r0 >> 2) + q;
            P.H[p * P.ld + f] = hv;
            P.DH[p * P.ld + f] = dv;
            dh_max = fmaxf(dh_max, fabsf(o[3]));
          }
        }
        if (f_ok && P.dh_cmax) atomicMax(P.dh_cmax + f, __float_as_uint(dh_max));
        if (ovf) *P.dh_flag = 1;
        if (planes) {
          const int64_t pg = P.point0 + (r0 >> 2);           // first of this thread's 32 points (a multiple of 32)
          int8_t* base = P.gplanes + (((pg >> 7) * 4) * FPW + f) * 128 + (pg & 127);
          // one 32-byte store per digit (STG.256): whole sectors - two 16-byte stores per run cost twice the sector
          // transactions and put the last layer 1.5 ms behind the hidden ones
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(base + (int64_t)s4 * FPW * 128),
                         "r"(dg[s4][0]), "r"(dg[s4][1]), "r"(dg[s4][2]), "r"(dg[s4][3]), "r"(dg[s4][4]), "r"(dg[s4][5]),
                         "r"(dg[s4][6]), "r"(dg[s4][7])
                         : "memory");
        }
      } else if (f_ok) {
        if (S == 4) {
#pragma unroll
          for (int q = 0; q < TN / 8; ++q) {
            float o[4];
            tanh_streams(fmaf(acc[4 * q], P.in_inv[0], bias), acc[4 * q + 1] * P.in_inv[1],
                         acc[4 * q + 2] * P.in_inv[2], acc[4 * q + 3] * P.in_inv[3], o[0], o[1], o[2], o[3]);
            if (tile_full || r0 + q * 4 < P.rows) {
#pragma unroll
              for (int s = 0; s < 4; ++s) store_planes_t<F16>(P.out, o0 + (size_t)(q * 4 + s) * FP, o[s], P.out_sc[s]);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < TN / 2; ++q) {
            const float h = tanhf(fmaf(acc[q], P.in_inv[0], bias));
            if (tile_full || r0 + q < P.rows) {
              if (LAST) P.H[(r0 + q) * P.ld + f] = h;
              else store_planes_t<F16>(P.out, o0 + (size_t)q * FP, h, P.out_sc[0]);
            }
          }
        }
      }
      if (warp == 4 && lane == 0) TC_STAT_ADD(5, clock64() - t_e);
    }
  }

  __syncwarp();
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int S, bool LAST, bool F16, int RB>
__global__ void __launch_bounds__(TC_THREADS, 1)
layer_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                const TcParams P) {
  layer_tc_body<S, LAST, 1, F16, RB>(tmWhi, tmWlo, tmAhi, tmAlo, P);
}

template <int S, bool LAST, bool F16, int RB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
layer_tc_pair_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                     const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                     const TcParams P) {
  layer_tc_body<S, LAST, 2, F16, RB>(tmWhi, tmWlo, tmAhi, tmAlo, P);
}

// ---- host side: tensor maps -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  CUtensorMap w_hi[NLAYER], w_lo[NLAYER], w_hi16[NLAYER], w_lo16[NLAYER];
  CUtensorMap act_hi[2][2], act_lo[2][2], act_hi16[2][2], act_lo16[2][2];  // [buffer][0: K = K0P, 1: K = FP]
  // 128-byte-row (SWIZZLE_128B) maps of the fp16 planes
  CUtensorMap w_hi16w[NLAYER], w_lo16w[NLAYER], act_hi16w[2][2], act_lo16w[2][2];
  int pair;
  int rb;  // shared-memory row bytes of the production (pair, FP16x2) kernel: 64 or 128
};

// 2-D K-major map: dims {K, rows}, box {rb bytes of K, 128 rows}, swizzle span = rb (64: 16 floats or 32 halves
// per row; 128: 64 halves per row).  A box that overhangs K (K0P = 416 with 64-wide slabs) is zero-filled.
static int make_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, uint64_t K, uint64_t pitch, uint64_t rows,
                    bool half, int rb = 64) {
  const uint64_t esz = half ? 2 : 4;
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {pitch * esz};   // pitch >= K: columns K.. of a row are never fetched (zero fill)
  cuuint32_t box[2] = {(cuuint32_t)(rb / esz), (cuuint32_t)BOX_ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for K=%llu rows=%llu half=%d", (int)r, (unsigned long long)K,
              (unsigned long long)rows, (int)half);
    return 1;
  }
  return 0;
}

template <typename Kern>
static int set_smem(Kern k) {
  FPS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  return 0;
}

template <bool F16>
static int set_smem_all() {
  return set_smem(layer_tc_kernel<4, false, F16, 64>) || set_smem(layer_tc_kernel<4, true, F16, 64>) ||
         set_smem(layer_tc_kernel<1, false, F16, 64>) || set_smem(layer_tc_kernel<1, true, F16, 64>) ||
         set_smem(layer_tc_pair_kernel<4, false, F16, 64>) || set_smem(layer_tc_pair_kernel<4, true, F16, 64>) ||
         set_smem(layer_tc_pair_kernel<1, false, F16, 64>) || set_smem(layer_tc_pair_kernel<1, true, F16, 64>);
}
static int set_smem_wide() {
  return set_smem(layer_tc_pair_kernel<4, false, true, 128>) || set_smem(layer_tc_pair_kernel<4, true, true, 128>) ||
         set_smem(layer_tc_pair_kernel<1, false, true, 128>) || set_smem(layer_tc_pair_kernel<1, true, true, 128>);
}

int tc_init(fps_ctx* ctx) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FPS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  FPS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  TcState* s = new TcState();
  ctx->tc_state = s;
  s->pair = getenv("FPS_TC_PAIR") ? atoi(getenv("FPS_TC_PAIR")) : 1;
  s->rb = getenv("FPS_TC_ROWB") ? atoi(getenv("FPS_TC_ROWB")) : 128;
  FPS_REQUIRE(s->rb == 64 || s->rb == 128, "FPS_TC_ROWB must be 64 or 128");
  for (int l = 0; l < NLAYER; ++l) {
    const Layer& L = ctx->layer[l];
    if (make_map(enc, &s->w_hi[l], L.w_hi, L.Kp, L.ldk, FPW, false)) return 1;
    if (make_map(enc, &s->w_lo[l], L.w_lo, L.Kp, L.ldk, FPW, false)) return 1;
    if (make_map(enc, &s->w_hi16[l], L.w_hi16, L.Kp, L.ldk, FPW, true)) return 1;
    if (make_map(enc, &s->w_lo16[l], L.w_lo16, L.Kp, L.ldk, FPW, true)) return 1;
    if (make_map(enc, &s->w_hi16w[l], L.w_hi16, L.Kp, L.ldk, FPW, true, 128)) return 1;
    if (make_map(enc, &s->w_lo16w[l], L.w_lo16, L.Kp, L.ldk, FPW, true, 128)) return 1;
  }
  const uint64_t cap_rows = (uint64_t)ctx->chunk * NSTREAM;
  for (int b = 0; b < 2; ++b) {
    // a plane holds cap_rows x FP elements; viewed with K = K0P it has room for more rows, but the same
    // row capacity is all a wave ever uses
    const uint64_t Ks[2] = {(uint64_t)K0P, (uint64_t)FP};
    const uint64_t Ls[2] = {(uint64_t)ctx->layer[0].ldk, (uint64_t)FP};
    for (int kv = 0; kv < 2; ++kv) {
      if (make_map(enc, &s->act_hi[b][kv], ctx->act[b].hi, Ks[kv], Ls[kv], cap_rows, false)) return 1;
      if (make_map(enc, &s->act_lo[b][kv], ctx->act[b].lo, Ks[kv], Ls[kv], cap_rows, false)) return 1;
      if (make_map(enc, &s->act_hi16[b][kv], ctx->act[b].hi16, Ks[kv], Ls[kv], cap_rows, true)) return 1;
      if (make_map(enc, &s->act_lo16[b][kv], ctx->act[b].lo16, Ks[kv], Ls[kv], cap_rows, true)) return 1;
      if (make_map(enc, &s->act_hi16w[b][kv], ctx->act[b].hi16, Ks[kv], Ls[kv], cap_rows, true, 128)) return 1;
      if (make_map(enc, &s->act_lo16w[b][kv], ctx->act[b].lo16, Ks[kv], Ls[kv], cap_rows, true, 128)) return 1;
    }
  }
  if (set_smem_all<false>() || set_smem_all<true>() || set_smem_wide()) return 1;
  return 0;
}

void tc_destroy(fps_ctx* ctx) {
  delete static_cast<TcState*>(ctx->tc_state);
  ctx->tc_state = nullptr;
}

// Truncation-bias compensation (see header): the tensor core rounds the fp32 accumulator toward zero after
// every MMA, which shrinks a segment sum by kappa on average.  Measured on B200 against the fp64 oracle
// (tools/calibrate_kappa.py) for the 3xTF32 scheme (3 MMAs per 8-wide k step): 4.2 ulp at K=32, 6 ulp at
// K=64, 9 ulp at K=128, i.e. kappa(K) = 4.2 * 2^-23 * (K/32)^0.55.  The TF32 + 2xFP16 scheme issues 2 MMAs
// per 8 k, so the same law is evaluated at 2/3 of the MMA count.  FPS_TC_KAPPA overrides (calibration).
static float kappa_of(int kseg, bool f16) {
  if (f16) {
    // FP16x2 split, calibrated on B200 against the fp64 oracle (tools/calibrate_kappa.py): optimal gain per
    // segment length K; piecewise linear in between, proportional below K = 128.
    static const float K[5] = {128.f, 192.f, 256.f, 352.f, 704.f};
    static const float V[5] = {3.0e-7f, 5.2e-7f, 7.6e-7f, 1.03e-6f, 2.05e-6f};
    const float k = (float)kseg;
    if (k <= K[0]) return V[0] * k / K[0];
    for (int i = 1; i < 5; ++i)
      if (k <= K[i]) return V[i - 1] + (V[i] - V[i - 1]) * (k - K[i - 1]) / (K[i] - K[i - 1]);
    return V[4] * k / K[4];
  }
  // 3xTF32: 4.2 ulp at K = 32, 6 ulp at K = 64, 9 ulp at K = 128
  return 4.2f * 1.1920929e-7f * powf((float)kseg / 32.f, 0.55f);
}

int launch_layer_tc(const fps_ctx* ctx, int l, ActPlanes in, ActPlanes out, int64_t points, int streams, bool last,
                    float* H, float* DH, int64_t ld, const FusedOut* fused, cudaStream_t st) {
  if (points <= 0) return 0;
  const TcState* s = static_cast<const TcState*>(ctx->tc_state);
  FPS_REQUIRE(s, "tcgen05 engine not initialised");
  const Layer& L = ctx->layer[l];
  const int b = (in.hi == ctx->act[0].hi) ? 0 : 1;
  FPS_REQUIRE(in.hi == ctx->act[b].hi, "layer_tc: input planes must be the context's workspace");
  const int kv = (L.Kp == K0P) ? 0 : 1;
  const bool f16 = ctx->f16 != 0;
  TcParams P;
  P.bias = L.bias;
  P.out = out;
  P.H = H;
  P.DH = DH;
  const bool fz = last && DH && fused && fused->xdh;
  P.dh_cmax = fz ? fused->xdh->cmax : nullptr;
  P.gplanes = fz ? fused->gplanes : nullptr;
  P.dh_scale = fz ? fused->xdh->scale : nullptr;
  P.dh_flag = fz ? fused->xdh->flags : nullptr;
  P.point0 = fz ? fused->point0 : 0;
  P.snap_h = (last && DH && fused) ? fused->snap_h : 0;
  FPS_REQUIRE(!P.gplanes || (P.point0 % 128 == 0 && (uintptr_t)P.gplanes % 16 == 0),
              "layer_tc: fused digit planes need a 128-point aligned wave and a 16-byte aligned buffer");
  P.sched = ctx->tc_sched;
  P.ld = ld;
  P.rows = points * streams;
  const bool wide = f16 && s->pair && s->rb == 128;     // 128-byte slab rows (SWIZZLE_128B)
  const int slab_k = (wide ? 128 : 64) / (f16 ? 2 : 4);   // K elements per slab
  P.nk = (L.Kp + slab_k - 1) / slab_k;                    // an overhanging last slab is zero-filled by TMA
  {
    const int step_k = f16 ? 16 : 8;                      // K elements per MMA instruction
    const int rem = L.Kp - (P.nk - 1) * slab_k;           // real K in the last slab
    P.tail_steps = (rem + step_k - 1) / step_k;
  }
  for (int i = 0; i < 4; ++i) {
    P.in_inv[i] = f16 ? 1.0f / (W16_SCALE * ctx->act_scale[l].s[i]) : 1.0f;
    P.out_sc[i] = ctx->act_scale[l + 1 < NLAYER ? l + 1 : l].s[i];
  }
  P.row_tiles = (int)((P.rows + TN - 1) / TN);
  static const int stats = getenv("FPS_TC_STATS") ? atoi(getenv("FPS_TC_STATS")) : 0;
  P.stats = stats;
  // K slabs per tensor-core accumulation segment.  3xTF32: 128.  FP16x2 (64-wide slabs): hidden layers 3 = segments of
  // 192, 192, 192, 128 per K = 704; first and last layer 4 = 256, 160 resp. 256, 256, 192.  The MMA issuer may run two
  // segments ahead of the accumulate warps (two TMEM buffers), and those warps spend ~8.2 K cycles per tile in the
  // epilogue before they drain the next tile's first segment: with 192-wide segments the look-ahead (9.2 K cycles of
  // MMAs) falls short of epilogue + drains in the last layer (fused digit split) and the first one (K = 416), and the
  // tensor pipe idles; 256-wide segments cover it (measured, interleaved runs under the power cap: last layer 1.084 ->
  // 0.972 of a hidden layer's time, first layer 0.185 -> 0.174 of the four hidden ones).  The hidden layers gain nothing
  // from longer segments and keep 192, because the truncation bias of a segment grows with its length: H / DH vs fp64
  // 6.98e-7 / 1.09e-6 (all 192), 7.44e-7 / 1.12e-6 (this mix; the fp32 FFMA engine: 7.3e-7 / 1.1e-6), 8.16e-7 / 1.22e-6
  // (all 256).  FPS_TC_SEG (all layers) / FPS_TC_SEG_L0 / _HID / _LAST override (measurements).
  static const int seg_env = getenv("FPS_TC_SEG") ? atoi(getenv("FPS_TC_SEG")) : 0;
  static const int seg_env_l0 = getenv("FPS_TC_SEG_L0") ? atoi(getenv("FPS_TC_SEG_L0")) : 0;
  static const int seg_env_hid = getenv("FPS_TC_SEG_HID") ? atoi(getenv("FPS_TC_SEG_HID")) : 0;
  static const int seg_env_last = getenv("FPS_TC_SEG_LAST") ? atoi(getenv("FPS_TC_SEG_LAST")) : 0;
  const int seg_layer = (l == 0) ? seg_env_l0 : (last ? seg_env_last : seg_env_hid);
  const int seg_default = f16 ? ((l == 0 || last) ? 256 : 192) / slab_k : 128 / slab_k;
  const int seg = seg_layer ? seg_layer : seg_env ? seg_env : seg_default;
  P.seg = seg;
  static const char* kenv = getenv("FPS_TC_KAPPA");
  const int tail = (P.nk % seg) ? (P.nk % seg) : seg;
  P.gain = 1.0f + (kenv ? (float)atof(kenv) : kappa_of(seg * slab_k, f16));
  const int tail_k = L.Kp - (P.nk - tail) * slab_k;   // real K of the last segment (the zero-filled MMA steps are skipped)
  P.gain_tail = 1.0f + (kenv ? (float)atof(kenv) * powf((float)tail / seg, 0.55f) : kappa_of(tail_k, f16));
  if (ctx->gain_off) P.gain = P.gain_tail = 1.0f;   // fps_calibrate_scales found the compensation harmful for these weights
  const int ncta = s->pair ? 2 : 1;
  {
    const int stage_bytes = (wide ? 128 : 64) * BOX_ROWS * (2 + 2 * (TN / BOX_ROWS / ncta));
    static const int stages_env = getenv("FPS_TC_STAGES") ? atoi(getenv("FPS_TC_STAGES")) : 0;
    const int max_stages = RING_BYTES / stage_bytes;
    P.stages = (stages_env > 0 && stages_env < max_stages) ? stages_env : max_stages;
  }
  const int total = P.row_tiles * (N_FEAT_TILES / ncta);          // tiles of a CTA (pair)
  const int groups_max = ctx->sm_count / ncta;
  const int grid = (total < groups_max ? total : groups_max) * ncta;
  const CUtensorMap& m1 = wide ? s->w_hi16w[l] : f16 ? s->w_hi16[l] : s->w_hi[l];
  const CUtensorMap& m2 = wide ? s->w_lo16w[l] : f16 ? s->w_lo16[l] : s->w_lo[l];
  const CUtensorMap& m3 = wide ? s->act_hi16w[b][kv] : f16 ? s->act_hi16[b][kv] : s->act_hi[b][kv];
  const CUtensorMap& m4 = wide ? s->act_lo16w[b][kv] : f16 ? s->act_lo16[b][kv] : s->act_lo[b][kv];
#define FPS_TC_LAUNCH(K_, S_, LAST_, F_, RB_) \
  K_<S_, LAST_, F_, RB_><<<grid, TC_THREADS, TC_SMEM, st>>>(m1, m2, m3, m4, P)
#define FPS_TC_PICK(K_, F_, RB_)                                                                     \
  do {                                                                                               \
    if (streams == 4) {                                                                              \
      if (last) FPS_TC_LAUNCH(K_, 4, true, F_, RB_); else FPS_TC_LAUNCH(K_, 4, false, F_, RB_);      \
    } else {                                                                                         \
      if (last) FPS_TC_LAUNCH(K_, 1, true, F_, RB_); else FPS_TC_LAUNCH(K_, 1, false, F_, RB_);      \
    }                                                                                                \
  } while (0)
  if (wide) {
    FPS_TC_PICK(layer_tc_pair_kernel, true, 128);
  } else if (s->pair) {
    if (f16) FPS_TC_PICK(layer_tc_pair_kernel, true, 64); else FPS_TC_PICK(layer_tc_pair_kernel, false, 64);
  } else {
    if (f16) FPS_TC_PICK(layer_tc_kernel, true, 64); else FPS_TC_PICK(layer_tc_kernel, false, 64);
  }
#undef FPS_TC_PICK
#undef FPS_TC_LAUNCH
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps

extern "C" int fps_debug_tc_stats(unsigned long long* out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, fps::g_tc_stats, sizeof(fps::g_tc_stats)) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(fps::g_tc_stats, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
