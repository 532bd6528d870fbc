// Single / few right-hand-side streaming kernels of solve() (solver.py:301 `RHSs @ f`, :327-330 `H @ w`, `DH @ w`): pure
// HBM-bound passes over the fp32 feature matrices (2 816 B per point and matrix).  The feature matrices are dense
// row-major with ld = 704, i.e. ONE contiguous array, so each persistent CTA (one per SM) streams its share of the rows
// through a 4-stage shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) issued by a
// producer thread, and eight consumer warps do the fp64 arithmetic out of shared memory.  ~45 KB per stage: 180 KB in
// flight per SM, independent of the consumers' register pressure (the register-staged predecessors, gemm_f64.cu
// atb_small_vec_kernel / nn_small_kernel, stayed at 60-70 % of the measured copy bandwidth).
//   proj_stream_kernel   partial[cta][f][c] = sum_{rows of the CTA} A[r][f] F[r][c]      (then atb_small_reduce)
//   recon_stream_kernel  out[r][c] = sum_f A[r][f] W[f][c] + bias[c]
#include "fps_common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <stdlib.h>

namespace fps {

constexpr int ST_ROWS = 16;                          // rows per chunk
constexpr int ST_CHUNK = ST_ROWS * FP * 4;           // 45 056 bytes
constexpr int ST_STAGES = 4;
constexpr int ST_THREADS = 288;                      // reconstruction: warp 0 producer, warps 1..8 consumers
constexpr int ST_THREADS_PROJ = 224;                 // projection: warp 0 producer, warps 1..6 consumers (176 column owners)
constexpr int ST_SMEM = ST_STAGES * ST_CHUNK + 1024 + 128;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct StreamSetup {
  uint32_t ring, bar_full, bar_empty;
};

__device__ __forceinline__ StreamSetup stream_setup(uint8_t* smem_raw, int consumer_warps) {
  StreamSetup s;
  s.ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  s.bar_full = s.ring + ST_STAGES * ST_CHUNK;
  s.bar_empty = s.bar_full + 8 * ST_STAGES;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST_STAGES; ++i) {
      mbar_init(s.bar_full + 8 * i, 1);
      mbar_init(s.bar_empty + 8 * i, consumer_warps);  // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  return s;
}

// producer loop: rows [r0, r1) of A in chunks of ST_ROWS
__device__ __forceinline__ void stream_produce(const StreamSetup& s, const float* A, int64_t r0, int64_t r1) {
  uint32_t stage = 0, phase = 0;
  for (int64_t r = r0; r < r1; r += ST_ROWS) {
    const uint32_t bytes = (uint32_t)(min((int64_t)ST_ROWS, r1 - r) * FP * 4);
    mbar_wait(s.bar_empty + 8 * stage, phase ^ 1);
    mbar_expect_tx(s.bar_full + 8 * stage, bytes);
    bulk_load_1d(s.ring + stage * ST_CHUNK, A + r * FP, bytes, s.bar_full + 8 * stage);
    if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
  }
}

// ---- projection: every consumer THREAD owns four consecutive feature columns (704 = 176 x 4: six warps, 176 of 192
// threads) and walks all rows of the chunk, one 16-byte word per row from shared memory.  The kernel sits on the fp64
// pipe (fp32 -> fp64 conversions + DFMA: ~2 clocks per warp instruction) as much as on HBM, so idle lanes cost: eight
// warps with 176 owners ran at 75 % of the copy bandwidth.  The right-hand-side value of a row is converted once and
// broadcast as a double.
template <int NB>
__global__ void __launch_bounds__(ST_THREADS_PROJ, 1)
proj_stream_kernel(const float* __restrict__ A, const float* __restrict__ Fm, int64_t ldf, int nb, int64_t n,
                   int64_t rows_per_cta, double* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  const StreamSetup s = stream_setup(smem_raw, ST_THREADS_PROJ / 32 - 1);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0) {
    if (lane == 0 && r0 < r1) stream_produce(s, A, r0, r1);
    return;
  }
  const int ct = threadIdx.x - 32;                   // consumer thread 0..191
  const bool owner = ct < FP / 4;
  double acc[4][NB];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < NB; ++c) acc[q][c] = 0.0;
  uint32_t stage = 0, phase = 0;
  for (int64_t r = r0; r < r1; r += ST_ROWS) {
    const int rows = (int)min((int64_t)ST_ROWS, r1 - r);
    // this chunk's right-hand-side values: lane l of every warp fetches row l, broadcast by shuffle below
    double fv[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) fv[c] = (lane < rows && c < nb) ? (double)__ldg(Fm + (r + lane) * ldf + c) : 0.0;
    mbar_wait(s.bar_full + 8 * stage, phase);
    const float4* src = reinterpret_cast<const float4*>(smem_raw + (s.ring - smem_u32(smem_raw)) + stage * ST_CHUNK) + ct;
#pragma unroll
    for (int k = 0; k < ST_ROWS; ++k) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (owner && k < rows) a = src[k * (FP / 4)];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const double f = __shfl_sync(0xffffffffu, fv[c], k);
        acc[0][c] = fma((double)a.x, f, acc[0][c]);
        acc[1][c] = fma((double)a.y, f, acc[1][c]);
        acc[2][c] = fma((double)a.z, f, acc[2][c]);
        acc[3][c] = fma((double)a.w, f, acc[3][c]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_release_cta(s.bar_empty + 8 * stage);
    if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
  }
  if (owner) {
    double* out = partial + (size_t)blockIdx.x * (768 * NB);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int c = 0; c < NB; ++c) out[(4 * ct + q) * NB + c] = acc[q][c];
  }
}

// ---- reconstruction: consumer warp w takes rows w-1, w-1+8 of every chunk; lane l keeps the weights of its feature
// words (float4 word index l + 32 q, q < 6: 176 words) in registers, fp64 accumulate, warp-shuffle reduction per row.
template <int NB>
__global__ void __launch_bounds__(ST_THREADS, 1)
recon_stream_kernel(const float* __restrict__ A, int64_t n, int64_t rows_per_cta, const double* __restrict__ W, int64_t ldw,
                    const double* __restrict__ bias, int nb, float* __restrict__ out, int64_t ldo) {
  extern __shared__ uint8_t smem_raw[];
  const StreamSetup s = stream_setup(smem_raw, ST_THREADS / 32 - 1);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0) {
    if (lane == 0 && r0 < r1) stream_produce(s, A, r0, r1);
    return;
  }
  double w[6][4][NB];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int word = lane + 32 * q;
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int c = 0; c < NB; ++c)
        w[q][e][c] = (word < FP / 4 && c < nb) ? __ldg(W + (int64_t)(4 * word + e) * ldw + c) : 0.0;
  }
  double bj[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) bj[c] = (bias && c < nb) ? bias[c] : 0.0;
  uint32_t stage = 0, phase = 0;
  for (int64_t r = r0; r < r1; r += ST_ROWS) {
    const int rows = (int)min((int64_t)ST_ROWS, r1 - r);
    mbar_wait(s.bar_full + 8 * stage, phase);
    const float4* base = reinterpret_cast<const float4*>(smem_raw + (s.ring - smem_u32(smem_raw)) + stage * ST_CHUNK);
#pragma unroll
    for (int h = 0; h < ST_ROWS / 8; ++h) {
      const int k = warp - 1 + 8 * h;
      double acc[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] = 0.0;
      if (k < rows) {
        const float4* row = base + k * (FP / 4);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int word = lane + 32 * q;
          if (word < FP / 4) {
            const float4 a = row[word];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
              acc[c] = fma((double)a.x, w[q][0][c], acc[c]);
              acc[c] = fma((double)a.y, w[q][1][c], acc[c]);
              acc[c] = fma((double)a.z, w[q][2][c], acc[c]);
              acc[c] = fma((double)a.w, w[q][3][c], acc[c]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NB; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
      if (k < rows && lane < nb) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) v = (lane == c) ? acc[c] + bj[c] : v;
        out[(r + k) * ldo + lane] = (float)v;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_release_cta(s.bar_empty + 8 * stage);
    if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
  }
}

static bool stream_ok(const void* A, int64_t ld) { return ld == FP && (uintptr_t)A % 16 == 0; }

template <typename K>
static int stream_attr(K kern) {
  FPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
  return 0;
}

int stream_init(fps_ctx* ctx) {
  (void)ctx;
  return stream_attr(proj_stream_kernel<1>) || stream_attr(proj_stream_kernel<2>) || stream_attr(proj_stream_kernel<4>) ||
         stream_attr(proj_stream_kernel<8>) || stream_attr(recon_stream_kernel<1>) || stream_attr(recon_stream_kernel<2>);
}

static void stream_split(const fps_ctx* ctx, int64_t n, int* ctas, int64_t* rows_per_cta) {
  int64_t c = std::min<int64_t>(ctx->sm_count, (n + ST_ROWS - 1) / ST_ROWS);
  const int64_t rpc = ((n + c - 1) / c + ST_ROWS - 1) / ST_ROWS * ST_ROWS;   // chunk-aligned shares: 16-byte aligned sources
  *ctas = (int)((n + rpc - 1) / rpc);
  *rows_per_cta = rpc;
}

// returns 1 when the streaming form does not apply (caller falls back), 0 on success, > 1 on error
int launch_project_stream(const fps_ctx* ctx, const float* A, int64_t lda, const float* Fm, int64_t ldf, int B, int64_t n,
                          double* partial, int* nctas_out, cudaStream_t st) {
  static const int on = getenv("FPS_STREAM") ? atoi(getenv("FPS_STREAM")) : 1;
  if (!on || !stream_ok(A, lda) || B > 8 || n < 4096) return 1;
  int ctas;
  int64_t rpc;
  stream_split(ctx, n, &ctas, &rpc);
  if ((size_t)ctas * 768 * 8 * sizeof(double) > ctx->partial_bytes) return 1;
#define FPS_PS(NB_) proj_stream_kernel<NB_><<<ctas, ST_THREADS_PROJ, ST_SMEM, st>>>(A, Fm, ldf, B, n, rpc, partial)
  if (B == 1) FPS_PS(1); else if (B == 2) FPS_PS(2); else if (B <= 4) FPS_PS(4); else FPS_PS(8);
#undef FPS_PS
  count_launch();
  if (cudaGetLastError() != cudaSuccess) { set_error("proj_stream launch failed"); return 2; }
  *nctas_out = ctas;
  return 0;
}

int launch_recon_stream(const fps_ctx* ctx, const float* A, int64_t n, int64_t lda, const double* W, int64_t ldw,
                        const double* bias, int B, float* out, int64_t ldo, cudaStream_t st) {
  static const int on = getenv("FPS_STREAM") ? atoi(getenv("FPS_STREAM")) : 1;
  if (!on || !stream_ok(A, lda) || B > 2 || n < 4096) return 1;   // (the weights of a lane's 24 features live in registers)
  int ctas;
  int64_t rpc;
  stream_split(ctx, n, &ctas, &rpc);
#define FPS_RS(NB_) recon_stream_kernel<NB_><<<ctas, ST_THREADS, ST_SMEM, st>>>(A, n, rpc, W, ldw, bias, B, out, ldo)
  if (B == 1) FPS_RS(1); else FPS_RS(2);
#undef FPS_RS
  count_launch();
  if (cudaGetLastError() != cudaSuccess) { set_error("recon_stream launch failed"); return 2; }
  return 0;
}

}  // namespace fps
