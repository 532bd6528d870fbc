// FP64-accumulated contractions over the feature matrices (DMMA m8n8k4 tensor path).
//
//   atb      C(M x Nc) += A^T B   over n rows      Gram  DH^T DH        solver.py:167
//                                                  BC    H_bc^T H_bc    solver.py:196
//                                                  RHS   DH^T F         solver.py:197 + :301
//   nn       out(n x B) = A W + bias               u = H w + b, f = DH w, u_bc = H_bc w + b
//                                                                        solver.py:327-330
// Why fp64: the normal matrix has cond ~1e15; accumulating the Gram or the RHS
// projection at fp32 grade moves u by 1e-4..3e-2, far beyond the 1e-5 parity bar.  In the default
// (precision=float32) mode the inputs stay fp32 in HBM (that is what the feature kernel produces) and
// are widened on the way into shared memory, so HBM traffic is the fp32 footprint; in
// precision=float64 mode the same kernels read fp64 features (template parameter).
//
// Split-K (over rows) partial tiles are written to a workspace and summed in a fixed order by a
// second kernel, so results are bit-reproducible for a given (n, chunking) - no fp64 atomics.
#include "fps_common.cuh"
#include <algorithm>
#include <type_traits>

namespace fps {

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// sm_90+ shape: 16x8x16 per instruction (8x fewer issues than m8n8k4).  Fragments (CUTLASS
// SM90_16x8x16_F64F64F64F64_TN traits; g = lane / 4, t = lane % 4):
//   a[v0 + 2 v1] = A[g + 8 v0][t + 4 v1]   (v0 < 2, v1 < 4)      b[v] = B[t + 4 v][g]   (v < 4)
//   c[v0 + 2 v1] = C[g + 8 v1][2 t + v0]   (v0, v1 < 2)
__device__ __forceinline__ void dmma16x8x16(double* c, const double* a, const double* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}

constexpr int GT = 64;         // output tile edge
constexpr int GK = 32;         // rows (contraction) per step
constexpr int GS = GT + 4;     // smem row stride in doubles: == 4 (mod 16) -> conflict-free fragment loads
constexpr int GTHREADS = 128;  // 4 warps, each a 32x32 sub-tile

// Four consecutive elements of a row, widened to double; zero outside [ncols).
template <typename T>
__device__ __forceinline__ void load4(const T* __restrict__ p, int c, int ncols, bool vec_ok, double* o);

template <>
__device__ __forceinline__ void load4<float>(const float* __restrict__ p, int c, int ncols, bool vec_ok, double* o) {
  if (vec_ok && c + 3 < ncols) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = (c + e < ncols) ? (double)__ldg(p + e) : 0.0;
  }
}
template <>
__device__ __forceinline__ void load4<double>(const double* __restrict__ p, int c, int ncols, bool vec_ok, double* o) {
  if (vec_ok && c + 3 < ncols) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = (c + e < ncols) ? __ldg(p + e) : 0.0;
  }
}

// Register image of a (GK x GT) tile whose rows are contraction indices: 16 elements per thread.
// fp32 inputs are kept as fp32 in registers (half the footprint) and widened when stored to smem.
template <typename T>
struct TileRegs {
  T v[16];
};

template <typename T>
__device__ __forceinline__ void load_tile_kn(TileRegs<T>& R, const T* __restrict__ P, int64_t ld, int64_t k0,
                                             int64_t kend, int c0, int ncols, bool vec_ok) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = q * GTHREADS + threadIdx.x;
    const int r = idx / 16, c = (idx % 16) * 4;
    double o[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t k = k0 + r;
    if (k < kend) load4<T>(P + k * ld + c0 + c, c0 + c, ncols, vec_ok, o);
#pragma unroll
    for (int e = 0; e < 4; ++e) R.v[q * 4 + e] = (T)o[e];
  }
}

template <typename T>
__device__ __forceinline__ void store_tile_kn(const TileRegs<T>& R, double* S) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = q * GTHREADS + threadIdx.x;
    const int r = idx / 16, c = (idx % 16) * 4;
    double2* d = reinterpret_cast<double2*>(S + r * GS + c);
    d[0] = make_double2((double)R.v[q * 4 + 0], (double)R.v[q * 4 + 1]);
    d[1] = make_double2((double)R.v[q * 4 + 2], (double)R.v[q * 4 + 3]);
  }
}

__device__ __forceinline__ void tri_tile(int t, int& ti, int& tj) {  // t enumerates the lower triangle, ti >= tj
  ti = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  tj = t - ti * (ti + 1) / 2;
}

// ------------------------------------------------------------------------------------------------
// C_partial[slice][tile] = A^T B over this slice's rows
// ------------------------------------------------------------------------------------------------
template <typename TA, typename TB>
__global__ void __launch_bounds__(GTHREADS)
atb_f64_kernel(const TA* __restrict__ A, int64_t lda, int M, const TB* __restrict__ B, int64_t ldb, int Nc,
               int64_t n, int64_t rows_per_slice, int tiles_j, int symmetric, double* __restrict__ partial,
               int vecA, int vecB) {
  __shared__ __align__(16) double As[GK * GS];
  __shared__ __align__(16) double Bs[GK * GS];
  int ti, tj;
  if (symmetric) {
    tri_tile(blockIdx.x, ti, tj);
  } else {
    ti = blockIdx.x / tiles_j;
    tj = blockIdx.x % tiles_j;
  }
  const int64_t kbeg = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t kend = min(n, kbeg + rows_per_slice);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int wi = warp / 2, wj = warp % 2;
  const int lr = lane / 4, lc = lane % 4;

  double c[2][4][4];  // [m16 tile][n8 tile][fragment]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) c[i][j][v] = 0.0;

  TileRegs<TA> ra;
  TileRegs<TB> rb;
  load_tile_kn<TA>(ra, A, lda, kbeg, kend, ti * GT, M, vecA);
  load_tile_kn<TB>(rb, B, ldb, kbeg, kend, tj * GT, Nc, vecB);
  for (int64_t k0 = kbeg; k0 < kend; k0 += GK) {
    __syncthreads();
    store_tile_kn<TA>(ra, As);
    store_tile_kn<TB>(rb, Bs);
    __syncthreads();
    if (k0 + GK < kend) {
      load_tile_kn<TA>(ra, A, lda, k0 + GK, kend, ti * GT, M, vecA);
      load_tile_kn<TB>(rb, B, ldb, k0 + GK, kend, tj * GT, Nc, vecB);
    }
#pragma unroll
    for (int kk = 0; kk < GK / 16; ++kk) {
      double a[2][8], b[4][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int v1 = 0; v1 < 4; ++v1)
#pragma unroll
          for (int v0 = 0; v0 < 2; ++v0)
            a[i][v0 + 2 * v1] = As[(kk * 16 + lc + 4 * v1) * GS + wi * 32 + i * 16 + lr + 8 * v0];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) b[j][v] = Bs[(kk * 16 + lc + 4 * v) * GS + wj * 32 + j * 8 + lr];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma16x8x16(c[i][j], a[i], b[j]);
    }
  }
  double* out = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (GT * GT);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int v1 = 0; v1 < 2; ++v1) {
        const int r = wi * 32 + i * 16 + lr + 8 * v1, cc = wj * 32 + j * 8 + lc * 2;
        *reinterpret_cast<double2*>(out + r * GT + cc) = make_double2(c[i][j][2 * v1], c[i][j][2 * v1 + 1]);
      }
}

// C += sum_slices partial (fixed order), mirrored for symmetric products.
__global__ void atb_reduce_kernel(const double* __restrict__ partial, int ntiles, int nslices, int tiles_j,
                                  int symmetric, int M, int Nc, double* __restrict__ C, int64_t ldc) {
  const int tile = blockIdx.x;
  int ti, tj;
  if (symmetric) {
    tri_tile(tile, ti, tj);
  } else {
    ti = tile / tiles_j;
    tj = tile % tiles_j;
  }
  for (int e = threadIdx.x; e < GT * GT; e += blockDim.x) {
    double s = 0.0;
    for (int sl = 0; sl < nslices; ++sl) s += partial[((size_t)sl * ntiles + tile) * (GT * GT) + e];
    const int r = ti * GT + e / GT, c = tj * GT + e % GT;
    if (r < M && c < Nc) {
      if (!symmetric) {
        C[(int64_t)r * ldc + c] += s;
      } else if (c <= r) {
        C[(int64_t)r * ldc + c] += s;
        if (c < r) C[(int64_t)c * ldc + r] += s;
      }
    }
  }
}

// Few right-hand sides (B <= 8): R(FP x B) += DH^T F is a streaming GEMV bounded by the HBM read of DH.
// Each thread owns up to 3 feature columns (t, t+256, t+512) and B fp64 accumulators per column.
template <int NB, typename TA, typename TB>
__global__ void __launch_bounds__(256)
atb_small_kernel(const TA* __restrict__ A, int64_t lda, int M, const TB* __restrict__ B, int64_t ldb,
                 int nb, int64_t n, int64_t rows_per_cta, double* __restrict__ partial) {
  const int t = threadIdx.x;
  double acc[3][NB];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[q][j] = 0.0;
  const int64_t kbeg = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t kend = min(n, kbeg + rows_per_cta);
  const bool has2 = t + 512 < M;
  // U rows per step, every load of the step issued before the first use and no per-row predicates in the main
  // loop: the earlier form (4 rows, bounds test per row) compiled to ~3 loads in flight per thread and streamed
  // DH at 2.9 TB/s.
  constexpr int U = (NB <= 2) ? 8 : 4;
  int64_t k = kbeg;
  for (; k + U <= kend; k += U) {
    TA a[U][3];
    TB f[U][NB];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const TA* row = A + (k + u) * lda;
      a[u][0] = __ldg(row + t);
      a[u][1] = __ldg(row + t + 256);
      a[u][2] = has2 ? __ldg(row + t + 512) : (TA)0;
#pragma unroll
      for (int j = 0; j < NB; ++j) f[u][j] = (j < nb) ? __ldg(B + (k + u) * ldb + j) : (TB)0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[q][j] = fma((double)a[u][q], (double)f[u][j], acc[q][j]);
  }
  for (; k < kend; ++k) {   // tail rows
    const TA* row = A + k * lda;
    const TA a0 = __ldg(row + t), a1 = __ldg(row + t + 256), a2 = has2 ? __ldg(row + t + 512) : (TA)0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const double fj = (j < nb) ? (double)__ldg(B + k * ldb + j) : 0.0;
      acc[0][j] = fma((double)a0, fj, acc[0][j]);
      acc[1][j] = fma((double)a1, fj, acc[1][j]);
      acc[2][j] = fma((double)a2, fj, acc[2][j]);
    }
  }
  double* out = partial + (size_t)blockIdx.x * (768 * NB);
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int j = 0; j < NB; ++j) out[(t + q * 256) * NB + j] = acc[q][j];
}

// Vector form of the same kernel (lda % 4 == 0, A 16-byte aligned, M % 4 == 0): thread t owns the four columns
// 4t..4t+3 and reads them with one 16-byte load per row (two for fp64 inputs), U rows per step.  176 of the 192
// threads are active for M = 704.  Four times the bytes per load instruction of the scalar form.
template <typename T>
__device__ __forceinline__ void ldg4(const T* p, double* o);
template <>
__device__ __forceinline__ void ldg4<float>(const float* p, double* o) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void ldg4<double>(const double* p, double* o) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

template <int NB, typename TA, typename TB>
__global__ void __launch_bounds__(192)
atb_small_vec_kernel(const TA* __restrict__ A, int64_t lda, int M, const TB* __restrict__ B, int64_t ldb,
                     int nb, int64_t n, int64_t rows_per_cta, double* __restrict__ partial) {
  const int c0 = 4 * threadIdx.x;
  if (c0 >= M) return;
  double acc[4][NB];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[q][j] = 0.0;
  const int64_t kbeg = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t kend = min(n, kbeg + rows_per_cta);
  constexpr int U = (NB <= 2) ? 8 : 4;
  int64_t k = kbeg;
  for (; k + U <= kend; k += U) {
    double a[U][4];
    double f[U][NB];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ldg4<TA>(A + (k + u) * lda + c0, a[u]);
#pragma unroll
      for (int j = 0; j < NB; ++j) f[u][j] = (j < nb) ? (double)__ldg(B + (k + u) * ldb + j) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[q][j] = fma(a[u][q], f[u][j], acc[q][j]);
  }
  for (; k < kend; ++k) {
    double a[4];
    ldg4<TA>(A + k * lda + c0, a);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const double fj = (j < nb) ? (double)__ldg(B + k * ldb + j) : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q][j] = fma(a[q], fj, acc[q][j]);
    }
  }
  double* out = partial + (size_t)blockIdx.x * (768 * NB);
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < NB; ++j) out[(c0 + q) * NB + j] = acc[q][j];
}

template <int NB>
__global__ void atb_small_reduce_kernel(const double* __restrict__ partial, int nctas, int M, int B,
                                        double* __restrict__ C, int64_t ldc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * NB) return;
  const int r = e / NB, j = e % NB;
  if (j >= B) return;
  double s = 0.0;
  for (int c = 0; c < nctas; ++c) s += partial[(size_t)c * (768 * NB) + e];
  C[(int64_t)r * ldc + j] += s;
}

template <typename TA, typename TB>
int launch_atb_t(const fps_ctx* ctx, const TA* A, int64_t lda, int M, const TB* B, int64_t ldb, int Nc, int64_t n,
                 double* C, int64_t ldc, bool symmetric, cudaStream_t st) {
  if (n <= 0 || M <= 0 || Nc <= 0) return 0;
  FPS_REQUIRE(M <= FP, "atb: M=%d exceeds %d", M, FP);
  if (!symmetric && Nc <= 8 && M == FP && std::is_same<TA, float>::value && std::is_same<TB, float>::value) {
    // TMA-staged streaming form (stream_f32.cu); the register-staged kernel below serves what it does not take
    int nctas = 0;
    const int rc = launch_project_stream(ctx, reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(B), ldb,
                                         Nc, n, ctx->partial, &nctas, st);
    if (rc > 1) return rc;
    if (rc == 0) {
      const int nbk = Nc == 1 ? 1 : Nc == 2 ? 2 : Nc <= 4 ? 4 : 8;
      if (nbk == 1) atb_small_reduce_kernel<1><<<(M * 1 + 255) / 256, 256, 0, st>>>(ctx->partial, nctas, M, Nc, C, ldc);
      else if (nbk == 2) atb_small_reduce_kernel<2><<<(M * 2 + 255) / 256, 256, 0, st>>>(ctx->partial, nctas, M, Nc, C, ldc);
      else if (nbk == 4) atb_small_reduce_kernel<4><<<(M * 4 + 255) / 256, 256, 0, st>>>(ctx->partial, nctas, M, Nc, C, ldc);
      else atb_small_reduce_kernel<8><<<(M * 8 + 255) / 256, 256, 0, st>>>(ctx->partial, nctas, M, Nc, C, ldc);
      count_launch();
      FPS_CUDA(cudaGetLastError());
      return 0;
    }
  }
  if (!symmetric && Nc <= 8) {
    int64_t ctas = std::min<int64_t>((int64_t)ctx->sm_count * 8, (n + 63) / 64);
    ctas = std::min<int64_t>(ctas, (int64_t)(ctx->partial_bytes / (768 * 8 * sizeof(double))));
    const int64_t rpc = ((n + ctas - 1) / ctas + 7) / 8 * 8;
    ctas = (n + rpc - 1) / rpc;
    // columns t and t+256 are always < 704 <= lda; t+512 is guarded by has2
    FPS_REQUIRE(M > 512 && lda >= M, "atb small path expects the padded feature matrix");
    const bool vec = (lda % 4 == 0) && (M % 4 == 0) && (M <= 768) && ((uintptr_t)A % 16 == 0);
#define FPS_SMALL(NB_)                                                                                         \
  if (vec) atb_small_vec_kernel<NB_, TA, TB><<<(unsigned)ctas, 192, 0, st>>>(A, lda, M, B, ldb, Nc, n, rpc, ctx->partial); \
  else atb_small_kernel<NB_, TA, TB><<<(unsigned)ctas, 256, 0, st>>>(A, lda, M, B, ldb, Nc, n, rpc, ctx->partial);   \
  atb_small_reduce_kernel<NB_><<<(M * NB_ + 255) / 256, 256, 0, st>>>(ctx->partial, (int)ctas, M, Nc, C, ldc)
    if (Nc == 1) { FPS_SMALL(1); }
    else if (Nc == 2) { FPS_SMALL(2); }
    else if (Nc <= 4) { FPS_SMALL(4); }
    else { FPS_SMALL(8); }
#undef FPS_SMALL
    count_launch(2);
    FPS_CUDA(cudaGetLastError());
    return 0;
  }
  const int tiles_i = (M + GT - 1) / GT, tiles_j = (Nc + GT - 1) / GT;
  const int ntiles = symmetric ? tiles_i * (tiles_i + 1) / 2 : tiles_i * tiles_j;
  const size_t tile_bytes = (size_t)GT * GT * sizeof(double);
  const int64_t max_ctas = (int64_t)(ctx->partial_bytes / tile_bytes);
  FPS_REQUIRE(ntiles <= max_ctas, "atb: %d tiles exceed the partial workspace", ntiles);
  // Split the rows into `slices` so that tiles x slices fills whole waves of resident CTAs
  // (66 Gram tiles x 9 slices on 148 SMs x 3 CTAs was 1.34 waves = 67 % wave efficiency).
  static int occ = 0;
  if (!occ) {
    FPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, atb_f64_kernel<TA, TB>, GTHREADS, 0));
    if (occ < 1) occ = 1;
  }
  const int64_t wave = (int64_t)ctx->sm_count * occ;
  const int64_t smax = std::max<int64_t>(1, std::min<int64_t>({(int64_t)32, max_ctas / ntiles, (n + 4 * GK - 1) / (4 * GK)}));
  int64_t slices = 1;
  double best = 0.0;
  for (int64_t sl = 1; sl <= smax; ++sl) {
    const int64_t ctas = ntiles * sl;
    const double eff = (double)ctas / (double)(((ctas + wave - 1) / wave) * wave);
    if (eff > best + 0.02) { best = eff; slices = sl; }
  }
  const int64_t rps = ((n + slices - 1) / slices + GK - 1) / GK * GK;
  slices = (n + rps - 1) / rps;
  const int vecA = (lda % 4 == 0) && ((uintptr_t)A % 16 == 0);
  const int vecB = (ldb % 4 == 0) && ((uintptr_t)B % 16 == 0);
  dim3 grid((unsigned)ntiles, (unsigned)slices);
  atb_f64_kernel<TA, TB><<<grid, GTHREADS, 0, st>>>(A, lda, M, B, ldb, Nc, n, rps, tiles_j, symmetric ? 1 : 0,
                                                    ctx->partial, vecA, vecB);
  FPS_CUDA(cudaGetLastError());
  atb_reduce_kernel<<<ntiles, 256, 0, st>>>(ctx->partial, ntiles, (int)slices, tiles_j, symmetric ? 1 : 0, M, Nc, C,
                                            ldc);
  count_launch(2);
  FPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_atb_f64(const fps_ctx* ctx, const float* A, int64_t lda, int M, const float* B, int64_t ldb, int Ncols,
                   int64_t n, double* C, int64_t ldc, bool symmetric, cudaStream_t st) {
  return launch_atb_t<float, float>(ctx, A, lda, M, B, ldb, Ncols, n, C, ldc, symmetric, st);
}
int launch_atb_f64_d(const fps_ctx* ctx, const double* A, int64_t lda, int M, const double* B, int64_t ldb, int Ncols,
                     int64_t n, double* C, int64_t ldc, bool symmetric, cudaStream_t st) {
  return launch_atb_t<double, double>(ctx, A, lda, M, B, ldb, Ncols, n, C, ldc, symmetric, st);
}

// ------------------------------------------------------------------------------------------------
// Column sums in fp64 (H_bc^T 1, solver.py:176-178).  N_bc is small (4G+4); one CTA per 64 columns.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ A, int64_t n, int64_t ld, double* __restrict__ s64) {
  __shared__ double red[4][64];
  const int c = blockIdx.x * 64 + threadIdx.x % 64;
  const int g = threadIdx.x / 64;
  double s = 0.0;
  for (int64_t r = g; r < n; r += 4) s += (double)A[r * ld + c];
  red[g][threadIdx.x % 64] = s;
  __syncthreads();
  if (g == 0) s64[c] += red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

int launch_colsum(const float* A, int64_t n, int64_t ld, double* s64, cudaStream_t st) {
  if (n <= 0) return 0;
  colsum_kernel<float><<<FP / 64, 256, 0, st>>>(A, n, ld, s64);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}
int launch_colsum_d(const double* A, int64_t n, int64_t ld, double* s64, cudaStream_t st) {
  if (n <= 0) return 0;
  colsum_kernel<double><<<FP / 64, 256, 0, st>>>(A, n, ld, s64);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// out(n x B) = A(n x FP) W(FP x B) + bias, fp64 accumulation.
// Tile 64 rows x 64 columns, K steps of 32; A fragment loads want a row stride == 4 (mod 16).
// ------------------------------------------------------------------------------------------------
// Tile 128 rows x 128 columns, 8 warps (4 x 2, each 32 x 64), K steps of 16 with register prefetch of the
// next A / W slabs; fragment loads want row strides == 4 (mod 16) doubles.
constexpr int NT = 128, NK = 16, NS_A = NK + 4 /*20*/, NS_W = NT + 4 /*132*/, NTHREADS = 256;

template <typename TA, typename TO>
__global__ void __launch_bounds__(NTHREADS)
nn_f64_kernel(const TA* __restrict__ A, int64_t n, int64_t ld, const double* __restrict__ W, int64_t ldw,
              const double* __restrict__ bias, int B, TO* __restrict__ out, int64_t ldo, int vecW, int col_tiles) {
  __shared__ __align__(16) double As[NT * NS_A];
  __shared__ __align__(16) double Ws[NK * NS_W];
  const int64_t r0 = (int64_t)(blockIdx.x / col_tiles) * NT;  // 1-D grid: the row-tile count can exceed 65535
  const int c0 = (int)(blockIdx.x % col_tiles) * NT;
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int wi = warp / 2, wj = warp % 2;
  const int lr = lane / 4, lc = lane % 4;
  double c[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) c[i][j][0] = c[i][j][1] = 0.0;

  // staging maps: A slab 128 rows x 16 k = 512 groups of 4 (2 per thread); W slab 16 k x 128 cols = 1024
  // pairs (4 per thread)
  TA ra[2][4];
  double2 rw[4];
  auto load_slab = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = q * NTHREADS + t;
      const int r = idx / 4, kq = (idx % 4) * 4;
      double o[4] = {0.0, 0.0, 0.0, 0.0};
      if (r0 + r < n) load4<TA>(A + (r0 + r) * ld + k0 + kq, 0, 4, true, o);
#pragma unroll
      for (int e = 0; e < 4; ++e) ra[q][e] = (TA)o[e];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = q * NTHREADS + t;
      const int r = idx / 64, cc = (idx % 64) * 2;
      const double* p = W + (int64_t)(k0 + r) * ldw + c0 + cc;
      if (vecW && c0 + cc + 1 < B) rw[q] = __ldg(reinterpret_cast<const double2*>(p));
      else rw[q] = make_double2(c0 + cc < B ? __ldg(p) : 0.0, c0 + cc + 1 < B ? __ldg(p + 1) : 0.0);
    }
  };
  load_slab(0);
  for (int k0 = 0; k0 < FP; k0 += NK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = q * NTHREADS + t;
      const int r = idx / 4, kq = (idx % 4) * 4;
      double2* d = reinterpret_cast<double2*>(As + r * NS_A + kq);
      d[0] = make_double2((double)ra[q][0], (double)ra[q][1]);
      d[1] = make_double2((double)ra[q][2], (double)ra[q][3]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = q * NTHREADS + t;
      const int r = idx / 64, cc = (idx % 64) * 2;
      *reinterpret_cast<double2*>(Ws + r * NS_W + cc) = rw[q];
    }
    __syncthreads();
    if (k0 + NK < FP) load_slab(k0 + NK);
#pragma unroll
    for (int kk = 0; kk < NK / 4; ++kk) {
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(wi * 32 + i * 8 + lr) * NS_A + kk * 4 + lc];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Ws[(kk * 4 + lc) * NS_W + wj * 64 + j * 8 + lr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma8x8x4(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + wi * 32 + i * 8 + lr;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cc = c0 + wj * 64 + j * 8 + lc * 2;
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (cc + e < B) out[r * ldo + cc + e] = (TO)(c[i][j][e] + (bias ? bias[cc + e] : 0.0));
    }
  }
}

// Few right-hand sides: one warp per row, W columns staged in shared memory; HBM-bound on A.
template <int NB, typename TA, typename TO>
__global__ void __launch_bounds__(256)
nn_small_kernel(const TA* __restrict__ A, int64_t n, int64_t ld, const double* __restrict__ W, int64_t ldw,
                const double* __restrict__ bias, int B, TO* __restrict__ out, int64_t ldo) {
  __shared__ double Ws[FP * NB];
  for (int e = threadIdx.x; e < FP * NB; e += blockDim.x) {
    const int k = e / NB, j = e % NB;
    Ws[e] = (j < B) ? W[(int64_t)k * ldw + j] : 0.0;
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < n; r += nwarps) {
    const TA* row = A + r * ld;
    double acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = 0.0;
    double v[6][4];
#pragma unroll
    for (int q = 0; q < 6; ++q) {  // 176 groups of 4 per row: 5 full rounds of 32 lanes + 16
      const int i4 = q * 32 + lane;
      v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0.0;
      if (i4 < FP / 4) load4<TA>(row + i4 * 4, 0, 4, true, v[q]);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int i4 = q * 32 + lane;
      if (i4 < FP / 4) {
        const double* w = Ws + (size_t)i4 * 4 * NB;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          acc[j] = fma(v[q][0], w[0 * NB + j], acc[j]);
          acc[j] = fma(v[q][1], w[1 * NB + j], acc[j]);
          acc[j] = fma(v[q][2], w[2 * NB + j], acc[j]);
          acc[j] = fma(v[q][3], w[3 * NB + j], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (j < B) out[r * ldo + j] = (TO)(acc[j] + (bias ? bias[j] : 0.0));
    }
  }
}

template <typename TA, typename TO>
int launch_reconstruct_t(const fps_ctx* ctx, const TA* A, int64_t n, int64_t ld, const double* W, int64_t ldw,
                         const double* bias, int B, TO* out, int64_t ldo, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  FPS_REQUIRE(ld % 4 == 0 && (uintptr_t)A % 16 == 0 && ld >= FP,
              "reconstruct: A must be 16-byte aligned, ld %% 4 == 0, ld >= %d", FP);
  if (std::is_same<TA, float>::value && std::is_same<TO, float>::value) {
    const int rc = launch_recon_stream(ctx, reinterpret_cast<const float*>(A), n, ld, W, ldw, bias, B,
                                       reinterpret_cast<float*>(out), ldo, st);
    if (rc != 1) return rc;
  }
  if (B <= 4) {
    const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 8);
    if (B == 1) nn_small_kernel<1, TA, TO><<<(unsigned)blocks, 256, 0, st>>>(A, n, ld, W, ldw, bias, B, out, ldo);
    else if (B == 2) nn_small_kernel<2, TA, TO><<<(unsigned)blocks, 256, 0, st>>>(A, n, ld, W, ldw, bias, B, out, ldo);
    else nn_small_kernel<4, TA, TO><<<(unsigned)blocks, 256, 0, st>>>(A, n, ld, W, ldw, bias, B, out, ldo);
  } else {
    // column tiles fastest: the CTAs that share an A row block run together and hit it in L2
    const int col_tiles = (B + NT - 1) / NT;
    const int64_t ctas = (int64_t)col_tiles * ((n + NT - 1) / NT);
    FPS_REQUIRE(ctas < ((int64_t)1 << 31), "reconstruct: too many tiles (%lld)", (long long)ctas);
    const int vecW = (ldw % 2 == 0) && ((uintptr_t)W % 16 == 0);
    nn_f64_kernel<TA, TO><<<(unsigned)ctas, NTHREADS, 0, st>>>(A, n, ld, W, ldw, bias, B, out, ldo, vecW, col_tiles);
  }
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_reconstruct(const fps_ctx* ctx, const float* A, int64_t n, int64_t ld, const double* W, int64_t ldw,
                       const double* bias, int B, float* out, int64_t ldo, cudaStream_t st) {
  return launch_reconstruct_t<float, float>(ctx, A, n, ld, W, ldw, bias, B, out, ldo, st);
}
int launch_reconstruct_d(const fps_ctx* ctx, const double* A, int64_t n, int64_t ld, const double* W, int64_t ldw,
                         const double* bias, int B, double* out, int64_t ldo, cudaStream_t st) {
  return launch_reconstruct_t<double, double>(ctx, A, n, ld, W, ldw, bias, B, out, ldo, st);
}

// Batched form (multi-lambda sweep over B right-hand sides): sse[j] += sum_i (pred[i][j] - ref[i][j])^2.
// CTA = 32 columns x 8 row lanes over one row chunk; partial sums per chunk go to the split-K workspace and are summed in
// chunk order (bit-reproducible).  ref_const: one reference value per column instead of a matrix.
template <typename T>
__global__ void __launch_bounds__(256)
sse_columns_kernel(const T* __restrict__ pred, int64_t ldp, const T* __restrict__ ref, int64_t ldr,
                   const double* __restrict__ ref_const, int64_t n, int B, int64_t rows_per_chunk,
                   double* __restrict__ partial) {
  __shared__ double red[8][33];
  const int cl = threadIdx.x % 32, rl = threadIdx.x / 32;
  const int j = blockIdx.x * 32 + cl;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
  double s = 0.0;
  if (j < B) {
    const double rc = ref_const ? ref_const[j] : 0.0;
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const double d = (double)pred[r * ldp + j] - (ref_const ? rc : (double)ref[r * ldr + j]);
      s = fma(d, d, s);
    }
  }
  red[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && j < B) {
#pragma unroll
    for (int q = 1; q < 8; ++q) s += red[q][cl];
    partial[(size_t)blockIdx.y * B + j] = s;
  }
}

__global__ void sse_reduce_kernel(const double* __restrict__ partial, int chunks, int B, double* __restrict__ sse) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= B) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * B + j];
  sse[j] += s;
}

template <typename T>
static int launch_sse_columns_t(const fps_ctx* ctx, const T* pred, int64_t ldp, const T* ref, int64_t ldr,
                                const double* ref_const, int64_t n, int B, double* sse, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  int64_t chunks = std::min<int64_t>({(int64_t)512, (n + 63) / 64, (int64_t)(ctx->partial_bytes / ((size_t)B * sizeof(double)))});
  FPS_REQUIRE(chunks >= 1, "sse_columns: %d columns exceed the partial workspace", B);
  const int64_t rpc = (n + chunks - 1) / chunks;
  chunks = (n + rpc - 1) / rpc;
  sse_columns_kernel<T><<<dim3((unsigned)((B + 31) / 32), (unsigned)chunks), 256, 0, st>>>(pred, ldp, ref, ldr, ref_const, n, B,
                                                                                          rpc, ctx->partial);
  sse_reduce_kernel<<<(B + 255) / 256, 256, 0, st>>>(ctx->partial, (int)chunks, B, sse);
  count_launch(2);
  FPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_sse_columns(const fps_ctx* ctx, const float* pred, int64_t ldp, const float* ref, int64_t ldr,
                       const double* ref_const, int64_t n, int B, double* sse, cudaStream_t st) {
  return launch_sse_columns_t<float>(ctx, pred, ldp, ref, ldr, ref_const, n, B, sse, st);
}
int launch_sse_columns_d(const fps_ctx* ctx, const double* pred, int64_t ldp, const double* ref, int64_t ldr,
                         const double* ref_const, int64_t n, int B, double* sse, cudaStream_t st) {
  return launch_sse_columns_t<double>(ctx, pred, ldp, ref, ldr, ref_const, n, B, sse, st);
}

}  // namespace fps
