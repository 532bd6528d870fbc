// Finite-difference accuracy oracle on the GPU: the reference's `numeric_solve` (numeric_solver.py:104-151 with the
// operators of diff_matrices.py:22-59) solves, by sparse LU, the 5-point system
//     (u[i-1,j] - 2 u[i,j] + u[i+1,j]) / hx^2 + (u[i,j-1] - 2 u[i,j] + u[i,j+1]) / hy^2 = f[i,j]     interior points
//     u = u_bc                                                                                        boundary rows
// (numeric_solver.py:112 builds D2x/dx^2 + D2y/dy^2, :138-146 overwrites every boundary row with an identity row; the
// one-sided rows of diff_matrices.py:26,33 never survive).  With the boundary values moved to the right-hand side the
// interior block is the Dirichlet 5-point operator, whose eigenvectors are sines: the SAME linear system is solved here
// exactly (to fp64 rounding) by a type-I discrete sine transform along each axis,
//     U = c (S_y ((S_y F S_x) ./ (lam_y[j] + lam_x[i])) S_x),   S_n[p][q] = sin(pi (p+1)(q+1) / (n+1)),
//     lam_n[p] = (2 cos(pi (p+1)/(n+1)) - 2) / h^2,   c = 4 / ((Gx+1)(Gy+1)),
// as four dense fp64 GEMMs per right-hand side on the DMMA tensor path (69 GFLOP for a 2048^2 grid: ~3 ms; the sparse
// LU of the reference needs 5 s for 400^2 and does not fit a host at 2048^2).  A dense sine matrix instead of an FFT:
// the transform is tiny next to the solver's own work, fp64-exact phases come for free from integer argument
// reduction, and no library FFT is involved.
//
// Layout: the caller's point ordering is data.py:192-198 (meshgrid: x fastest), i.e. grid point (iy, ix) is row
// iy * Gx + ix of the (Gx Gy, B) matrices F and U; one column per right-hand side.
#include "fps_common.cuh"
#include <algorithm>

namespace fps {

// S[p][q] = sin(pi (p+1)(q+1) / (n+1)), argument reduced in integers (exact), row-major n x n
__global__ void dst_matrix_kernel(double* __restrict__ S, int n) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * n) return;
  const int p = (int)(e / n), q = (int)(e % n);
  const int64_t m = ((int64_t)(p + 1) * (q + 1)) % (2 * (int64_t)(n + 1));
  S[e] = sinpi((double)m / (double)(n + 1));
}

// eigenvalues of the 1-D Dirichlet second difference, lam[p] = (2 cos(pi (p+1)/(n+1)) - 2) / h^2
__global__ void dst_eig_kernel(double* __restrict__ lam, int n, double inv_h2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) lam[p] = (2.0 * cospi((double)(p + 1) / (double)(n + 1)) - 2.0) * inv_h2;
}

// T[b][iy][ix] = F[(iy Gx + ix) ldf + b0 + b]  (gather one batch of columns into contiguous grids, widened to fp64)
template <typename TF>
__global__ void __launch_bounds__(256)
fd_gather_kernel(const TF* __restrict__ F, int64_t ldf, int b0, int nb, int64_t npts, double* __restrict__ T) {
  __shared__ double tile[32][33];
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    tile[r][tx] = (p < npts && c0 + tx < nb) ? (double)F[p * ldf + b0 + c0 + tx] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int b = c0 + r;
    const int64_t p = p0 + tx;
    if (b < nb && p < npts) T[(int64_t)b * npts + p] = tile[tx][r];
  }
}

// U[(iy Gx + ix) ldu + b0 + b] = scale * T[b][iy][ix] + bc[b0 + b]
template <typename TU>
__global__ void __launch_bounds__(256)
fd_scatter_kernel(const double* __restrict__ T, int64_t npts, int b0, int nb, double scale,
                  const double* __restrict__ bc, TU* __restrict__ U, int64_t ldu) {
  __shared__ double tile[32][33];
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int r = ty; r < 32; r += 8) {
    const int b = c0 + r;
    const int64_t p = p0 + tx;
    tile[r][tx] = (b < nb && p < npts) ? T[(int64_t)b * npts + p] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int b = c0 + tx;
    if (p < npts && b < nb) U[p * ldu + b0 + b] = (TU)(scale * tile[tx][r] + (bc ? bc[b0 + b] : 0.0));
  }
}

// ---- batched fp64 GEMM on the DMMA path:  C_b (m x n) = A_b (m x k) B_b (k x n), row-major, optional spectral divide
// 128 x 128 tile, 8 warps (4 x 2, each 32 x 64), k steps of 16, m8n8k4 DMMA; strides of 0 share an operand (the sine
// matrix) between the batches.  divide: C[i][j] /= (lam_r[i] + lam_c[j]).
constexpr int DT = 128, DK = 16, DS_A = DK + 4, DS_B = DT + 4, DTHREADS = 256;

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(DTHREADS)
dgemm_batched_kernel(const double* __restrict__ A, int64_t strideA, const double* __restrict__ B, int64_t strideB,
                     double* __restrict__ C, int64_t strideC, int m, int n, int k, const double* __restrict__ lam_r,
                     const double* __restrict__ lam_c) {
  __shared__ __align__(16) double As[DT * DS_A];
  __shared__ __align__(16) double Bs[DK * DS_B];
  A += (int64_t)blockIdx.z * strideA;
  B += (int64_t)blockIdx.z * strideB;
  C += (int64_t)blockIdx.z * strideC;
  const int r0 = blockIdx.y * DT, c0 = blockIdx.x * DT;
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int wi = warp / 2, wj = warp % 2;
  const int lr = lane / 4, lc = lane % 4;
  double c[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) c[i][j][0] = c[i][j][1] = 0.0;

  double ra[8], rb[8];
  auto load_slab = [&](int k0) {
    // A slab: 128 rows x 16 k = 2048 elements, 8 per thread (row = idx / 16: consecutive threads walk k)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = q * DTHREADS + t;
      const int r = idx / DK, kk = idx % DK;
      ra[q] = (r0 + r < m && k0 + kk < k) ? __ldg(A + (int64_t)(r0 + r) * k + k0 + kk) : 0.0;
    }
    // B slab: 16 k x 128 columns
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = q * DTHREADS + t;
      const int kk = idx / DT, cc = idx % DT;
      rb[q] = (k0 + kk < k && c0 + cc < n) ? __ldg(B + (int64_t)(k0 + kk) * n + c0 + cc) : 0.0;
    }
  };
  load_slab(0);
  for (int k0 = 0; k0 < k; k0 += DK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = q * DTHREADS + t;
      As[(idx / DK) * DS_A + idx % DK] = ra[q];
      Bs[(idx / DT) * DS_B + idx % DT] = rb[q];
    }
    __syncthreads();
    if (k0 + DK < k) load_slab(k0 + DK);
#pragma unroll
    for (int kk = 0; kk < DK / 4; ++kk) {
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(wi * 32 + i * 8 + lr) * DS_A + kk * 4 + lc];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[(kk * 4 + lc) * DS_B + wj * 64 + j * 8 + lr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma_8x8x4(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + wi * 32 + i * 8 + lr;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = c0 + wj * 64 + j * 8 + lc * 2 + e;
        if (cc >= n) continue;
        double v = c[i][j][e];
        if (lam_r) v /= (lam_r[r] + lam_c[cc]);
        C[(int64_t)r * n + cc] = v;
      }
    }
  }
}

static void dgemm_batched(const double* A, int64_t sA, const double* B, int64_t sB, double* C, int64_t sC, int m, int n,
                          int k, int batch, const double* lam_r, const double* lam_c, cudaStream_t st) {
  dim3 grid((unsigned)((n + DT - 1) / DT), (unsigned)((m + DT - 1) / DT), (unsigned)batch);
  dgemm_batched_kernel<<<grid, DTHREADS, 0, st>>>(A, sA, B, sB, C, sC, m, n, k, lam_r, lam_c);
  count_launch();
}

// work: caller-owned scratch of fps_fd_workspace_bytes(Gx, Gy, nb) bytes
template <typename TF, typename TU>
static int fd_solve_t(const fps_ctx* ctx, int Gx, int Gy, double hx, double hy, const TF* F, int64_t ldf, const double* bc,
                      int B, TU* U, int64_t ldu, void* work, size_t work_bytes, cudaStream_t st) {
  FPS_REQUIRE(Gx >= 1 && Gy >= 1 && B >= 1 && hx > 0 && hy > 0, "fps_fd_solve: bad grid");
  FPS_REQUIRE(ldf >= B && ldu >= B, "fps_fd_solve: bad leading dimension");
  const int64_t npts = (int64_t)Gx * Gy;
  const size_t fixed = ((size_t)Gx * Gx + (size_t)Gy * Gy + Gx + Gy) * sizeof(double);
  FPS_REQUIRE(work && work_bytes >= fixed + 2 * (size_t)npts * sizeof(double),
              "fps_fd_solve: workspace of %zu bytes is too small (need >= %zu)", work_bytes,
              fixed + 2 * (size_t)npts * sizeof(double));
  double* Sx = static_cast<double*>(work);
  double* Sy = Sx + (size_t)Gx * Gx;
  double* lx = Sy + (size_t)Gy * Gy;
  double* ly = lx + Gx;
  double* T0 = ly + Gy;
  const int64_t nb_max = (int64_t)((work_bytes - fixed) / (2 * (size_t)npts * sizeof(double)));
  const int nbatch = (int)std::min<int64_t>(nb_max, B);
  double* T1 = T0 + (size_t)nbatch * npts;
  dst_matrix_kernel<<<(unsigned)(((int64_t)Gx * Gx + 255) / 256), 256, 0, st>>>(Sx, Gx);
  dst_matrix_kernel<<<(unsigned)(((int64_t)Gy * Gy + 255) / 256), 256, 0, st>>>(Sy, Gy);
  dst_eig_kernel<<<(Gx + 255) / 256, 256, 0, st>>>(lx, Gx, 1.0 / (hx * hx));
  dst_eig_kernel<<<(Gy + 255) / 256, 256, 0, st>>>(ly, Gy, 1.0 / (hy * hy));
  count_launch(4);
  const double scale = 4.0 / ((double)(Gx + 1) * (double)(Gy + 1));
  for (int b0 = 0; b0 < B; b0 += nbatch) {
    const int nb = std::min(nbatch, B - b0);
    dim3 gg((unsigned)((npts + 31) / 32), (unsigned)((nb + 31) / 32));
    fd_gather_kernel<TF><<<gg, 256, 0, st>>>(F, ldf, b0, nb, npts, T0);
    count_launch();
    // forward transform: T1 = Sy T0 (along iy), T0 = (T1 Sx) ./ (ly[iy] + lx[ix])
    dgemm_batched(Sy, 0, T0, npts, T1, npts, Gy, Gx, Gy, nb, nullptr, nullptr, st);
    dgemm_batched(T1, npts, Sx, 0, T0, npts, Gy, Gx, Gx, nb, ly, lx, st);
    // inverse transform (DST-I is its own inverse up to the factor folded into `scale`)
    dgemm_batched(Sy, 0, T0, npts, T1, npts, Gy, Gx, Gy, nb, nullptr, nullptr, st);
    dgemm_batched(T1, npts, Sx, 0, T0, npts, Gy, Gx, Gx, nb, nullptr, nullptr, st);
    fd_scatter_kernel<TU><<<gg, 256, 0, st>>>(T0, npts, b0, nb, scale, bc, U, ldu);
    count_launch();
  }
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps

using namespace fps;

extern "C" {

size_t fps_fd_workspace_bytes(int Gx, int Gy, int batch) {
  if (Gx < 1 || Gy < 1 || batch < 1) return 0;
  return ((size_t)Gx * Gx + (size_t)Gy * Gy + Gx + Gy + 2 * (size_t)batch * Gx * Gy) * sizeof(double);
}

int fps_fd_solve(fps_ctx* ctx, int Gx, int Gy, double hx, double hy, const double* F, int64_t ldf, const double* bc,
                 int B, double* U, int64_t ldu, void* work, size_t work_bytes, void* stream) {
  FPS_REQUIRE(ctx && F && U, "fps_fd_solve: null argument");
  return fd_solve_t<double, double>(ctx, Gx, Gy, hx, hy, F, ldf, bc, B, U, ldu, work, work_bytes, (cudaStream_t)stream);
}

int fps_fd_solve_f32(fps_ctx* ctx, int Gx, int Gy, double hx, double hy, const float* F, int64_t ldf, const double* bc,
                     int B, double* U, int64_t ldu, void* work, size_t work_bytes, void* stream) {
  FPS_REQUIRE(ctx && F && U, "fps_fd_solve_f32: null argument");
  return fd_solve_t<float, double>(ctx, Gx, Gy, hx, hy, F, ldf, bc, B, U, ldu, work, work_bytes, (cudaStream_t)stream);
}

}  // extern "C"
