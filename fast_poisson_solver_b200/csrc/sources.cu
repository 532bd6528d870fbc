// On-device synthetic right-hand sides: the sine family of the reference's data generator,
//   f_j(x, y) = sum_i s_ij * sin(k1_ij*pi*(x - o1_ij)) * {sin|cos}(k2_ij*pi*(y - o2_ij))
// (data_utils/source_functions.py:24-73: k ~ U(-10,10), o ~ U(-1,1), s ~ U(-100,100), 1-9 terms), written
// straight into the (n, B) matrix that fps_project_rhs consumes, so that thousands of sources never
// exist on the host (16M points x 4096 columns would be 275 GB).
#include "fps_common.cuh"
#include <algorithm>
#include <math.h>

namespace fps {

// params: (B, T, 6) doubles = {k1, o1, k2, o2, s, use_cos}; terms with s == 0 are padding.
template <typename TO>
__global__ void __launch_bounds__(256)
sine_sources_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                    const double* __restrict__ params, int T, int B, TO* __restrict__ F, int64_t ldf) {
  extern __shared__ double sp[];  // this CTA's 32 columns x T terms x 6
  const int c0 = blockIdx.y * 32;
  const int nc = min(32, B - c0);
  for (int e = threadIdx.x; e < nc * T * 6; e += blockDim.x) sp[e] = params[(size_t)c0 * T * 6 + e];
  __syncthreads();
  const int col = threadIdx.x % 32, rsub = threadIdx.x / 32;
  for (int64_t r = (int64_t)blockIdx.x * 8 + rsub; r < n; r += (int64_t)gridDim.x * 8) {
    if (col >= nc) continue;
    const double xv = x[r], yv = y[r];
    double acc = 0.0;
    for (int t = 0; t < T; ++t) {
      const double* q = sp + ((size_t)col * T + t) * 6;
      if (q[4] == 0.0) continue;
      const double a = sinpi(q[0] * (xv - q[1]));
      const double ph = q[2] * (yv - q[3]);
      acc = fma(q[4] * a, q[5] != 0.0 ? cospi(ph) : sinpi(ph), acc);
    }
    F[r * ldf + c0 + col] = (TO)acc;
  }
}

template <typename TO>
static int launch_sine_sources_t(const fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params,
                                 int T, int B, TO* F, int64_t ldf, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  FPS_REQUIRE(T >= 1 && T <= 32 && ldf >= B, "sine_sources: bad argument (T=%d)", T);
  dim3 grid((unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 16), (unsigned)((B + 31) / 32));
  sine_sources_kernel<TO><<<grid, 256, 32 * T * 6 * sizeof(double), st>>>(x, y, n, params, T, B, F, ldf);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

// ---- Perlin family (data_utils/source_functions.py:173-227) ----------------------------------------------------
// params: (B, 6, 3) doubles; rows 0..4 = additive layers {octave, amplitude, seed} (amplitude 0 = unused), row 5 =
// the multiplicative modulation layer of source_functions.py:213-220 (amplitude 0 = none).  One layer is classic
// gradient noise on the lattice of (x, y) * octave with the quintic fade; the lattice gradients come from an integer
// hash (the reference's `perlin_noise` package - a per-point Python loop, absent here - draws them from a seeded
// RNG: same family, different sample; sources are INPUTS to the path).  Host mirror: sources.py:perlin_sources_numpy.
__device__ __forceinline__ uint32_t lattice_hash(int ix, int iy, uint32_t seed) {
  uint32_t h = ((uint32_t)ix * 0x9E3779B1u) ^ ((uint32_t)iy * 0x85EBCA77u) ^ (seed * 0xC2B2AE3Du);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}

// 64 unit gradient directions, angle 2 pi k / 64 (filled once per context by sources_init: the same fp64 values as
// numpy's cos / sin in sources.py); the top 6 bits of the lattice hash pick one
__constant__ double c_grad[64][2];

__device__ __forceinline__ double perlin_layer(double x, double y, double octave, uint32_t seed) {
  const double px = x * octave, py = y * octave;
  const double fx0 = floor(px), fy0 = floor(py);
  const int ix = (int)fx0, iy = (int)fy0;
  const double fx = px - fx0, fy = py - fy0;
  const double sx = fx * fx * fx * (fx * (fx * 6.0 - 15.0) + 10.0);
  const double sy = fy * fy * fy * (fy * (fy * 6.0 - 15.0) + 10.0);
  double n[2][2];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const uint32_t g = lattice_hash(ix + dx, iy + dy, seed) >> 26;
      n[dy][dx] = c_grad[g][0] * (fx - dx) + c_grad[g][1] * (fy - dy);
    }
  return (n[0][0] * (1.0 - sx) + n[0][1] * sx) * (1.0 - sy) + (n[1][0] * (1.0 - sx) + n[1][1] * sx) * sy;
}

template <typename TO>
__global__ void __launch_bounds__(256)
perlin_sources_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                      const double* __restrict__ params, int B, TO* __restrict__ F, int64_t ldf) {
  __shared__ double sp[32 * 18];  // this CTA's 32 columns x 6 rows x 3
  const int c0 = blockIdx.y * 32;
  const int nc = min(32, B - c0);
  for (int e = threadIdx.x; e < nc * 18; e += blockDim.x) sp[e] = params[(size_t)c0 * 18 + e];
  __syncthreads();
  const int col = threadIdx.x % 32, rsub = threadIdx.x / 32;
  for (int64_t r = (int64_t)blockIdx.x * 8 + rsub; r < n; r += (int64_t)gridDim.x * 8) {
    if (col >= nc) continue;
    const double xv = x[r], yv = y[r];
    const double* q = sp + col * 18;
    double u = 0.0;
#pragma unroll 1
    for (int l = 0; l < 5; ++l)
      if (q[3 * l + 1] != 0.0) u += q[3 * l + 1] * perlin_layer(xv, yv, q[3 * l], (uint32_t)q[3 * l + 2]);
    if (q[16] != 0.0) u *= q[16] * perlin_layer(xv, yv, q[15], (uint32_t)q[17]);
    F[r * ldf + c0 + col] = (TO)u;
  }
}

int sources_init(fps_ctx* ctx) {
  (void)ctx;
  double g[64][2];
  for (int k = 0; k < 64; ++k) {
    const double a = 2.0 * 3.14159265358979323846 * (double)k / 64.0;
    g[k][0] = cos(a);
    g[k][1] = sin(a);
  }
  FPS_CUDA(cudaMemcpyToSymbol(c_grad, g, sizeof(g)));
  return 0;
}

template <typename TO>
static int launch_perlin_sources_t(const fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params,
                                   int B, TO* F, int64_t ldf, cudaStream_t st) {
  if (n <= 0 || B <= 0) return 0;
  FPS_REQUIRE(ldf >= B, "perlin_sources: ldf < B");
  dim3 grid((unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 16), (unsigned)((B + 31) / 32));
  perlin_sources_kernel<TO><<<grid, 256, 0, st>>>(x, y, n, params, B, F, ldf);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

// ---- problem geometry of the reference's generator (data.py:192-217) ---------------------------------------------
// interior point p (global index, x fastest like np.meshgrid + flatten): x = g[1 + p % G], y = g[1 + p / G] with
// g = linspace(0, 1, G + 2); boundary ring of 4G + 4 points in the order [bottom row (G+2), top row (G+2), left
// column (G), right column (G)].  [first, first + count) selects a rank's shard.
__global__ void grid_points_kernel(int G, int64_t first, int64_t count, double* __restrict__ x, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t p = first + i;
  const double step = 1.0 / (double)(G + 1);
  const int ix = (int)(p % G), iy = (int)(p / G);
  // linspace(0, 1, G + 2)[k] = k * step computed as numpy does (start + k * step), the last point forced to 1
  x[i] = (double)(ix + 1) * step;
  y[i] = (double)(iy + 1) * step;
}

__global__ void ring_points_kernel(int G, int64_t first, int64_t count, double* __restrict__ x, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t p = first + i;
  const double step = 1.0 / (double)(G + 1);
  auto lin = [&](int64_t k) { return k == G + 1 ? 1.0 : (double)k * step; };
  double xv, yv;
  if (p < G + 2) { xv = lin(p); yv = 0.0; }
  else if (p < 2 * (int64_t)(G + 2)) { xv = lin(p - (G + 2)); yv = 1.0; }
  else if (p < 2 * (int64_t)(G + 2) + G) { xv = 0.0; yv = lin(p - 2 * (int64_t)(G + 2) + 1); }
  else { xv = 1.0; yv = lin(p - 2 * (int64_t)(G + 2) - G + 1); }
  x[i] = xv;
  y[i] = yv;
}

}  // namespace fps

using namespace fps;

extern "C" {
int fps_sine_sources(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params, int n_terms,
                     int B, float* F, int64_t ldf, void* stream) {
  FPS_REQUIRE(ctx && x && y && params && F, "fps_sine_sources: null argument");
  return launch_sine_sources_t<float>(ctx, x, y, n, params, n_terms, B, F, ldf, (cudaStream_t)stream);
}
int fps_sine_sources_f64(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params, int n_terms,
                         int B, double* F, int64_t ldf, void* stream) {
  FPS_REQUIRE(ctx && x && y && params && F, "fps_sine_sources_f64: null argument");
  return launch_sine_sources_t<double>(ctx, x, y, n, params, n_terms, B, F, ldf, (cudaStream_t)stream);
}
int fps_perlin_sources(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params, int B,
                       float* F, int64_t ldf, void* stream) {
  FPS_REQUIRE(ctx && x && y && params && F, "fps_perlin_sources: null argument");
  return launch_perlin_sources_t<float>(ctx, x, y, n, params, B, F, ldf, (cudaStream_t)stream);
}
int fps_perlin_sources_f64(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params, int B,
                           double* F, int64_t ldf, void* stream) {
  FPS_REQUIRE(ctx && x && y && params && F, "fps_perlin_sources_f64: null argument");
  return launch_perlin_sources_t<double>(ctx, x, y, n, params, B, F, ldf, (cudaStream_t)stream);
}
int fps_grid_points(fps_ctx* ctx, int G, int64_t first_pde, int64_t n_pde, double* x_pde, double* y_pde,
                    int64_t first_bc, int64_t n_bc, double* x_bc, double* y_bc, void* stream) {
  FPS_REQUIRE(ctx && G >= 1, "fps_grid_points: bad argument");
  FPS_REQUIRE(first_pde >= 0 && n_pde >= 0 && first_pde + n_pde <= (int64_t)G * G, "fps_grid_points: interior range");
  FPS_REQUIRE(first_bc >= 0 && n_bc >= 0 && first_bc + n_bc <= 4 * (int64_t)G + 4, "fps_grid_points: boundary range");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_pde > 0) {
    FPS_REQUIRE(x_pde && y_pde, "fps_grid_points: null interior output");
    grid_points_kernel<<<(unsigned)((n_pde + 255) / 256), 256, 0, st>>>(G, first_pde, n_pde, x_pde, y_pde);
    count_launch();
  }
  if (n_bc > 0) {
    FPS_REQUIRE(x_bc && y_bc, "fps_grid_points: null boundary output");
    ring_points_kernel<<<(unsigned)((n_bc + 255) / 256), 256, 0, st>>>(G, first_bc, n_bc, x_bc, y_bc);
    count_launch();
  }
  FPS_CUDA(cudaGetLastError());
  return 0;
}
}
