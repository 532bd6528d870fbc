// On-device synthetic right-hand sides: the sine family of the reference's data generator,
//   f_j(x, y) = sum_i s_ij * sin(k1_ij*pi*(x - o1_ij)) * {sin|cos}(k2_ij*pi*(y - o2_ij))
// (data_utils/source_functions.py:24-73: k ~ U(-10,10), o ~ U(-1,1), s ~ U(-100,100), 1-9 terms), written
// straight into the (n, B) matrix that fps_project_rhs consumes, so that thousands of sources never
// exist on the host (16M points x 4096 columns would be 275 GB).
#include "fps_common.cuh"
#include <algorithm>

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
}
