// Exact-operand descriptors (struct Xop, fps_common.cuh): per-column maxima -> exponents -> fixed-point scales of
// the int8 engine (gram_i8.cu: Gram `DH.t() @ DH` solver.py:167 and projection `RHSs @ f` :301; recon_i8.cu: the GEMVs
// of solver.py:327-330).  One rule for every consumer, so that a matrix snapped by one of them stays put in the others.
#include "fps_common.cuh"
#include <algorithm>

namespace fps {

// per-column max |a| over rows [r0, r1) of this CTA, merged with atomicMax on the float bits
__global__ void __launch_bounds__(256)
xop_colmax_kernel(const float* __restrict__ A, int64_t ld, int64_t n, int ncols, int64_t rows_per_cta,
                  unsigned* __restrict__ cmax) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  const int c0 = blockIdx.y * 768;
  const int t = c0 + threadIdx.x;
  const bool h0 = t < ncols, h1 = t + 256 < ncols, h2 = t + 512 < ncols;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f;
  for (int64_t r = r0; r < r1; r += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r + u < r1) {
        const float* row = A + (r + u) * ld;
        if (h0) m0 = fmaxf(m0, fabsf(__ldg(row + t)));
        if (h1) m1 = fmaxf(m1, fabsf(__ldg(row + t + 256)));
        if (h2) m2 = fmaxf(m2, fabsf(__ldg(row + t + 512)));
      }
    }
  }
  if (h0) atomicMax(cmax + t, __float_as_uint(m0));
  if (h1) atomicMax(cmax + t + 256, __float_as_uint(m1));
  if (h2) atomicMax(cmax + t + 512, __float_as_uint(m2));
}

// exponent e with max < 2^e, and the fixed-point scale 2^(30 - e) with its inverse (fp32 powers of two; e is
// clamped to >= -90 so that both stay normal numbers: a column below 2^-90 is kept on a coarser grid).
// headroom: added to e (the calibrated grids of the fused feature epilogue are chosen from a sample of the points).
// cond != null: run only when *cond != 0 (device-side conditional for the overflow fallback).
__global__ void xop_exps_kernel(const unsigned* __restrict__ cmax, int ncols, int rows, int headroom,
                                int* __restrict__ exps, float2* __restrict__ scale, const int* __restrict__ cond) {
  if (cond && *cond == 0) return;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= rows) return;
  int e = 0;
  if (f < ncols) {
    const unsigned bits = cmax[f];
    const int E = (int)(bits >> 23);
    e = bits ? ((E > 0 ? E : 1) - 126 + headroom) : 0;
    if (e < -90) e = -90;
    if (e > 120) e = 120;
  }
  exps[f] = e;
  scale[f] = make_float2(ldexpf(1.0f, 30 - e), ldexpf(1.0f, e - 30));
}

// every column on the same grid 2^(e - 30) (H = tanh(.) lies in [-1, 1]: e = 0 needs no maxima at all)
__global__ void xop_uniform_kernel(int e, int rows, Xop* __restrict__ x) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= rows) return;
  x->exps[f] = e;
  x->scale[f] = make_float2(ldexpf(1.0f, 30 - e), ldexpf(1.0f, e - 30));
  x->cmax[f] = 0u;
  if (f < 4) x->flags[f] = 0;
}

// A[i][f] <- rint(A[i][f] 2^(30 - e_f)) 2^(e_f - 30) in place (only elements that move are written)
__global__ void __launch_bounds__(256)
xop_snap_kernel(float* __restrict__ A, int64_t ld, int64_t n, int ncols, const float2* __restrict__ scale) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = idx / (ncols / 4);
  const int k = (int)(idx % (ncols / 4)) * 4;
  if (i >= n) return;
  float4 v = *reinterpret_cast<const float4*>(A + i * ld + k);
  float vv[4] = {v.x, v.y, v.z, v.w};
  bool changed = false;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 sc = scale[k + u];
    const float snapped = (float)__float2int_rn(vv[u] * sc.x) * sc.y;
    if (snapped != vv[u]) { vv[u] = snapped; changed = true; }
  }
  if (changed) *reinterpret_cast<float4*>(A + i * ld + k) = make_float4(vv[0], vv[1], vv[2], vv[3]);
}

int xop_snap(float* A, int64_t lda, int64_t n, int ncols, const Xop* x, cudaStream_t st) {
  FPS_REQUIRE(ncols % 4 == 0 && lda % 4 == 0 && (uintptr_t)A % 16 == 0, "xop_snap: alignment");
  const int64_t threads = n * (ncols / 4);
  xop_snap_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(A, lda, n, ncols, x->scale);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

int xop_colmax(const fps_ctx* ctx, const float* A, int64_t lda, int64_t n, int ncols, unsigned* cmax, cudaStream_t st) {
  FPS_CUDA(cudaMemsetAsync(cmax, 0, (size_t)ncols * sizeof(unsigned), st));
  if (n <= 0) return 0;
  int64_t ctas = std::min<int64_t>((int64_t)ctx->sm_count * 8, (n + 63) / 64);
  const int64_t rpc = ((n + ctas - 1) / ctas + 3) / 4 * 4;
  ctas = (n + rpc - 1) / rpc;
  xop_colmax_kernel<<<dim3((unsigned)ctas, (unsigned)((ncols + 767) / 768)), 256, 0, st>>>(A, lda, n, ncols, rpc, cmax);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

int xop_exps(const unsigned* cmax, int ncols, int rows, int headroom, int* exps, float2* scale, const int* cond,
             cudaStream_t st) {
  xop_exps_kernel<<<(rows + 255) / 256, 256, 0, st>>>(cmax, ncols, rows, headroom, exps, scale, cond);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

int xop_uniform(Xop* x, int e, cudaStream_t st) {
  xop_uniform_kernel<<<(XOP_ROWS + 255) / 256, 256, 0, st>>>(e, XOP_ROWS, x);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
