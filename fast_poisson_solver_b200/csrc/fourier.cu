// Fourier feature layer with analytic derivative streams.
// Reference: FourierFeatureMapping.forward, model.py:134-139 (x@freqs + biases, sin/cos, cat) and
// the derivatives that utils.py:54-68 obtains from autograd, written out in closed form
// (SURVEY.md 3.3):  z = x F0 + y F1 + beta ; phi = [c sin z, c cos z]
//   phi_x = [ c cos z F0, -c sin z F0 ],  phi_y likewise with F1,  phi_lap = -(F0^2+F1^2) phi.
// z and sin/cos are evaluated in fp64 (|z| reaches ~40 rad and the Laplacian multiplies by up to
// ~800, so an fp32 argument would cost ~1e-6 of relative accuracy before the first GEMM); the
// results are rounded once to fp32 and split into the hi/lo TF32 planes.
#include "fps_common.cuh"

namespace fps {

constexpr int FOURIER_THREADS = 128;  // 100 frequency-pair lanes + 4 pad lanes, rounded to warps

// sin and cos of a moderate fp64 argument (|z| < ~1e3 here: |z| <= |x F0| + |y F1| + |beta| ~ 40), branch free:
// two-term Cody-Waite reduction by pi/2 and the fdlibm kernel polynomials on |r| <= pi/4 (~1 ulp in fp64; the
// results are rounded to fp32 afterwards).  The library sincos() with its large-argument path cost three times
// as many fp64 instructions and made this kernel compute bound.
__device__ __forceinline__ void sincos_reduced(double z, double& sn, double& cs) {
  const double kf = rint(z * 6.36619772367581382433e-01);           // z * 2/pi
  const int q = (int)kf;
  double r = fma(-kf, 1.57079632679489655800e+00, z);
  r = fma(-kf, 6.12323399573676603587e-17, r);
  const double r2 = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, r2, -2.50507602534068634195e-08);
  ps = fma(ps, r2, 2.75573137070700676789e-06);
  ps = fma(ps, r2, -1.98412698298579493134e-04);
  ps = fma(ps, r2, 8.33333333332248946124e-03);
  ps = fma(ps, r2, -1.66666666666666324348e-01);
  const double s = fma(ps * r2, r, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, r2, 2.08757232129817482790e-09);
  pc = fma(pc, r2, -2.75573143513906633035e-07);
  pc = fma(pc, r2, 2.48015872894767294178e-05);
  pc = fma(pc, r2, -1.38888888888741095749e-03);
  pc = fma(pc, r2, 4.16666666666666019037e-02);
  const double c = fma(pc * r2, r2, fma(-0.5, r2, 1.0));
  const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;           // quadrant: (s, c), (c, -s), (-s, -c), (-c, s)
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}

// two adjacent columns at once (idx even): 4-byte stores into the fp16 planes, 8-byte into the fp32 planes
__device__ __forceinline__ void store_planes2(const ActPlanes& P, size_t idx, float v0, float v1, float scale) {
  if (P.f16) {
    const float t0 = v0 * scale, t1 = v1 * scale;
    const __half2 h = __floats2half2_rn(t0, t1);
    const float2 hf = __half22float2(h);
    *reinterpret_cast<__half2*>(P.hi16 + idx) = h;
    *reinterpret_cast<__half2*>(P.lo16 + idx) = __floats2half2_rn(t0 - hf.x, t1 - hf.y);
  } else {
    const float h0 = tf32_hi(v0), h1 = tf32_hi(v1);
    *reinterpret_cast<float2*>(P.hi + idx) = make_float2(h0, h1);
    *reinterpret_cast<float2*>(P.lo + idx) = make_float2(v0 - h0, v1 - h1);
  }
}

// Thread j2 < 100 owns the frequency pair (2 j2, 2 j2 + 1); threads 100..103 zero the K padding (columns
// 400..415).  The kernel is bound by its stores (6.6 KB per point in 2-byte elements): pairing the columns
// halves the store instructions.
template <int S>
__global__ void __launch_bounds__(FOURIER_THREADS)
fourier_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
               const double* __restrict__ ffm, const ActPlanes P, const StreamScale sc, const int ldk) {
  const int j = 2 * threadIdx.x;
  double f0[2] = {0, 0}, f1[2] = {0, 0}, c[2] = {0, 0}, beta[2] = {0, 0}, k2[2];
  if (j < NFREQ) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      f0[u] = ffm[j + u];
      f1[u] = ffm[NFREQ + j + u];
      c[u] = ffm[2 * NFREQ + j + u];
      beta[u] = ffm[3 * NFREQ + j + u];
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) k2[u] = f0[u] * f0[u] + f1[u] * f1[u];
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const size_t row = (size_t)p * S * ldk;
    if (j < NFREQ) {
      const double xp = x[p], yp = y[p];
      double sn[2], cs[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        sincos_reduced(fma(xp, f0[u], fma(yp, f1[u], beta[u])), sn[u], cs[u]);
        sn[u] *= c[u];
        cs[u] *= c[u];
      }
      store_planes2(P, row + j, (float)sn[0], (float)sn[1], sc.s[0]);
      store_planes2(P, row + NFREQ + j, (float)cs[0], (float)cs[1], sc.s[0]);
      if (S == 4) {
        store_planes2(P, row + ldk + j, (float)(cs[0] * f0[0]), (float)(cs[1] * f0[1]), sc.s[1]);
        store_planes2(P, row + ldk + NFREQ + j, (float)(-sn[0] * f0[0]), (float)(-sn[1] * f0[1]), sc.s[1]);
        store_planes2(P, row + 2 * ldk + j, (float)(cs[0] * f1[0]), (float)(cs[1] * f1[1]), sc.s[2]);
        store_planes2(P, row + 2 * ldk + NFREQ + j, (float)(-sn[0] * f1[0]), (float)(-sn[1] * f1[1]), sc.s[2]);
        store_planes2(P, row + 3 * ldk + j, (float)(-k2[0] * sn[0]), (float)(-k2[1] * sn[1]), sc.s[3]);
        store_planes2(P, row + 3 * ldk + NFREQ + j, (float)(-k2[0] * cs[0]), (float)(-k2[1] * cs[1]), sc.s[3]);
      }
    } else if (j < NFREQ + 8) {
      // zero the K padding (columns 400..415) of every stream row: 4 threads x 4 columns
      const int c0 = K0 + 2 * (j - NFREQ);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        store_planes2(P, row + s * ldk + c0, 0.f, 0.f, 1.f);
        store_planes2(P, row + s * ldk + c0 + 2, 0.f, 0.f, 1.f);
      }
    }
  }
}

int launch_fourier(const fps_ctx* ctx, const double* x, const double* y, int64_t n, int streams, ActPlanes out,
                   cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t blocks = n < (int64_t)ctx->sm_count * 16 ? n : (int64_t)ctx->sm_count * 16;
  if (streams == 4)
    fourier_kernel<4><<<(unsigned)blocks, FOURIER_THREADS, 0, st>>>(x, y, n, ctx->ffm, out, ctx->act_scale[0], ctx->layer[0].ldk);
  else
    fourier_kernel<1><<<(unsigned)blocks, FOURIER_THREADS, 0, st>>>(x, y, n, ctx->ffm, out, ctx->act_scale[0], ctx->layer[0].ldk);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
