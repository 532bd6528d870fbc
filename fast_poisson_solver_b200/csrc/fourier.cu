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

constexpr int FOURIER_THREADS = 224;  // 200 frequency lanes + 8 pad lanes, rounded to warps

template <int S>
__global__ void __launch_bounds__(FOURIER_THREADS)
fourier_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
               const double* __restrict__ ffm, const ActPlanes P, const StreamScale sc) {
  const int j = threadIdx.x;
  double f0 = 0, f1 = 0, c = 0, beta = 0;
  if (j < NFREQ) {
    f0 = ffm[j];
    f1 = ffm[NFREQ + j];
    c = ffm[2 * NFREQ + j];
    beta = ffm[3 * NFREQ + j];
  }
  const double k2 = f0 * f0 + f1 * f1;
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const size_t row = (size_t)p * S * K0P;
    if (j < NFREQ) {
      const double z = fma(x[p], f0, fma(y[p], f1, beta));
      double sn, cs;
      sincos(z, &sn, &cs);
      sn *= c;
      cs *= c;
      store_planes(P, row + j, (float)sn, sc.s[0]);
      store_planes(P, row + NFREQ + j, (float)cs, sc.s[0]);
      if (S == 4) {
        store_planes(P, row + K0P + j, (float)(cs * f0), sc.s[1]);
        store_planes(P, row + K0P + NFREQ + j, (float)(-sn * f0), sc.s[1]);
        store_planes(P, row + 2 * K0P + j, (float)(cs * f1), sc.s[2]);
        store_planes(P, row + 2 * K0P + NFREQ + j, (float)(-sn * f1), sc.s[2]);
        store_planes(P, row + 3 * K0P + j, (float)(-k2 * sn), sc.s[3]);
        store_planes(P, row + 3 * K0P + NFREQ + j, (float)(-k2 * cs), sc.s[3]);
      }
    } else if (j < NFREQ + 8) {
      // zero the K padding (columns 400..415) of every stream row
      const int c0 = K0 + 2 * (j - NFREQ);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        store_planes(P, row + s * K0P + c0, 0.f, 1.f);
        store_planes(P, row + s * K0P + c0 + 1, 0.f, 1.f);
      }
    }
  }
}

int launch_fourier(const fps_ctx* ctx, const double* x, const double* y, int64_t n, int streams, ActPlanes out,
                   cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t blocks = n < (int64_t)ctx->sm_count * 8 ? n : (int64_t)ctx->sm_count * 8;
  if (streams == 4)
    fourier_kernel<4><<<(unsigned)blocks, FOURIER_THREADS, 0, st>>>(x, y, n, ctx->ffm, out, ctx->act_scale[0]);
  else
    fourier_kernel<1><<<(unsigned)blocks, FOURIER_THREADS, 0, st>>>(x, y, n, ctx->ffm, out, ctx->act_scale[0]);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
