// Full fp64 feature engine (Solver(precision=torch.float64), test_solverclass.py:65-68 in the reference):
// the same Fourier layer + dense/tanh layers + forward-mode Laplacian streams as the fp32 path, but
// every value, weight, product and accumulation in double precision on the DMMA m8n8k4 tensor path.
//
// Why it exists: with default-initialised weights and rough sources |w| reaches 1e5, so u amplifies
// feature rounding by ~1e3; fp32 *storage* of H alone then costs ~1e-4 in u (precision
// budget).  The reference's `precision=torch.float64` mode has no such floor, and neither has this.
//
// Reference maths: model.py:134-146 (Fourier), model.py:53-60/:144 (Linear+Tanh), utils.py:54-68
// (Laplacian) - closed-form streams as in layer_simt.cu / layer_tc.cu.
#include "fps_common.cuh"

namespace fps {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- Fourier layer ------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(224)
fourier_f64_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                   const double* __restrict__ ffm, double* __restrict__ out) {
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
    double* row = out + (size_t)p * S * K0P;
    if (j < NFREQ) {
      const double z = fma(x[p], f0, fma(y[p], f1, beta));
      double sn, cs;
      sincos(z, &sn, &cs);
      sn *= c;
      cs *= c;
      row[j] = sn;
      row[NFREQ + j] = cs;
      if (S == 4) {
        row[K0P + j] = cs * f0;
        row[K0P + NFREQ + j] = -sn * f0;
        row[2 * K0P + j] = cs * f1;
        row[2 * K0P + NFREQ + j] = -sn * f1;
        row[3 * K0P + j] = -k2 * sn;
        row[3 * K0P + NFREQ + j] = -k2 * cs;
      }
    } else if (j < NFREQ + 8) {
      const int c0 = K0 + 2 * (j - NFREQ);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        row[s * K0P + c0] = 0.0;
        row[s * K0P + c0 + 1] = 0.0;
      }
    }
  }
}

// ---- dense layer + tanh streams ---------------------------------------------------------------
// Tile: 128 activation rows x 64 features, K steps of 16; 8 warps as 4 (rows) x 2 (features), each
// a 32 x 32 sub-tile of 4 x 4 DMMA fragments.  Inside a warp's 32-row slab the rows are permuted to
// [stream][point] order when they are staged (smem row = 8*stream + point), so that the DMMA
// accumulator fragment of one thread (row = lane/4 of each 8-row fragment) holds the four streams
// of ONE point and the tanh recurrences need no shuffles.
constexpr int DM = 128, DN = 64, DK = 16, DS = DK + 4;  // DS == 4 (mod 16): conflict-free fragment loads

template <int S, bool LAST>
__global__ void __launch_bounds__(256)
layer_f64_kernel(const double* __restrict__ in, int Kp, const double* __restrict__ W, const double* __restrict__ bias,
                 double* __restrict__ out, double* __restrict__ H, double* __restrict__ DH, int64_t ld, int64_t rows) {
  __shared__ __align__(16) double As[DM * DS];
  __shared__ __align__(16) double Ws[DN * DS];
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int wi = warp / 2, wj = warp % 2, lr = lane / 4, lc = lane % 4;
  const int64_t row0 = (int64_t)blockIdx.x * DM;
  const int col0 = blockIdx.y * DN;

  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;

  // staging: A 128 x 16 doubles = 1024 double2, 4 per thread; W 64 x 16 = 512 double2, 2 per thread
  for (int k0 = 0; k0 < Kp; k0 += DK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = q * 256 + t;
      const int r = idx / 8, kq = (idx % 8) * 2;
      double2 v = make_double2(0.0, 0.0);
      if (row0 + r < rows) v = *reinterpret_cast<const double2*>(in + (row0 + r) * Kp + k0 + kq);
      int m = r;
      if (S == 4) {  // r = slab*32 + 4*point + stream  ->  slab*32 + 8*stream + point
        const int slab = r / 32, pt = (r % 32) / 4, s = r % 4;
        m = slab * 32 + s * 8 + pt;
      }
      *reinterpret_cast<double2*>(As + m * DS + kq) = v;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = q * 256 + t;
      const int r = idx / 8, kq = (idx % 8) * 2;
      *reinterpret_cast<double2*>(Ws + r * DS + kq) =
          *reinterpret_cast<const double2*>(W + (size_t)(col0 + r) * Kp + k0 + kq);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DK / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(wi * 32 + i * 8 + lr) * DS + kk * 4 + lc];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[(wj * 32 + j * 8 + lr) * DS + kk * 4 + lc];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }

  if (S == 4) {
    // fragment i = stream, row lr = point within the slab
    const int64_t r = row0 + wi * 32 + lr * 4;  // global row of the value stream of this point
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col0 + wj * 32 + j * 8 + lc * 2;
        double o[4][2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double a = c[0][j][e] + bias[col + e], ax = c[1][j][e], ay = c[2][j][e], al = c[3][j][e];
          const double h = tanh(a);
          const double d1 = fma(-h, h, 1.0), d2 = -2.0 * h * d1;
          o[0][e] = h;
          o[1][e] = d1 * ax;
          o[2][e] = d1 * ay;
          o[3][e] = fma(d2, fma(ax, ax, ay * ay), d1 * al);
        }
        if (LAST) {
          const int64_t p = r / 4;
          *reinterpret_cast<double2*>(H + p * ld + col) = make_double2(o[0][0], o[0][1]);
          *reinterpret_cast<double2*>(DH + p * ld + col) = make_double2(o[3][0], o[3][1]);
        } else {
#pragma unroll
          for (int s = 0; s < 4; ++s)
            *reinterpret_cast<double2*>(out + (r + s) * FP + col) = make_double2(o[s][0], o[s][1]);
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = row0 + wi * 32 + i * 8 + lr;
      if (r >= rows) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col0 + wj * 32 + j * 8 + lc * 2;
        const double2 o = make_double2(tanh(c[i][j][0] + bias[col]), tanh(c[i][j][1] + bias[col + 1]));
        if (LAST) *reinterpret_cast<double2*>(H + r * ld + col) = o;
        else *reinterpret_cast<double2*>(out + r * FP + col) = o;
      }
    }
  }
}

int launch_fourier_f64(const fps_ctx* ctx, const double* x, const double* y, int64_t n, int streams, double* out,
                       cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t blocks = n < (int64_t)ctx->sm_count * 8 ? n : (int64_t)ctx->sm_count * 8;
  if (streams == 4) fourier_f64_kernel<4><<<(unsigned)blocks, 224, 0, st>>>(x, y, n, ctx->ffm, out);
  else fourier_f64_kernel<1><<<(unsigned)blocks, 224, 0, st>>>(x, y, n, ctx->ffm, out);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

int launch_layer_f64(const fps_ctx* ctx, int l, const double* in, double* out, int64_t points, int streams, bool last,
                     double* H, double* DH, int64_t ld, cudaStream_t st) {
  if (points <= 0) return 0;
  const int Kp = ctx->layer[l].Kp;
  const double* W = ctx->w64[l];
  const double* bias = ctx->b64[l];
  const int64_t rows = points * streams;
  dim3 grid((unsigned)((rows + DM - 1) / DM), FP / DN);
#define FPS_F64_LAUNCH(S_, LAST_) \
  layer_f64_kernel<S_, LAST_><<<grid, 256, 0, st>>>(in, Kp, W, bias, out, H, DH, ld, rows)
  if (streams == 4) {
    if (last) FPS_F64_LAUNCH(4, true); else FPS_F64_LAUNCH(4, false);
  } else {
    if (last) FPS_F64_LAUNCH(1, true); else FPS_F64_LAUNCH(1, false);
  }
#undef FPS_F64_LAUNCH
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
