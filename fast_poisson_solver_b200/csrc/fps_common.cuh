// Shared declarations for libfps_sm100.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fps_sm100.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfps_sm100 is written for sm_100a (B200) only"
#endif

namespace fps {

constexpr int F = FPS_F;          // 700 features
constexpr int FP = FPS_FP;        // padded feature dimension (row stride of every activation)
constexpr int FPW = 768;          // weight rows padded to 6 x 128 (UMMA M tiles); rows >= F are zero
constexpr int K0 = FPS_K0;        // 400 Fourier inputs
constexpr int K0P = 416;          // K0 padded to a multiple of 32 (one 128-byte TMA/UMMA K-chunk)
// Row pitch (elements) of the first layer's operand planes - Fourier activations and W0 alike.  416 halves are 832 bytes:
// every other row would start in the middle of a 128-byte line and each 128-byte slab row of a TMA box would touch two L2
// lines.  448 halves = 7 x 128 bytes keeps every slab row inside one line; columns 416..447 are never written nor read
// (the tensor maps end at K0P and zero-fill beyond).  FPS_K0_LD overrides (A/B measurements).
constexpr int K0LD = 448;
constexpr int NFREQ = FPS_NFREQ;  // 200
constexpr int NLAYER = 1 + FPS_NHIDDEN;  // mapping.linear + 5 hidden
constexpr int NSTREAM = 4;        // value, d/dx, d/dy, laplacian

// Activation planes.  A layer input for P points is a (rows = P*S, K) row-major matrix, row index
// r = point*S + stream (S = 4 for PDE points, 1 for boundary points), stored as two fp32 planes:
//   hi = value with the low 13 mantissa bits cleared (exactly a TF32 number)
//   lo = value - hi (exact in fp32; its top 11 bits are what the tensor core consumes)
// so that hi + lo reproduces the fp32 activation exactly and A*B ~ Ahi*Bhi + Ahi*Blo + Alo*Bhi.
//
// f16 = 1 ("FP16x2 split", the production format): the same value is stored as two fp16 planes of a
// power-of-two scaled copy t = v * S (S per layer input and stream, chosen so that typical |t| ~ 1):
//   hi16 = fp16(t),  lo16 = fp16(t - hi16)            (22+ significant bits, 4 bytes per element)
// and the contraction runs as three kind::f16 MMAs  Whi*Ahi + Whi*Alo + Wlo*Ahi  with the weights split the
// same way at scale W16_SCALE; the epilogue multiplies by 1 / (W16_SCALE * S).  Half the operand bytes of
// the TF32 format and half the tensor-core cycles, at the same ~2^-22 product accuracy (fp16 and tf32 both
// carry 11 significand bits); values far below the scale fall into fp16 subnormals, whose 6e-8 absolute
// spacing relative to |t| ~ 1 is still fp32 grade for the dot product.
struct ActPlanes {
  float* hi;
  float* lo;
  __half* hi16;
  __half* lo16;
  int f16;
};
constexpr float W16_SCALE = 16.0f;   // |w| <= 1/sqrt(400) = 0.05 -> |w * 16| < 1

struct Layer {
  float* w_hi;   // (FPW, Kp) row-major, K-major == torch weight[out][in], zero padded
  float* w_lo;
  float* bias;   // (FPW) zero padded
  __half* w_hi16;  // (FPW, Kp) fp16(w * W16_SCALE)
  __half* w_lo16;  // (FPW, Kp) fp16(w * W16_SCALE - w_hi16)
  int Kp;        // padded K (K0P for layer 0, FP otherwise)
  int ldk;       // row pitch in elements of the weight planes AND of this layer's input planes (K0LD for layer 0, else FP)
};

struct StreamScale {  // per-stream power-of-two scales of one layer's input planes (f16 format)
  float s[4];
};

// Exact-operand descriptor ("xop", FPS_XOP_BYTES of caller- or context-owned device memory): the per-column fixed-point
// grid on which one fp32 feature matrix A (n x ld) enters the int8 engine (gram_i8.cu, recon_i8.cu).
//   exps[f]   e_f with |A[:, f]| < 2^e_f          scale[f] = {2^(30 - e_f), 2^(e_f - 30)}
//   cmax[f]   max |A[:, f]| as float bits (atomicMax), recorded by the kernel that produced A
//   flags[0]  set by the fused last-layer epilogue when an element did not fit the calibrated grid: the consumers then
//             rebuild exps from cmax and re-split A (device-side conditional, no host round trip)
constexpr int XOP_ROWS = 768;
struct Xop {
  int exps[XOP_ROWS];
  float2 scale[XOP_ROWS];
  unsigned cmax[XOP_ROWS];
  int flags[4];
};
static_assert(sizeof(Xop) <= FPS_XOP_BYTES, "Xop must fit FPS_XOP_BYTES");

}  // namespace fps

namespace fps {
struct Prof;  // event-pair recorder, fps_api.cu
extern unsigned long long g_launches;  // kernels launched by this library (host-side counter)
}

struct fps_ctx {
  fps::Prof* prof;
  int device;
  int engine;
  int chunk;              // points per wave
  int f16;                // 1: FP16x2-split operand planes (production), 0: 3xTF32 planes
  fps::StreamScale act_scale[fps::NLAYER];  // input scales per layer (f16 format)
  int sm_count;
  fps::Layer layer[fps::NLAYER];
  double* ffm;            // 4*NFREQ doubles: F0 | F1 | coefficients | biases
  fps::ActPlanes act[2];  // ping-pong, each plane chunk*4*FP floats
  double* partial;        // split-K partial tiles for the fp64 GEMMs
  size_t partial_bytes;
  size_t workspace_bytes;
  void* tc_state;         // engine-private (tensor maps), owned by layer_tc.cu
  int* tc_sched;          // 2 ints for the dynamic tile scheduler of layer_tc_pair_kernel (zero between launches)
  void* gi_state;         // int8 exact-Gram workspace, owned by gram_i8.cu (allocated on first use)
  void* ri_state;         // int8 reconstruction workspace, owned by recon_i8.cu (allocated on first use)
  fps::Xop* xop_tmp;      // context-owned descriptors for the uncached entry points: [0] A operand, [1] scratch
  unsigned* dh_cal_cmax;  // FPW words: max |DH| per feature over the calibration sample (float bits); 0 = not calibrated
  int dh_cal_valid;
  int gain_checked;       // the accumulator-gain compensation of layer_tc.cu has been checked against the fp32 engine
  int gain_off;           // 1: compensation disabled for this weight set (the check found it harmful)
  double cal_err[4];      // rel-L2 vs the fp32 engine on the calibration sample: {H, DH} with gain, {H, DH} without
  double* w64[fps::NLAYER]; // fp64 engine: (FP, Kp) row-major weights, zero padded
  double* b64[fps::NLAYER]; // fp64 engine: (FP) biases
};

namespace fps {

void set_error(const char* fmt, ...);

#define FPS_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      fps::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define FPS_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      fps::set_error(__VA_ARGS__);    \
      return 2;                       \
    }                                 \
  } while (0)

// ---- kernel launchers (one translation unit each) -------------------------------------------
int launch_fourier(const fps_ctx* ctx, const double* x, const double* y, int64_t n, int streams,
                   ActPlanes out, cudaStream_t st);

// One dense layer + tanh-derivative epilogue for `points` points with `streams` streams.
// last = write H (and DH when streams == 4) instead of the next layer's planes.
int launch_layer_simt(const fps_ctx* ctx, int l, ActPlanes in, ActPlanes out, int64_t points, int streams,
                      bool last, float* H, float* DH, int64_t ld, cudaStream_t st);
// Exact-engine outputs of the LAST layer of a PDE-point wave (tcgen05 engine; all null = plain fp32 H / DH):
//   xdh      grid of DH.  Its cmax always receives max |DH| per feature.  With gplanes != null the epilogue snaps DH to
//            the grid already stored in xdh->scale (calibrated, set by the caller), writes the point-major digit planes
//            of gram_i8.cu for the points [point0, point0 + points) and raises xdh->flags[0] if an element does not fit.
//   snap_h   snap H to the uniform 2^-30 grid (|H| <= 1: no maxima needed, cannot overflow)
struct FusedOut {
  Xop* xdh;
  int8_t* gplanes;
  int64_t point0;
  int snap_h;
};
int launch_layer_tc(const fps_ctx* ctx, int l, ActPlanes in, ActPlanes out, int64_t points, int streams,
                    bool last, float* H, float* DH, int64_t ld, const FusedOut* fused, cudaStream_t st);
int tc_init(fps_ctx* ctx);
int chol_init(fps_ctx* ctx);
int sources_init(fps_ctx* ctx);
int stream_init(fps_ctx* ctx);
// streaming single / few right-hand-side kernels (stream_f32.cu); return 1 = not applicable (fall back), 0 = done
int launch_project_stream(const fps_ctx* ctx, const float* A, int64_t lda, const float* Fm, int64_t ldf, int B, int64_t n,
                          double* partial, int* nctas_out, cudaStream_t st);
int launch_recon_stream(const fps_ctx* ctx, const float* A, int64_t n, int64_t lda, const double* W, int64_t ldw,
                        const double* bias, int B, float* out, int64_t ldo, cudaStream_t st);
void tc_destroy(fps_ctx* ctx);

// C (M x Ncols, ldc) fp64 += A^T B over n rows; A (n, lda) fp32 M columns, B (n, ldb) fp32 Ncols columns.
// symmetric: A == B, only tiles with j <= i are computed and mirrored.
int launch_atb_f64(const fps_ctx* ctx, const float* A, int64_t lda, int M, const float* B, int64_t ldb, int Ncols,
                   int64_t n, double* C, int64_t ldc, bool symmetric, cudaStream_t st);
// Same Gram (symmetric A^T A, M = FP) on the int8 tensor cores with exact products (gram_i8.cu)
int launch_gram_i8(fps_ctx* ctx, float* A, int64_t lda, int64_t n, double* G, int64_t ldg, cudaStream_t st);
// cached forms: caller-owned descriptor + digit planes of all n rows (gram_i8.cu, recon_i8.cu)
size_t gram_planes_bytes(int64_t n);
size_t recon_planes_bytes(int64_t n);
int launch_gram_x(fps_ctx* ctx, float* A, int64_t lda, int64_t n, Xop* x, int8_t* planes, bool fused, double* G,
                  int64_t ldg, cudaStream_t st);
int launch_project_x(fps_ctx* ctx, const Xop* x, const int8_t* planes, int64_t n, const float* F, int64_t ldf, int B,
                     double* R, int64_t ldr, cudaStream_t st);
int launch_recon_planes_build(fps_ctx* ctx, float* A, int64_t lda, int64_t n, Xop* x, int8_t* planes, bool need_grid,
                              cudaStream_t st);
int launch_recon_x(fps_ctx* ctx, const Xop* x, const int8_t* planes, int64_t n, const double* W, int64_t ldw,
                   const double* bias, int B, float* out, int64_t ldo, cudaStream_t st);
int gram_i8_reserve(fps_ctx* ctx, int64_t n, int B);
int recon_i8_reserve(fps_ctx* ctx, int B);
// xop.cu
int xop_colmax(const fps_ctx* ctx, const float* A, int64_t lda, int64_t n, int ncols, unsigned* cmax, cudaStream_t st);
int xop_exps(const unsigned* cmax, int ncols, int rows, int headroom, int* exps, float2* scale, const int* cond,
             cudaStream_t st);
int xop_uniform(Xop* x, int e, cudaStream_t st);
int xop_snap(float* A, int64_t lda, int64_t n, int ncols, const Xop* x, cudaStream_t st);
int launch_project_i8(fps_ctx* ctx, float* DH, int64_t lda, const float* F, int64_t ldf, int B, int64_t n, double* R,
                      int64_t ldr, cudaStream_t st);
void gram_i8_destroy(fps_ctx* ctx);
// out (n x B) = A W + bias on the int8 tensor cores with exact products (recon_i8.cu); snaps A in place
int launch_recon_i8(fps_ctx* ctx, float* A, int64_t lda, int64_t n, const double* W, int64_t ldw, const double* bias,
                    int B, float* out, int64_t ldo, cudaStream_t st);
void recon_i8_destroy(fps_ctx* ctx);
int launch_colsum(const float* A, int64_t n, int64_t ld, double* s64, cudaStream_t st);
int launch_assemble_factor(const fps_ctx* ctx, const double* G, const double* Gbc, const double* s, int64_t n_total,
                           int64_t nbc_total, const double* d_counts, const double* h_lambdas, const double* h_jitter,
                           int n_lambda, double* L, int* info, cudaStream_t st);
int launch_solve_factored(const fps_ctx* ctx, const double* L, const double* R, int64_t ldr, int B, double scale,
                          const double* s, const double* sum_bc, int64_t nbc_total, double* W, double* bias,
                          cudaStream_t st);
int launch_reconstruct(const fps_ctx* ctx, const float* A, int64_t n, int64_t ld, const double* W, int64_t ldw,
                       const double* bias, int B, float* out, int64_t ldo, cudaStream_t st);
int launch_sse_columns(const fps_ctx* ctx, const float* pred, int64_t ldp, const float* ref, int64_t ldr,
                       const double* ref_const, int64_t n, int B, double* sse, cudaStream_t st);
int launch_sse_columns_d(const fps_ctx* ctx, const double* pred, int64_t ldp, const double* ref, int64_t ldr,
                         const double* ref_const, int64_t n, int B, double* sse, cudaStream_t st);

// count `n` kernel launches (instrumentation for bench.py's gpu_launches)
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// tanh layer on the four forward-mode streams (the closed form of what
// utils.py:54-68 obtains by autograd):  h = tanh a, d1 = 1-h^2, d2 = -2 h d1,
// h_x = d1 a_x, h_y = d1 a_y, h_lap = d2 (a_x^2 + a_y^2) + d1 a_lap.
__device__ __forceinline__ void tanh_streams(float a, float ax, float ay, float al, float& h, float& hx, float& hy,
                                             float& hl) {
  h = tanhf(a);
  const float d1 = fmaf(-h, h, 1.0f);
  const float d2 = -2.0f * h * d1;
  hx = d1 * ax;
  hy = d1 * ay;
  hl = fmaf(d2, fmaf(ax, ax, ay * ay), d1 * al);
}

// store one activation value into the operand planes of the next layer (see ActPlanes)
__device__ __forceinline__ void store_planes(const ActPlanes& P, size_t idx, float v, float scale) {
  if (P.f16) {
    const float t = v * scale;
    const __half h = __float2half_rn(t);
    P.hi16[idx] = h;
    P.lo16[idx] = __float2half_rn(t - __half2float(h));
  } else {
    const float h = tf32_hi(v);
    P.hi[idx] = h;
    P.lo[idx] = v - h;
  }
}
// same with the format known at compile time (tensor-core epilogue)
template <bool F16>
__device__ __forceinline__ void store_planes_t(const ActPlanes& P, size_t idx, float v, float scale) {
  if (F16) {
    const float t = v * scale;
    const __half h = __float2half_rn(t);
    P.hi16[idx] = h;
    P.lo16[idx] = __float2half_rn(t - __half2float(h));
  } else {
    const float h = tf32_hi(v);
    P.hi[idx] = h;
    P.lo[idx] = v - h;
  }
}
// read it back (SIMT cross-check engine)
__device__ __forceinline__ float load_planes(const ActPlanes& P, size_t idx, float inv_scale) {
  if (P.f16) return (__half2float(P.hi16[idx]) + __half2float(P.lo16[idx])) * inv_scale;
  return P.hi[idx] + P.lo[idx];
}

// Balanced radix-256 digits of a 31-bit integer X = d0 2^24 + d1 2^16 + d2 2^8 + d3 with d1..d3 in [-128, 127]
// (|X| <= 2^30, so |d0| <= 64), all four at once: adding 128 at the three lower byte positions turns every balanced
// digit into its offset-binary byte, and flipping those sign bits back gives the two's complement bytes
// [d0 | d1 | d2 | d3] (d0 in the top byte).
__device__ __forceinline__ uint32_t balanced_digits4(int X) {
  return ((uint32_t)X + 0x00808080u) ^ 0x00808080u;
}
// 4 x 4 byte transpose: z[u] = digits word of element u -> w[s] = digit s of the elements 0..3 (element 0 in the low byte)
__device__ __forceinline__ void digits_transpose4(uint32_t z0, uint32_t z1, uint32_t z2, uint32_t z3, uint32_t* w) {
  const uint32_t t0 = __byte_perm(z0, z1, 0x5140), t1 = __byte_perm(z0, z1, 0x7362);
  const uint32_t t2 = __byte_perm(z2, z3, 0x5140), t3 = __byte_perm(z2, z3, 0x7362);
  w[3] = __byte_perm(t0, t2, 0x5410);
  w[2] = __byte_perm(t0, t2, 0x7632);
  w[1] = __byte_perm(t1, t3, 0x5410);
  w[0] = __byte_perm(t1, t3, 0x7632);
}

__device__ __forceinline__ void split_store(float* hi, float* lo, float v) {
  float h = tf32_hi(v);
  *hi = h;
  *lo = v - h;
}

}  // namespace fps

namespace fps {
// ---- fp64 engine (precision = float64) --------------------------------------------------------
int launch_fourier_f64(const fps_ctx* ctx, const double* x, const double* y, int64_t n, int streams, double* out,
                       cudaStream_t st);
int launch_layer_f64(const fps_ctx* ctx, int l, const double* in, double* out, int64_t points, int streams, bool last,
                     double* H, double* DH, int64_t ld, cudaStream_t st);
int launch_atb_f64_d(const fps_ctx* ctx, const double* A, int64_t lda, int M, const double* B, int64_t ldb, int Ncols,
                     int64_t n, double* C, int64_t ldc, bool symmetric, cudaStream_t st);
int launch_colsum_d(const double* A, int64_t n, int64_t ld, double* s64, cudaStream_t st);
int launch_reconstruct_d(const fps_ctx* ctx, const double* A, int64_t n, int64_t ld, const double* W, int64_t ldw,
                         const double* bias, int B, double* out, int64_t ldo, cudaStream_t st);
}  // namespace fps
