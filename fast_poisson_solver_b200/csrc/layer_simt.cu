// Dense layer + tanh with forward-mode derivative streams, fp32 FFMA engine (FPS_ENGINE_SIMT).
// This is the bring-up / cross-check engine: same buffers, same layout and same epilogue maths as
// the tcgen05 engine in layer_tc.cu, but the contraction runs on the FP32 pipe so that it can
// serve as an on-device reference for the tensor-core kernel at sizes the CPU oracle cannot reach.
//
// Reference: nn.Linear + nn.Tanh blocks of model.py:53-60 / model.py:144, with the derivative
// recurrences that utils.py:54-68 gets from autograd written in closed form:
//   a = W v + b ; a_x = W v_x ; a_y = W v_y ; a_l = W v_l
//   h = tanh a ; d1 = 1 - h^2 ; d2 = -2 h d1
//   h_x = d1 a_x ; h_y = d1 a_y ; h_l = d2 (a_x^2 + a_y^2) + d1 a_l
#include "fps_common.cuh"

namespace fps {

constexpr int SBM = 128, SBN = 64, SBK = 16, STHREADS = 256;

template <int S, bool LAST>
__global__ void __launch_bounds__(STHREADS)
layer_simt_kernel(const ActPlanes in, const StreamScale in_sc, int Kp, int ldk,
                  const float* __restrict__ w_hi, const float* __restrict__ w_lo, const float* __restrict__ bias,
                  const ActPlanes outp, const StreamScale out_sc, float* __restrict__ H,
                  float* __restrict__ DH, int64_t ld, int64_t rows) {
  const float* __restrict__ in_hi = in.hi;
  const float* __restrict__ in_lo = in.lo;
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Bs[SBK][SBN + 4];
  const int t = threadIdx.x;
  const int tx = t % 16, ty = t / 16;
  const int64_t row0 = (int64_t)blockIdx.x * SBM;
  const int col0 = blockIdx.y * SBN;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = t / 2, a_k = (t % 2) * 8;
  const int b_row = t / 4, b_k = (t % 4) * 4;
  const bool a_ok = row0 + a_row < rows;
  const float* a_hi = in_hi + (row0 + a_row) * ldk + a_k;
  const float* a_lo = in_lo + (row0 + a_row) * ldk + a_k;
  const float* b_hi = w_hi + (size_t)(col0 + b_row) * ldk + b_k;
  const float* b_lo = w_lo + (size_t)(col0 + b_row) * ldk + b_k;

  for (int k0 = 0; k0 < Kp; k0 += SBK) {
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
    if (a_ok && in.f16) {
      // FP16x2-split planes: value = (hi16 + lo16) / scale[stream of this row]
      const size_t base = (size_t)(row0 + a_row) * ldk + a_k + k0;
      const float inv = 1.0f / in_sc.s[S == 4 ? (int)((row0 + a_row) & 3) : 0];
      const uint4 qh = *reinterpret_cast<const uint4*>(in.hi16 + base);
      const uint4 ql = *reinterpret_cast<const uint4*>(in.lo16 + base);
      const __half2* hh = reinterpret_cast<const __half2*>(&qh);
      const __half2* hl = reinterpret_cast<const __half2*>(&ql);
      float v[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fh = __half22float2(hh[e]), fl = __half22float2(hl[e]);
        v[2 * e] = (fh.x + fl.x) * inv;
        v[2 * e + 1] = (fh.y + fl.y) * inv;
      }
      a0 = make_float4(v[0], v[1], v[2], v[3]);
      a1 = make_float4(v[4], v[5], v[6], v[7]);
    } else if (a_ok) {
      const float4 h0 = *reinterpret_cast<const float4*>(a_hi + k0);
      const float4 h1 = *reinterpret_cast<const float4*>(a_hi + k0 + 4);
      const float4 l0 = *reinterpret_cast<const float4*>(a_lo + k0);
      const float4 l1 = *reinterpret_cast<const float4*>(a_lo + k0 + 4);
      a0 = make_float4(h0.x + l0.x, h0.y + l0.y, h0.z + l0.z, h0.w + l0.w);
      a1 = make_float4(h1.x + l1.x, h1.y + l1.y, h1.z + l1.z, h1.w + l1.w);
    }
    const float4 bh = *reinterpret_cast<const float4*>(b_hi + k0);
    const float4 bl = *reinterpret_cast<const float4*>(b_lo + k0);
    __syncthreads();
    As[a_k + 0][a_row] = a0.x; As[a_k + 1][a_row] = a0.y; As[a_k + 2][a_row] = a0.z; As[a_k + 3][a_row] = a0.w;
    As[a_k + 4][a_row] = a1.x; As[a_k + 5][a_row] = a1.y; As[a_k + 6][a_row] = a1.z; As[a_k + 7][a_row] = a1.w;
    Bs[b_k + 0][b_row] = bh.x + bl.x; Bs[b_k + 1][b_row] = bh.y + bl.y;
    Bs[b_k + 2][b_row] = bh.z + bl.z; Bs[b_k + 3][b_row] = bh.w + bl.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 x0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }

  const int col = col0 + tx * 4;
  const float4 bv = *reinterpret_cast<const float4*>(bias + col);
  const float bj[4] = {bv.x, bv.y, bv.z, bv.w};

  if (S == 4) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t r = row0 + ty * 8 + q * 4;  // row of the value stream of this point
      if (r >= rows) continue;
      float o[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        tanh_streams(acc[q * 4 + 0][j] + bj[j], acc[q * 4 + 1][j], acc[q * 4 + 2][j], acc[q * 4 + 3][j], o[0][j],
                     o[1][j], o[2][j], o[3][j]);
      if (LAST) {
        const int64_t p = r / 4;
        *reinterpret_cast<float4*>(H + p * ld + col) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
        *reinterpret_cast<float4*>(DH + p * ld + col) = make_float4(o[3][0], o[3][1], o[3][2], o[3][3]);
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
          for (int j = 0; j < 4; ++j) store_planes(outp, (size_t)(r + s) * FP + col + j, o[s][j], out_sc.s[s]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = row0 + ty * 8 + i;
      if (r >= rows) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = tanhf(acc[i][j] + bj[j]);
      if (LAST) {
        *reinterpret_cast<float4*>(H + r * ld + col) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) store_planes(outp, (size_t)r * FP + col + j, o[j], out_sc.s[0]);
      }
    }
  }
}

int launch_layer_simt(const fps_ctx* ctx, int l, ActPlanes in, ActPlanes out, int64_t points, int streams, bool last,
                      float* H, float* DH, int64_t ld, cudaStream_t st) {
  if (points <= 0) return 0;
  const Layer& L = ctx->layer[l];
  const int64_t rows = points * streams;
  dim3 grid((unsigned)((rows + SBM - 1) / SBM), FP / SBN);
#define FPS_SIMT_LAUNCH(S_, LAST_)                                                                           \
  layer_simt_kernel<S_, LAST_><<<grid, STHREADS, 0, st>>>(in, ctx->act_scale[l], L.Kp, L.ldk, L.w_hi, L.w_lo, L.bias, out, \
                                                           ctx->act_scale[l + 1 < NLAYER ? l + 1 : l], H, DH, ld, rows)
  if (streams == 4) {
    if (last) FPS_SIMT_LAUNCH(4, true); else FPS_SIMT_LAUNCH(4, false);
  } else {
    if (last) FPS_SIMT_LAUNCH(1, true); else FPS_SIMT_LAUNCH(1, false);
  }
#undef FPS_SIMT_LAUNCH
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
