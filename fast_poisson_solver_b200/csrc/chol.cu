// Normal-equation assembly, blocked FP64 Cholesky and batched triangular solves on the 704x704
// feature-dimension system.
//
// Reference: precompute_LHS_RHS (solver.py:187-199) builds, per lambda,
//     LHS = lambda/N * DH^T DH + 1/N_bc * H_bc^T (I - 11^T/N_bc) H_bc
// with a dense N_bc x N_bc centering matrix; solve() then calls torch.linalg.solve (LU) per
// right-hand side (solver.py:305) and computes bias = -(1^T H_bc w - sum(bc)) / N_bc (:306).
// Here the centering matrix is folded in as the rank-1 correction  Gbc - s s^T / N_bc  with
// s = H_bc^T 1, the matrix is factored ONCE (it is symmetric positive semi-definite by
// construction) and every solve is two triangular sweeps over all right-hand sides at once.
//
// Conditioning: cond(LHS) ~ 1e15 in fp64 and the matrix is exactly singular when
// N + N_bc < 700 (the reference's own test fixture, grid_num=5).  fps_assemble_factor therefore
// takes a diagonal shift `jitter` (absolute) chosen by the host: it starts at 0 and is raised only
// when the factorisation reports a pivot <= eps*max_diag through `info`.
#include "fps_common.cuh"
#include <float.h>
#include <stdlib.h>

namespace fps {

constexpr int CN = FP;   // matrix order incl. padding
constexpr int CB = 32;   // panel width
constexpr int CNB = CN / CB;  // 22 panels

// ---- assemble ----------------------------------------------------------------------------------
// counts != null: {N, N_bc} are read from device memory (the all-reduced packed buffer) and lam_over_n holds lambda
__global__ void assemble_kernel(const double* __restrict__ G, const double* __restrict__ Gbc,
                                const double* __restrict__ s, double lam_over_n, double inv_nbc, double jitter,
                                const double* __restrict__ counts, double* __restrict__ L) {
  const int i = blockIdx.x;
  if (counts) {
    lam_over_n = lam_over_n / counts[0];
    inv_nbc = 1.0 / counts[1];
  }
  for (int j = threadIdx.x; j < CN; j += blockDim.x) {
    double v;
    if (i < F && j < F) {
      v = lam_over_n * G[(size_t)i * CN + j] + (Gbc[(size_t)i * CN + j] - s[i] * s[j] * inv_nbc) * inv_nbc;
      if (i == j) v += jitter;
    } else {
      v = (i == j) ? 1.0 : 0.0;
    }
    L[(size_t)i * CN + j] = v;
  }
}

// max diagonal of the assembled matrix (first F entries) -> scal[0]; resets info.
__global__ void diagmax_kernel(const double* __restrict__ L, double* __restrict__ scal, int* __restrict__ info) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < F; i += blockDim.x) m = fmax(m, L[(size_t)i * CN + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)blockDim.x / 32; ++w) m = fmax(m, red[w]);
    scal[0] = m;
    *info = 0;
  }
}

// ---- panel: factor the 32x32 diagonal block (every CTA, redundantly) and solve the rows below ----
// L21 = A21 * L11^-T.  The factored diagonal block goes to Dfac (CTA 0 only) so that CTAs that are
// still reading the unfactored block from L never race with the write-back; fixup_kernel installs
// the diagonal blocks at the end.
__global__ void __launch_bounds__(256)
chol_panel_kernel(double* __restrict__ L, int kb, double* __restrict__ Dfac, const double* __restrict__ scal,
                  int* __restrict__ info) {
  __shared__ double D[CB][CB + 1];
  const int k0 = kb * CB;
  const int t = threadIdx.x;
  for (int e = t; e < CB * CB; e += 256) D[e / CB][e % CB] = L[(size_t)(k0 + e / CB) * CN + k0 + e % CB];
  const double tol = DBL_EPSILON * scal[0];
  const double safe = scal[0] > 0.0 ? scal[0] : 1.0;
  for (int j = 0; j < CB; ++j) {
    __syncthreads();
    double d = D[j][j];
    const bool bad = !(d > tol);
    if (bad) d = safe;  // keep the arithmetic finite; the host sees info != 0 and retries with a shift
    const double ljj = sqrt(d);
    __syncthreads();
    if (t < CB) {
      if (t == j) {
        D[j][j] = ljj;
        if (bad && blockIdx.x == 0) atomicCAS(info, 0, k0 + j + 1);
      } else if (t > j) {
        D[t][j] = D[t][j] / ljj;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = t + 256 * q;
      const int i = e / CB, m = e % CB;
      if (m > j && i >= m) D[i][m] -= D[i][j] * D[m][j];
    }
  }
  __syncthreads();
  if (blockIdx.x == 0)
    for (int e = t; e < CB * CB; e += 256) Dfac[(size_t)kb * CB * CB + e] = (e % CB <= e / CB) ? D[e / CB][e % CB] : 0.0;

  const int r = k0 + CB + blockIdx.x * 256 + t;
  if (r < CN) {
    double x[CB];
    double* row = L + (size_t)r * CN + k0;
#pragma unroll
    for (int j = 0; j < CB; ++j) x[j] = row[j];
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      double sacc = x[j];
#pragma unroll
      for (int m = 0; m < j; ++m) sacc = fma(-x[m], D[j][m], sacc);
      x[j] = sacc / D[j][j];
    }
#pragma unroll
    for (int j = 0; j < CB; ++j) row[j] = x[j];
  }
}

// ---- trailing update: A22 -= L21 L21^T on lower-triangular 64x64 tiles -----------------------------
__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ L, int kb) {
  __shared__ double Pi[64][CB + 1];
  __shared__ double Pj[64][CB + 1];
  const int base = (kb + 1) * CB;
  int tile = blockIdx.x, ti, tj;
  ti = (int)((sqrtf(8.f * tile + 1.f) - 1.f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
  while (ti * (ti + 1) / 2 > tile) --ti;
  tj = tile - ti * (ti + 1) / 2;
  const int i0 = base + ti * 64, j0 = base + tj * 64;
  const int k0 = kb * CB;
  for (int e = threadIdx.x; e < 64 * CB; e += 256) {
    const int r = e / CB, c = e % CB;
    Pi[r][c] = (i0 + r < CN) ? L[(size_t)(i0 + r) * CN + k0 + c] : 0.0;
    Pj[r][c] = (j0 + r < CN) ? L[(size_t)(j0 + r) * CN + k0 + c] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
  for (int k = 0; k < CB; ++k) {
    double pa[4], pb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) pa[a] = Pi[ty + 16 * a][k];
#pragma unroll
    for (int b = 0; b < 4; ++b) pb[b] = Pj[tx + 16 * b][k];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(pa[a], pb[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < CN && j <= i) L[(size_t)i * CN + j] -= acc[a][b];
    }
}

// ---- fixup: install factored diagonal blocks, mirror L into the upper triangle (so that both
// triangular sweeps read rows contiguously) --------------------------------------------------------
__global__ void chol_fixup_kernel(double* __restrict__ L, const double* __restrict__ Dfac) {
  const int i = blockIdx.x;
  const int kb = i / CB;
  for (int j = threadIdx.x; j <= i; j += blockDim.x) {
    double v;
    if (j >= kb * CB) v = Dfac[(size_t)kb * CB * CB + (i - kb * CB) * CB + (j - kb * CB)];
    else v = L[(size_t)i * CN + j];
    L[(size_t)i * CN + j] = v;
    if (j < i) L[(size_t)j * CN + i] = v;
  }
}

// Inverses of the 22 diagonal 32x32 blocks of L, stored in rows 704..735 of the factor buffer (block kb in
// columns 32 kb .. 32 kb + 31).  The batched solve applies them as small mat-vecs instead of 32 sequential
// substitution steps; one step of iterative refinement against the block itself keeps the block solve
// backward stable.  One warp per block, lane j builds column j by forward substitution on e_j.
__global__ void chol_dinv_kernel(double* __restrict__ L) {
  __shared__ double D[CB][CB + 1];
  const int kb = blockIdx.x, k0 = kb * CB, j = threadIdx.x;
  for (int e = threadIdx.x; e < CB * CB; e += 32) D[e / CB][e % CB] = L[(size_t)(k0 + e / CB) * CN + k0 + e % CB];
  __syncwarp();
  double x[CB];
#pragma unroll
  for (int i = 0; i < CB; ++i) {
    double sacc = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int m = 0; m < i; ++m) sacc = fma(-D[i][m], x[m], sacc);
    x[i] = (i >= j) ? sacc / D[i][i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < CB; ++i) L[(size_t)(CN + i) * CN + k0 + j] = x[i];
}

// ---- the whole factorisation in ONE launch -------------------------------------------------------------------------
// assemble -> max diagonal -> 22 x (panel, trailing update) -> fixup + diagonal-block inverses, with a grid-wide barrier
// between the phases instead of 48 kernel launches (the 704 x 704 problem is 0.11 GFLOP: pure latency).  CF_CTAS CTAs,
// one per SM and all co-resident (the barrier spins): 66 = the tile count of the first trailing update.  Data that
// crosses CTAs goes through L2 (__ldcg / __stcg); the barrier is a monotone counter in global memory (zeroed by the
// host before the launch), bounded spin + trap like the mbarrier waits of the tensor-core kernels.
constexpr int CF_CTAS = 66;

__device__ __forceinline__ void cf_grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) {
      if (++spins > 20000000u) {
        printf("fps chol_fused: grid barrier timeout (block %d target %u)\n", blockIdx.x, target);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1)
chol_fused_kernel(const double* __restrict__ G, const double* __restrict__ Gbc, const double* __restrict__ sv,
                  double lam, double n_total, double nbc_total, const double* __restrict__ counts, double jitter,
                  double* __restrict__ L, double* __restrict__ Dfac, unsigned long long* __restrict__ dmax_bits,
                  unsigned* __restrict__ bar, int* __restrict__ info) {
  __shared__ double D[CB][CB + 1];
  __shared__ double Dr[CB];            // reciprocals of the diagonal of the factored block
  __shared__ double Pi[64][CB + 1];
  __shared__ double Pj[64][CB + 1];
  __shared__ double red[8];
  const int t = threadIdx.x;
  unsigned target = 0;
  if (counts) { n_total = counts[0]; nbc_total = counts[1]; }
  const double lam_over_n = lam / n_total, inv_nbc = 1.0 / nbc_total;

  // ---- assemble (solver.py:195-196 with the centering matrix folded in) + max diagonal
  double m = 0.0;
  for (int i = blockIdx.x; i < CN; i += gridDim.x) {
    for (int j = t; j < CN; j += 256) {
      double v;
      if (i < F && j < F) {
        v = lam_over_n * G[(size_t)i * CN + j] + (Gbc[(size_t)i * CN + j] - sv[i] * sv[j] * inv_nbc) * inv_nbc;
        if (i == j) { v += jitter; m = fmax(m, v); }
      } else {
        v = (i == j) ? 1.0 : 0.0;
      }
      __stcg(L + (size_t)i * CN + j, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (t % 32 == 0) red[t / 32] = m;
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
    atomicMax(dmax_bits, (unsigned long long)__double_as_longlong(m));   // m >= 0: the bit patterns order like the values
    if (blockIdx.x == 0) *info = 0;
  }
  cf_grid_barrier(bar, target);
  const double dmax = __longlong_as_double((long long)__ldcg(dmax_bits));
  const double tol = DBL_EPSILON * dmax;
  const double safe = dmax > 0.0 ? dmax : 1.0;

  for (int kb = 0; kb < CNB; ++kb) {
    const int k0 = kb * CB;
    const int below = CN - (k0 + CB);
    // ---- panel: CTAs that own rows below (256 per CTA) factor the diagonal block redundantly, then solve their rows
    const int my_r = k0 + CB + blockIdx.x * 256 + t;
    const bool cta_has_rows = blockIdx.x * 256 < below || blockIdx.x == 0;
    if (cta_has_rows) {
      for (int e = t; e < CB * CB; e += 256) D[e / CB][e % CB] = __ldcg(L + (size_t)(k0 + e / CB) * CN + k0 + e % CB);
      for (int j = 0; j < CB; ++j) {
        __syncthreads();
        double d = D[j][j];
        const bool bad = !(d > tol);
        if (bad) d = safe;  // keep the arithmetic finite; the host sees info != 0 and retries with a shift
        const double rs = rsqrt(d);          // the serial chain of 704 columns: one rsqrt + multiplies, no divisions
        __syncthreads();
        if (t < CB) {
          if (t == j) {
            D[j][j] = d * rs;
            Dr[j] = rs;
            if (bad && blockIdx.x == 0) atomicCAS(info, 0, k0 + j + 1);
          } else if (t > j) {
            D[t][j] = D[t][j] * rs;
          }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = t + 256 * q;
          const int i = e / CB, mm = e % CB;
          if (mm > j && i >= mm) D[i][mm] -= D[i][j] * D[mm][j];
        }
      }
      __syncthreads();
      if (blockIdx.x == 0)
        for (int e = t; e < CB * CB; e += 256) Dfac[(size_t)kb * CB * CB + e] = (e % CB <= e / CB) ? D[e / CB][e % CB] : 0.0;
      if (my_r < CN) {
        double x[CB];
        double* row = L + (size_t)my_r * CN + k0;
#pragma unroll
        for (int j = 0; j < CB; ++j) x[j] = __ldcg(row + j);
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          double sacc = x[j];
#pragma unroll
          for (int mm = 0; mm < j; ++mm) sacc = fma(-x[mm], D[j][mm], sacc);
          x[j] = sacc * Dr[j];
        }
#pragma unroll
        for (int j = 0; j < CB; ++j) __stcg(row + j, x[j]);
      }
    }
    if (below <= 0) break;
    cf_grid_barrier(bar, target);
    // ---- trailing update A22 -= L21 L21^T on lower-triangular 64 x 64 tiles
    const int base = k0 + CB;
    const int t64 = (below + 63) / 64;
    const int ntile = t64 * (t64 + 1) / 2;
    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
      int ti = (int)((sqrtf(8.f * tile + 1.f) - 1.f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
      while (ti * (ti + 1) / 2 > tile) --ti;
      const int tj = tile - ti * (ti + 1) / 2;
      const int i0 = base + ti * 64, j0 = base + tj * 64;
      __syncthreads();
      for (int e = t; e < 64 * CB; e += 256) {
        const int r = e / CB, c = e % CB;
        Pi[r][c] = (i0 + r < CN) ? __ldcg(L + (size_t)(i0 + r) * CN + k0 + c) : 0.0;
        Pj[r][c] = (j0 + r < CN) ? __ldcg(L + (size_t)(j0 + r) * CN + k0 + c) : 0.0;
      }
      __syncthreads();
      const int tx = t % 16, ty = t / 16;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
      for (int k = 0; k < CB; ++k) {
        double pa[4], pb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) pa[a] = Pi[ty + 16 * a][k];
#pragma unroll
        for (int b = 0; b < 4; ++b) pb[b] = Pj[tx + 16 * b][k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fma(pa[a], pb[b], acc[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
          if (i < CN && j <= i) {
            double* q = L + (size_t)i * CN + j;
            __stcg(q, __ldcg(q) - acc[a][b]);
          }
        }
    }
    cf_grid_barrier(bar, target);
  }
  cf_grid_barrier(bar, target);
  // ---- fixup: install the factored diagonal blocks, mirror L into the upper triangle
  for (int i = blockIdx.x; i < CN; i += gridDim.x) {
    const int kb = i / CB;
    for (int j = t; j <= i; j += 256) {
      double v;
      if (j >= kb * CB) v = __ldcg(Dfac + (size_t)kb * CB * CB + (i - kb * CB) * CB + (j - kb * CB));
      else v = __ldcg(L + (size_t)i * CN + j);
      __stcg(L + (size_t)i * CN + j, v);
      if (j < i) __stcg(L + (size_t)j * CN + i, v);
    }
  }
  // ---- inverses of the diagonal blocks (rows 704..735 of the factor buffer), one warp per block, from Dfac
  for (int kb = blockIdx.x; kb < CNB; kb += gridDim.x) {
    __syncthreads();
    for (int e = t; e < CB * CB; e += 256) D[e / CB][e % CB] = __ldcg(Dfac + (size_t)kb * CB * CB + e);
    __syncthreads();
    if (t < CB) {
      const int j = t;
      double x[CB];
#pragma unroll
      for (int i = 0; i < CB; ++i) {
        double sacc = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int mm = 0; mm < i; ++mm) sacc = fma(-D[i][mm], x[mm], sacc);
        x[i] = (i >= j) ? sacc / D[i][i] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < CB; ++i) L[(size_t)(CN + i) * CN + kb * CB + j] = x[i];
    }
  }
}

int launch_assemble_factor(const fps_ctx* ctx, const double* G, const double* Gbc, const double* s, int64_t n_total,
                           int64_t nbc_total, const double* d_counts, const double* h_lambdas, const double* h_jitter,
                           int n_lambda, double* L, int* info, cudaStream_t st) {
  FPS_REQUIRE(d_counts || (n_total > 0 && nbc_total > 0), "assemble: empty problem (N=%lld, N_bc=%lld)", (long long)n_total,
              (long long)nbc_total);
  // scratch in the partial workspace: [0] = max diag, then the 22 factored diagonal blocks
  double* scal = ctx->partial;
  double* Dfac = ctx->partial + 8;
  static const int fused_env = getenv("FPS_CHOL_FUSED") ? atoi(getenv("FPS_CHOL_FUSED")) : 1;
  if (fused_env && ctx->sm_count >= CF_CTAS) {
    for (int li = 0; li < n_lambda; ++li) {
      double* Ll = L + (size_t)li * FPS_LROWS * CN;
      // scal[0] = max diagonal (as ordered bits), scal[1] = barrier counter
      FPS_CUDA(cudaMemsetAsync(scal, 0, 2 * sizeof(double), st));
      chol_fused_kernel<<<CF_CTAS, 256, 0, st>>>(G, Gbc, s, h_lambdas[li], (double)n_total, (double)nbc_total, d_counts,
                                                 h_jitter ? h_jitter[li] : 0.0, Ll, Dfac,
                                                 reinterpret_cast<unsigned long long*>(scal),
                                                 reinterpret_cast<unsigned*>(scal + 1), info + li);
    }
    count_launch(n_lambda);
    FPS_CUDA(cudaGetLastError());
    return 0;
  }
  for (int li = 0; li < n_lambda; ++li) {
    double* Ll = L + (size_t)li * FPS_LROWS * CN;
    const double lam = h_lambdas[li];
    const double jitter = h_jitter ? h_jitter[li] : 0.0;
    if (d_counts) assemble_kernel<<<CN, 256, 0, st>>>(G, Gbc, s, lam, 0.0, jitter, d_counts, Ll);
    else assemble_kernel<<<CN, 256, 0, st>>>(G, Gbc, s, lam / (double)n_total, 1.0 / (double)nbc_total, jitter, nullptr, Ll);
    diagmax_kernel<<<1, 256, 0, st>>>(Ll, scal, info + li);
    for (int kb = 0; kb < CNB; ++kb) {
      const int below = CN - (kb + 1) * CB;
      const int pc = below > 0 ? (below + 255) / 256 : 1;
      chol_panel_kernel<<<pc, 256, 0, st>>>(Ll, kb, Dfac, scal, info + li);
      if (below > 0) {
        const int t64 = (below + 63) / 64;
        chol_update_kernel<<<t64 * (t64 + 1) / 2, 256, 0, st>>>(Ll, kb);
      }
    }
    chol_fixup_kernel<<<CN, 256, 0, st>>>(Ll, Dfac);
    chol_dinv_kernel<<<CNB, 32, 0, st>>>(Ll);
  }
  count_launch(n_lambda * (4 + CNB + (CNB - 1)));
  FPS_CUDA(cudaGetLastError());
  return 0;
}

// ---- batched solve  W = (L L^T)^-1 (scale R),  bias = -(s^T W - sum_bc)/N_bc ----------------------
// One CTA owns SC right-hand-side columns and runs both sweeps entirely in shared memory; L (full,
// mirrored) is streamed from L2.  Warp c solves the 32x32 diagonal systems of column c with lane = row.
constexpr int SC = 8;

__global__ void __launch_bounds__(256)
solve_factored_kernel(const double* __restrict__ L, const double* __restrict__ R, int64_t ldr, int B, double scale,
                      const double* __restrict__ s, const double* __restrict__ sum_bc, double inv_nbc,
                      double* __restrict__ W, double* __restrict__ bias) {
  extern __shared__ double sm[];
  double* X = sm;                  // [CN][SC]
  double* D = sm + CN * SC;        // [CB][CB+1] diagonal block (lower triangle, zeros above)
  double* Di = D + CB * (CB + 1);  // [CB][CB+1] its inverse
  double* xb = Di + CB * (CB + 1); // [CB][SC] solved block
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int c0 = blockIdx.x * SC;
  // The sweeps are a chain of 44 dependent steps that each wait for a block of L.  Between the factorisation and
  // this kernel the projection has streamed gigabytes through L2, so L (4 MB) would come from HBM one dependent
  // miss at a time: pull it into L2 up front (a few CTAs' worth of prefetches; harmless when many CTAs run).
  if (blockIdx.x < 4) {
    const char* Lb = reinterpret_cast<const char*>(L);
    const size_t bytes = (size_t)(CN + CB) * CN * sizeof(double);
    for (size_t off = ((size_t)blockIdx.x * 256 + t) * 128; off < bytes; off += (size_t)min(gridDim.x, 4u) * 256 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Lb + off));
  }
  for (int e = t; e < CN * SC; e += 256) {
    const int r = e / SC, c = e % SC;
    X[e] = (c0 + c < B) ? scale * R[(int64_t)r * ldr + c0 + c] : 0.0;
  }
  // forward: L y = b
  for (int kb = 0; kb < CNB; ++kb) {
    const int k0 = kb * CB;
    __syncthreads();
    for (int e = t; e < CB * CB; e += 256) {
      const int i = e / CB, j = e % CB;
      D[i * (CB + 1) + j] = (j <= i) ? L[(size_t)(k0 + i) * CN + k0 + j] : 0.0;
      Di[i * (CB + 1) + j] = L[(size_t)(CN + i) * CN + k0 + j];
    }
    __syncthreads();
    {
      // x = Dinv b, then one refinement step x += Dinv (b - D x); lane = row, warp = right-hand side
      const double b = X[(k0 + lane) * SC + warp];
      double x = 0.0;
#pragma unroll 8
      for (int j = 0; j < CB; ++j) x = fma(Di[lane * (CB + 1) + j], __shfl_sync(0xffffffffu, b, j), x);
      double r = b;
#pragma unroll 8
      for (int j = 0; j < CB; ++j) r = fma(-D[lane * (CB + 1) + j], __shfl_sync(0xffffffffu, x, j), r);
#pragma unroll 8
      for (int j = 0; j < CB; ++j) x = fma(Di[lane * (CB + 1) + j], __shfl_sync(0xffffffffu, r, j), x);
      X[(k0 + lane) * SC + warp] = x;
      xb[lane * SC + warp] = x;
    }
    __syncthreads();
    // rows below: X[r] -= sum_m L[r][k0+m] xb[m]  (read through the mirrored upper triangle: coalesced)
    for (int r = k0 + CB + t; r < CN; r += 256) {
      // all 32 L2 loads of the panel row in flight at once (the sweep is latency bound)
      double l[CB];
#pragma unroll
      for (int m = 0; m < CB; ++m) l[m] = __ldg(L + (size_t)(k0 + m) * CN + r);
      double acc[SC];
#pragma unroll
      for (int c = 0; c < SC; ++c) acc[c] = X[r * SC + c];
#pragma unroll
      for (int m = 0; m < CB; ++m) {
#pragma unroll
        for (int c = 0; c < SC; ++c) acc[c] = fma(-l[m], xb[m * SC + c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < SC; ++c) X[r * SC + c] = acc[c];
    }
  }
  // backward: L^T w = y
  for (int kb = CNB - 1; kb >= 0; --kb) {
    const int k0 = kb * CB;
    __syncthreads();
    for (int e = t; e < CB * CB; e += 256) {
      const int i = e / CB, j = e % CB;
      D[i * (CB + 1) + j] = (j <= i) ? L[(size_t)(k0 + i) * CN + k0 + j] : 0.0;
      Di[i * (CB + 1) + j] = L[(size_t)(CN + i) * CN + k0 + j];
    }
    __syncthreads();
    {
      // transposed block: x = Dinv^T b, refined against D^T
      const double b = X[(k0 + lane) * SC + warp];
      double x = 0.0;
#pragma unroll 8
      for (int j = 0; j < CB; ++j) x = fma(Di[j * (CB + 1) + lane], __shfl_sync(0xffffffffu, b, j), x);
      double r = b;
#pragma unroll 8
      for (int j = 0; j < CB; ++j) r = fma(-D[j * (CB + 1) + lane], __shfl_sync(0xffffffffu, x, j), r);
#pragma unroll 8
      for (int j = 0; j < CB; ++j) x = fma(Di[j * (CB + 1) + lane], __shfl_sync(0xffffffffu, r, j), x);
      X[(k0 + lane) * SC + warp] = x;
      xb[lane * SC + warp] = x;
    }
    __syncthreads();
    // rows above: X[r] -= sum_m L[k0+m][r] xb[m]
    for (int r = t; r < k0; r += 256) {
      // all 32 L2 loads of the panel row in flight at once (the sweep is latency bound)
      double l[CB];
#pragma unroll
      for (int m = 0; m < CB; ++m) l[m] = __ldg(L + (size_t)(k0 + m) * CN + r);
      double acc[SC];
#pragma unroll
      for (int c = 0; c < SC; ++c) acc[c] = X[r * SC + c];
#pragma unroll
      for (int m = 0; m < CB; ++m) {
#pragma unroll
        for (int c = 0; c < SC; ++c) acc[c] = fma(-l[m], xb[m * SC + c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < SC; ++c) X[r * SC + c] = acc[c];
    }
  }
  __syncthreads();
  for (int e = t; e < CN * SC; e += 256) {
    const int r = e / SC, c = e % SC;
    if (c0 + c < B) W[(int64_t)r * ldr + c0 + c] = X[e];
  }
  // bias (solver.py:306): warp c reduces s^T w over the F real features
  if (c0 + warp < B) {
    double acc = 0.0;
    for (int r = lane; r < F; r += 32) acc = fma(s[r], X[r * SC + warp], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) bias[c0 + warp] = -(acc - sum_bc[c0 + warp]) * inv_nbc;
  }
}

constexpr size_t SOLVE_SMEM = (size_t)(CN * SC + 2 * CB * (CB + 1) + CB * SC) * sizeof(double);

// per context (= per device): the function attribute is a per-device setting
int chol_init(fps_ctx* ctx) {
  (void)ctx;
  FPS_CUDA(cudaFuncSetAttribute(solve_factored_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM));
  return 0;
}

int launch_solve_factored(const double* L, const double* R, int64_t ldr, int B, double scale, const double* s,
                          const double* sum_bc, int64_t nbc_total, double* W, double* bias, cudaStream_t st) {
  if (B <= 0) return 0;
  const size_t smem = SOLVE_SMEM;
  solve_factored_kernel<<<(B + SC - 1) / SC, 256, smem, st>>>(L, R, ldr, B, scale, s, sum_bc, 1.0 / (double)nbc_total,
                                                              W, bias);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
