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
  // ---- inverses of the diagonal blocks (rows 704..735 of the factor buffer), one warp per block, from Dfac; they also
  // seed the diagonal of the explicit inverse below
  double* Xf = L + (size_t)(CN + CB) * CN;   // (CN x CN): lower triangle X = L^-1, upper triangle X^T
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
      for (int i = 0; i < CB; ++i) {
        L[(size_t)(CN + i) * CN + kb * CB + j] = x[i];
        __stcg(Xf + (size_t)(kb * CB + i) * CN + kb * CB + j, x[i]);
      }
    }
  }
  // ---- explicit inverse X = L^-1 (applied by the few-right-hand-side solve as GEMVs with one refinement step against
  // L itself, instead of 44 dependent substitution steps).  Recursive merging of diagonal segments: for a pair of
  // adjacent segments [a0, a1), [a1, a2) whose inverses X11, X22 are known,
  //     X21 = -X22 (L21 X11)
  // - two rounds of independent 64 x 64 tile products per level, five levels for the 22 blocks (22 -> 11 -> 6 -> 3 -> 2
  // -> 1 segments).  T = L21 X11 is parked transposed in the (still unused) upper triangle of Xf.
  __shared__ int segb[CNB + 1];
  int nseg = CNB;
  if (t <= CNB) segb[t] = t * CB;
  cf_grid_barrier(bar, target);
  while (nseg > 1) {
    const int npair = nseg / 2;
    for (int phase = 0; phase < 2; ++phase) {
      // tiles of all pairs of this level, dealt round-robin to the CTAs
      int tile = blockIdx.x;
      for (int p = 0; p < npair; ++p) {
        const int a0 = segb[2 * p], a1 = segb[2 * p + 1], a2 = segb[2 * p + 2];
        const int tr = (a2 - a1 + 63) / 64, tc = (a1 - a0 + 63) / 64;
        for (; tile < tr * tc; tile += gridDim.x) {
          const int r0 = a1 + (tile / tc) * 64, c0 = a0 + (tile % tc) * 64;
          // phase 0: T[r][c] = sum_{k = c .. a1-1} L[r][k] X11[k][c]      (X11 lower triangular)
          // phase 1: X21[r][c] = - sum_{k = a1 .. r} X22[r][k] T[k][c]    (X22 lower triangular; T read from the mirror)
          const int kbeg = phase == 0 ? (c0 / CB) * CB : a1;
          const int kend = phase == 0 ? a1 : min(a2, r0 + 64);
          const int tx = t % 16, ty = t / 16;
          double acc[4][4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
          for (int k0 = kbeg; k0 < kend; k0 += CB) {
            __syncthreads();
            for (int e = t; e < 64 * CB; e += 256) {
              const int rr = e / CB, kk = e % CB;          // Pi[row][k]: the left operand, read along k
              const int r = r0 + rr, k = k0 + kk;
              double v = 0.0;
              if (r < a2 && k < kend) {
                if (phase == 0) v = __ldcg(L + (size_t)r * CN + k);
                else if (k <= r) v = __ldcg(Xf + (size_t)r * CN + k);
              }
              Pi[rr][kk] = v;
            }
            for (int e = t; e < 64 * CB; e += 256) {
              // Pj[col][k]: the right operand; phase 0 reads rows of X11 (columns contiguous), phase 1 rows of T^T (k contiguous)
              const int kk = phase == 0 ? e / 64 : e % CB, cc = phase == 0 ? e % 64 : e / CB;
              const int c = c0 + cc, k = k0 + kk;
              double v = 0.0;
              if (c < a1 && k < kend) {
                if (phase == 0) { if (k >= c) v = __ldcg(Xf + (size_t)k * CN + c); }
                else v = __ldcg(Xf + (size_t)c * CN + k);    // T[k][c], stored transposed
              }
              Pj[cc][kk] = v;
            }
            __syncthreads();
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
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int r = r0 + ty + 16 * a, c = c0 + tx + 16 * b;
              if (r < a2 && c < a1) {
                if (phase == 0) __stcg(Xf + (size_t)c * CN + r, acc[a][b]);
                else __stcg(Xf + (size_t)r * CN + c, -acc[a][b]);
              }
            }
        }
        tile -= tr * tc;
      }
      cf_grid_barrier(bar, target);
    }
    // next level: every pair becomes one segment, an odd last segment is carried over
    __syncthreads();
    const int nnew = npair + (nseg & 1);
    int keep = 0;
    if (t <= nnew) keep = (t < npair) ? segb[2 * t] : (t == nnew ? CN : segb[2 * npair]);
    __syncthreads();
    if (t <= nnew) segb[t] = keep;
    __syncthreads();
    nseg = nnew;
  }
  // ---- mirror: upper triangle of Xf = X^T (the transposed products of the backward sweep read rows, too)
  for (int i = blockIdx.x; i < CN; i += gridDim.x)
    for (int j = t; j < i; j += 256) __stcg(Xf + (size_t)j * CN + i, __ldcg(Xf + (size_t)i * CN + j));
}

int launch_assemble_factor(const fps_ctx* ctx, const double* G, const double* Gbc, const double* s, int64_t n_total,
                           int64_t nbc_total, const double* d_counts, const double* h_lambdas, const double* h_jitter,
                           int n_lambda, double* L, int* info, cudaStream_t st) {
  FPS_REQUIRE(d_counts || (n_total > 0 && nbc_total > 0), "assemble: empty problem (N=%lld, N_bc=%lld)", (long long)n_total,
              (long long)nbc_total);
  // scratch in the partial workspace: [0] = max diag, then the 22 factored diagonal blocks
  double* scal = ctx->partial;
  double* Dfac = ctx->partial + 8;
  FPS_REQUIRE(ctx->sm_count >= CF_CTAS, "assemble_factor: %d SMs < %d co-resident CTAs", ctx->sm_count, CF_CTAS);
  for (int li = 0; li < n_lambda; ++li) {
    double* Ll = L + (size_t)li * FPS_LROWS * CN;
    // scal[0] = max diagonal (as ordered bits), scal[1] = barrier counter
    FPS_CUDA(cudaMemsetAsync(scal, 0, 2 * sizeof(double), st));
    // cooperative launch: the 66 CTAs wait on one another at the grid barriers, so the driver must place them all at once
    double lam = h_lambdas[li], nt = (double)n_total, nbt = (double)nbc_total, jit = h_jitter ? h_jitter[li] : 0.0;
    unsigned long long* dmax = reinterpret_cast<unsigned long long*>(scal);
    unsigned* bar = reinterpret_cast<unsigned*>(scal + 1);
    int* info_l = info + li;
    void* args[] = {(void*)&G, (void*)&Gbc, (void*)&s, &lam, &nt, &nbt, (void*)&d_counts, &jit, &Ll, &Dfac, &dmax, &bar, &info_l};
    FPS_CUDA(cudaLaunchCooperativeKernel((const void*)chol_fused_kernel, dim3(CF_CTAS), dim3(256), args, 0, st));
  }
  count_launch(n_lambda);
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

// ---- few right-hand sides (B <= 8): the solve as GEMVs with the explicit inverse --------------------------------------
// out[i][c] = beta base[i][c] + alpha sum_k M[i][k] v[k][c] over the lower (k <= i) or upper (k >= i) part of row i of a
// (CN x CN) matrix; one warp per row, v staged in shared memory.  All 704 rows run in parallel on 88 CTAs; the
// substitution kernel above needs 44 dependent block steps on ONE SM (0.6 ms for a single right-hand side).
template <bool UPPER>
__global__ void __launch_bounds__(256)
tri_gemv_kernel(const double* __restrict__ M, const double* __restrict__ v, int64_t ldv, double alpha,
                const double* __restrict__ base, int64_t ldb, double beta, double* __restrict__ out, int64_t ldo, int B) {
  __shared__ double vs[CN * SC];
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  for (int e = t; e < CN * SC; e += 256) {
    const int k = e / SC, c = e % SC;
    vs[e] = (c < B) ? __ldcg(v + (int64_t)k * ldv + c) : 0.0;
  }
  __syncthreads();
  const int i = blockIdx.x * 8 + warp;
  if (i >= CN) return;
  const int k0 = UPPER ? i : 0, k1 = UPPER ? CN : i + 1;
  double acc[SC];
#pragma unroll
  for (int c = 0; c < SC; ++c) acc[c] = 0.0;
  const double* row = M + (size_t)i * CN;
  for (int k = (k0 & ~31) + lane; k < k1; k += 32) {
    if (k < k0) continue;
    const double m = __ldcg(row + k);
#pragma unroll
    for (int c = 0; c < SC; ++c) acc[c] = fma(m, vs[k * SC + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < SC; ++c)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  if (lane < B) {
    double r = 0.0;
#pragma unroll
    for (int c = 0; c < SC; ++c) r = (lane == c) ? acc[c] : r;
    r *= alpha;
    if (base) r = fma(beta, __ldcg(base + (int64_t)i * ldb + lane), r);
    out[(int64_t)i * ldo + lane] = r;
  }
}

// bias_c = -(s^T W_c - sum_bc_c) / N_bc  (solver.py:306), one warp per column
__global__ void bias_kernel(const double* __restrict__ W, int64_t ldw, int B, const double* __restrict__ s,
                            const double* __restrict__ sum_bc, double inv_nbc, double* __restrict__ bias) {
  const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (c >= B) return;
  double acc = 0.0;
  for (int r = lane; r < F; r += 32) acc = fma(s[r], W[(int64_t)r * ldw + c], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) bias[c] = -(acc - sum_bc[c]) * inv_nbc;
}

int launch_solve_factored(const fps_ctx* ctx, const double* L, const double* R, int64_t ldr, int B, double scale,
                          const double* s, const double* sum_bc, int64_t nbc_total, double* W, double* bias,
                          cudaStream_t st) {
  if (B <= 0) return 0;
  static const int gemv_env = getenv("FPS_SOLVE_GEMV") ? atoi(getenv("FPS_SOLVE_GEMV")) : 1;
  if (B <= SC && gemv_env) {
    // y0 = X b, r = b - L y0, y = y0 + X r (one refinement step against L restores the backward stability of
    // substitution: the explicit inverse alone leaves a residual ~ cond(L) eps);  then the same with the transposes.
    // b = scale R.  Temporaries (5 x 704 x 8 doubles) live behind the factorisation's scratch in the split-K workspace.
    const double* X = L + (size_t)(CN + CB) * CN;
    double* tmp = ctx->partial + 8 + (size_t)CNB * CB * CB;
    double *y0 = tmp, *r = tmp + CN * SC, *y = tmp + 2 * CN * SC, *w0 = tmp + 3 * CN * SC, *r2 = tmp + 4 * CN * SC;
    const int grid = CN / 8;
    tri_gemv_kernel<false><<<grid, 256, 0, st>>>(X, R, ldr, scale, nullptr, 0, 0.0, y0, SC, B);
    tri_gemv_kernel<false><<<grid, 256, 0, st>>>(L, y0, SC, -1.0, R, ldr, scale, r, SC, B);
    tri_gemv_kernel<false><<<grid, 256, 0, st>>>(X, r, SC, 1.0, y0, SC, 1.0, y, SC, B);
    tri_gemv_kernel<true><<<grid, 256, 0, st>>>(X, y, SC, 1.0, nullptr, 0, 0.0, w0, SC, B);
    tri_gemv_kernel<true><<<grid, 256, 0, st>>>(L, w0, SC, -1.0, y, SC, 1.0, r2, SC, B);
    tri_gemv_kernel<true><<<grid, 256, 0, st>>>(X, r2, SC, 1.0, w0, SC, 1.0, W, ldr, B);
    bias_kernel<<<1, 256, 0, st>>>(W, ldr, B, s, sum_bc, 1.0 / (double)nbc_total, bias);
    count_launch(7);
    FPS_CUDA(cudaGetLastError());
    return 0;
  }
  const size_t smem = SOLVE_SMEM;
  solve_factored_kernel<<<(B + SC - 1) / SC, 256, smem, st>>>(L, R, ldr, B, scale, s, sum_bc, 1.0 / (double)nbc_total,
                                                              W, bias);
  count_launch();
  FPS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fps
