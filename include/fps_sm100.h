/* fps_sm100.h  --  C ABI of libfps_sm100.so, the B200 (sm_100a) implementation of the
 * fast-poisson-solver hot path:  Solver.precompute(x_pde, y_pde, x_bc, y_bc) + Solver.solve(f, bc).
 *
 * The reference (matthiasnwt/fast-poisson-solver) is pure Python and has no FFI; every entry point
 * below replaces a span of reference Python that a maintainer would swap for a ctypes call
 * (see INTEGRATION.md for the binding).  Citations are file:line in the reference checkout,
 * relative to fast_poisson_solver/.
 *
 * Conventions
 *   - every data pointer is a caller-owned DEVICE pointer unless its name starts with "h_";
 *   - matrices are row-major with an explicit leading dimension in elements;
 *   - FPS_F = 700 features, padded to FPS_FP = 704 in every feature-dimension buffer; the pad
 *     columns of H / DH / H_bc are written as zero by the library;
 *   - stream is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless stated;
 *   - return value 0 = ok, non-zero = error; fps_last_error() gives the message (per thread);
 *   - one host thread per context and ONE stream at a time per context (the tile scheduler counters and the
 *     workspaces below are not shared-safe);
 *   - allocation: a context owns its packed weights, the activation workspace of one point wave, and grow-only
 *     scratch of the exact int8 engine (digit planes of the right-hand sides F and of the weights W, split-K partial
 *     tiles; the uncached entry points fps_gram_accum / fps_project_rhs / fps_reconstruct also keep the digit planes of
 *     up to 2^20 rows of their matrix there).  Growing such a buffer calls cudaFree + cudaMalloc, which synchronises
 *     the device: call fps_reserve_workspace(ctx, n, B) once per problem size and nothing allocates afterwards.
 *     fps_workspace_bytes() reports the current total.  The cached entry points (suffix _x) keep their large operands
 *     in CALLER-owned buffers.
 */
#ifndef FPS_SM100_H
#define FPS_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPS_F 700        /* width of the feature network, resources/infos.yaml:859           */
#define FPS_FP 704       /* FPS_F padded to a multiple of 64                                  */
#define FPS_K0 400       /* 2 * ffm.num Fourier inputs, resources/infos.yaml:843-848          */
#define FPS_NFREQ 200
#define FPS_LROWS 1440   /* rows of one factor buffer: 704 of L + 32 of diagonal-block inverses + 704 of L^-1 */
#define FPS_NHIDDEN 5    /* depth-1 Linear(700,700)+Tanh blocks, model.py:49-60               */

#define FPS_ENGINE_SIMT 0     /* fp32 FFMA layer kernel (bring-up / cross-check engine)        */
#define FPS_ENGINE_TCGEN05 1  /* tcgen05/TMEM/TMA layer kernel, fp16 hi/lo operand planes (production engine) */
#define FPS_XOP_BYTES 16384   /* size of one exact-operand descriptor (see "cached exact-engine operands")  */

typedef struct fps_ctx fps_ctx;

/* Host-side view of the reference state_dict (model.py:124-130, :53; key names in SURVEY 2 #5).
 * All pointers are HOST float32, row-major, torch nn.Linear convention weight[out][in].        */
typedef struct fps_weights {
  const float* h_freqs;        /* mapping.freqs         (2, 200)                               */
  const float* h_coefficients; /* mapping.coefficients  (200)                                  */
  const float* h_biases;       /* mapping.biases        (200)                                  */
  const float* h_w0;           /* mapping.linear.weight (700, 400)                             */
  const float* h_b0;           /* mapping.linear.bias   (700)                                  */
  const float* h_wh[FPS_NHIDDEN]; /* model.{1,4,7,10,13}.weight (700, 700)                     */
  const float* h_bh[FPS_NHIDDEN]; /* model.{1,4,7,10,13}.bias   (700)                          */
} fps_weights;

const char* fps_last_error(void);
int fps_version(void);

/* Replaces Solver.build_model (solver.py:137-151): uploads + packs the weights on `device`,
 * allocates the activation workspace for `chunk_points` collocation points per launch wave
 * (0 = default).  engine = FPS_ENGINE_*.                                                       */
int fps_create(fps_ctx** out, int device, const fps_weights* w, int chunk_points, int engine);
int fps_destroy(fps_ctx* ctx);
int fps_set_engine(fps_ctx* ctx, int engine);
/* The tcgen05 engine stores layer inputs as fp16 hi/lo planes of a power-of-two scaled copy.
 * fps_calibrate_scales runs up to 512 of the given points through the fp32 cross-check engine, measures
 * max |v| per layer input and stream (value, d/dx, d/dy, laplacian) and sets every scale to the largest
 * power of two with max * scale <= 64 (1000x headroom below the fp16 maximum).  Synchronises the stream.
 * fps_get_scales copies the 6 x 4 scales out (host float[24]).                                       */
int fps_calibrate_scales(fps_ctx* ctx, const double* x, const double* y, int64_t n, void* stream);
int fps_get_scales(const fps_ctx* ctx, float* h_out24);
/* Restore a calibration (a saved precompute state must be evaluated with the operand scales and the compensation
 * setting it was built with): scales from fps_get_scales, gain_on = element 4 of fps_get_calibration (-1: leave).    */
int fps_set_calibration(fps_ctx* ctx, const float* h_in24, int gain_on);
/* The same call (a) records max |DH| per feature over the sample, from which fps_features_x derives the fixed-point grid
 * of its fused digit split, and (b) once per context checks the tensor-core kernel's accumulator-gain compensation
 * (csrc/layer_tc.cu) against the fp32 FFMA engine on the sample, switching it off for weight sets where it hurts.
 * fps_get_calibration: h_out5 = {rel-L2 of H, DH vs the fp32 engine with the compensation, the same without,
 * 1 = compensation in use / 0 = switched off / -1 = not checked yet}.                                               */
int fps_get_calibration(const fps_ctx* ctx, double* h_out5);
/* The calibrated maxima of |DH| per feature (host float[768], rows 704.. are zero; all zero = not calibrated).  Together
 * with fps_get_scales / fps_set_calibration this is the whole calibration state: a caller that shards the points over
 * several contexts (GPUs) combines the states (minimum of the scales, maximum of the maxima) and sets the result in
 * every context, so that a point's features do not depend on which shard it fell into (fast_poisson_solver_b200/
 * solver.py does this with one small all-reduce per precompute).  Both calls synchronise the device.                 */
int fps_get_dh_calibration(const fps_ctx* ctx, float* h_out768);
int fps_set_dh_calibration(fps_ctx* ctx, const float* h_in768);
int fps_get_engine(const fps_ctx* ctx);
size_t fps_workspace_bytes(const fps_ctx* ctx);

/* Replaces PINN.h (model.py:74-79, :134-146) + calculate_laplace (utils.py:54-68) as called from
 * Solver.evaluate_network_pde / evaluate_network_bc (solver.py:162-178).
 * x, y: n float64 coordinates.  H: (n, ld) float32 out.  DH: (n, ld) float32 out or NULL
 * (NULL = value-only evaluation for boundary points, solver.py:174-175).  ld >= FPS_FP.        */
int fps_features(fps_ctx* ctx, const double* x, const double* y, int64_t n,
                 float* H, float* DH, int64_t ld, void* stream);

/* Replaces `DH.t() @ DH` (solver.py:167).  G64 (FPS_FP, FPS_FP) float64 += A^T A over the n rows
 * of A (n, ld) float32; both triangles are written.  Accumulating (beta = 1) so that callers can
 * chunk or shard rows; zero G64 first.
 * For n >= 16384 the product runs on the int8 tensor cores with exact integer arithmetic
 * (csrc/gram_i8.cu).  That path SNAPS A IN PLACE to a per-column 31-bit fixed-point grid: elements
 * below 2^-7 of their column maximum move by less than 2^-31 of that maximum (under 1 % of one
 * fp32 ulp of the maximum), all others are unchanged, and G64 receives the Gram of the stored A
 * to fp64 rounding.  Call it before anything else consumes A (fps_project_rhs, fps_reconstruct).
 * The column maxima are taken in a pass over A inside the call (no state is kept between calls).  */
int fps_gram_accum(fps_ctx* ctx, float* A, int64_t n, int64_t ld, double* G64, void* stream);

/* Replaces `Ht_bc @ ones` (solver.py:176-178): s64 (FPS_FP) float64 += column sums of A.        */
int fps_colsum_accum(fps_ctx* ctx, const float* A, int64_t n, int64_t ld, double* s64, void* stream);

/* Replaces precompute_LHS_RHS (solver.py:187-199) and the LU inside torch.linalg.solve
 * (solver.py:305):  for every lambda
 *     LHS = lambda/N * G + ( Gbc - s s^T / N_bc ) / N_bc          (centering matrix folded in)
 * then the lower Cholesky factor L of LHS is left in L64[(i), :FPS_FP, :] of the (n_lambda, FPS_LROWS,
 * FPS_FP) float64 buffer (mirrored into the upper triangle); rows FPS_FP .. FPS_FP + 31 hold the inverses of the 22
 * diagonal 32x32 blocks that the many-right-hand-side solve applies, rows FPS_FP + 32 .. the explicit inverse L^-1 (lower
 * triangle; its transpose above the diagonal) that the solve for <= 8 right-hand sides applies as GEMVs with one step
 * of refinement against L.  h_lambdas / h_jitter are HOST arrays of
 * n_lambda entries; h_jitter (may be NULL = 0) is an absolute diagonal shift added to LHS - the
 * matrix has cond ~1e15 and is exactly singular when N + N_bc < 700, so the caller starts at 0 and
 * raises the shift only when info reports a breakdown.  info (device int[n_lambda]) receives 0 or
 * the 1-based index of the first pivot <= eps * max_diag.                                        */
int fps_assemble_factor(fps_ctx* ctx, const double* G64, const double* Gbc64, const double* s64,
                        int64_t n_total, int64_t nbc_total, const double* h_lambdas,
                        const double* h_jitter, int n_lambda, double* L64, int* info, void* stream);

/* The same on the packed normal-equation buffer that the path's one all-reduce carries (FPS_PACK_DOUBLES float64:
 * [G (FP*FP) | Gbc (FP*FP) | s (FP) | N | N_bc]): the point counts are read on the device, so a multi-GPU caller needs
 * no host round trip between the all-reduce and the factorisation.                                               */
#define FPS_PACK_DOUBLES (2 * FPS_FP * FPS_FP + FPS_FP + 2)
int fps_assemble_factor_packed(fps_ctx* ctx, const double* pack64, const double* h_lambdas, const double* h_jitter,
                               int n_lambda, double* L64, int* info, void* stream);

/* Replaces `RHSs @ f` (solver.py:197 + :301) for B right-hand sides at once:
 * R64 (FPS_FP, ldr) float64 += DH^T F,  DH (n, ld) float32,  F (n, ldf) float32, B columns.
 * Accumulating so that row shards / chunks can be summed; zero R64 first.
 * For n >= 16384 and B >= 32 this runs on the int8 tensor cores with exact integer arithmetic like
 * fps_gram_accum, on the same per-column grid of DH (DH is snapped in place - a no-op after
 * fps_gram_accum) and on a private copy of F on a 31-bit grid per right-hand side.              */
int fps_project_rhs(fps_ctx* ctx, float* DH, int64_t n, int64_t ld, const float* F, int64_t ldf,
                    int B, double* R64, int64_t ldr, void* stream);

/* Replaces torch.linalg.solve + the bias formula (solver.py:305-306) for B right-hand sides:
 *     W = (L L^T)^-1 (scale * R),   bias_j = -( s^T W_j - sum_bc_j ) / N_bc
 * L64 = one (FPS_LROWS, FPS_FP) factor buffer; R64 (FPS_FP, ldr) float64 in, W64 (FPS_FP, ldr) float64 out
 * (may alias R64), sum_bc64 (B),
 * bias64 (B).  scale = lambda / N.                                                             */
int fps_solve_factored(fps_ctx* ctx, const double* L64, const double* R64, int64_t ldr, int B,
                       double scale, const double* s64, const double* sum_bc64, int64_t nbc_total,
                       double* W64, double* bias64, void* stream);

/* Replaces the three GEMVs of solver.py:327-330 for B right-hand sides:
 *     out[i, j] = sum_k A[i, k] W[k, j] + (bias ? bias[j] : 0)       i < n, j < B
 * A (n, ld) float32 (H, DH or H_bc), W64 (FPS_FP, ldw) float64, out (n, ldo) float32.
 * For n >= 16384 and B >= 32 this runs on the int8 tensor cores with exact products (csrc/recon_i8.cu):
 * A is snapped in place to the same per-column 31-bit grid as in fps_gram_accum (a no-op for a DH that
 * went through it), W is used on a 39-bit grid per right-hand side.                             */
int fps_reconstruct(fps_ctx* ctx, float* A, int64_t n, int64_t ld, const double* W64, int64_t ldw,
                    const double* bias64, int B, float* out, int64_t ldo, void* stream);

/* Multi-lambda losses of solver.py:312-314 for B right-hand sides at once: sse64[j] += sum_i (pred[i][j] - ref[i][j])^2,
 * j < B; pred (n, ldp), ref (n, ldr) float32.  ref_const != NULL: ref[i][j] = ref_const[j] (float64; constant boundary
 * values) and ref / ldr are ignored.  Accumulating (zero sse64 first); fixed summation order.                          */
int fps_sse_columns(fps_ctx* ctx, const float* pred, int64_t ldp, const float* ref, int64_t ldr,
                    const double* ref_const, int64_t n, int B, double* sse64, void* stream);
int fps_sse_columns_f64(fps_ctx* ctx, const double* pred, int64_t ldp, const double* ref, int64_t ldr,
                        const double* ref_const, int64_t n, int B, double* sse64, void* stream);

/* ---- cached exact-engine operands (n >= fps_exact_min_rows(); precision = float32) --------------------------------
 * Gram, batched projection and batched reconstruction run on the int8 tensor cores with exact integer products
 * (csrc/gram_i8.cu, csrc/recon_i8.cu).  Their operands are fixed-point images of H / DH: an exact-operand descriptor
 * ("xop": FPS_XOP_BYTES of caller-owned device memory holding the per-feature exponents, scales and maxima) plus
 * digit planes, also caller-owned:
 *   gram planes   point-major  [128-point slab][4 digits][768][128]   fps_gram_planes_bytes(n)   Gram + projection
 *   recon planes  feature-major [4 digits][n padded to 64][704]       fps_recon_planes_bytes(n)  reconstruction
 * Keeping them between calls is what makes thousands of solves on one domain cheap; a caller short of memory uses
 * the uncached entry points above instead (same results bit for bit, the planes are rebuilt per call).
 *
 * fps_features_x = fps_features for PDE points, plus: H is snapped to the uniform 2^-30 grid (|H| <= 1), xop_h / xop_dh
 * are initialised, max |DH| per feature is recorded, and - tcgen05 engine, gram_planes != NULL, after
 * fps_calibrate_scales - the last layer's epilogue itself snaps DH to the calibrated grid and writes DH's gram planes
 * (*h_planes_written = 1).  An element beyond the calibrated grid raises a device flag; fps_gram_accum_x then rebuilds
 * the grid from the true maxima and re-splits DH (device-side conditional, no host synchronisation).
 * After fps_features_x + fps_gram_accum_x, H and DH never change again: solve() does not mutate precomputed state.  */
int64_t fps_exact_min_rows(void);
size_t fps_gram_planes_bytes(int64_t n);
size_t fps_recon_planes_bytes(int64_t n);
int fps_reserve_workspace(fps_ctx* ctx, int64_t n, int B);
int fps_features_x(fps_ctx* ctx, const double* x, const double* y, int64_t n, float* H, float* DH, int64_t ld,
                   void* xop_h, void* xop_dh, void* gram_planes, int* h_planes_written, void* stream);
/* Replaces `DH.t() @ DH` (solver.py:167): G64 += A^T A exactly, leaving A's gram planes in `planes` and A snapped to
 * the grid in xop.  planes_written = the value fps_features_x returned for this matrix (0: split here).              */
int fps_gram_accum_x(fps_ctx* ctx, float* A, int64_t n, int64_t ld, void* xop, void* planes, int planes_written,
                     double* G64, void* stream);
/* Replaces `RHSs @ f` (solver.py:301) for B right-hand sides from DH's cached gram planes: R64 += DH^T F.            */
int fps_project_rhs_x(fps_ctx* ctx, const void* xop, const void* planes, int64_t n, const float* F, int64_t ldf,
                      int B, double* R64, int64_t ldr, void* stream);
/* Recon planes of A (H or DH) on the grid of xop; derive_grid != 0 computes that grid from A's column maxima first
 * (a matrix that did not come from fps_features_x).  A is snapped in place (a no-op for an already snapped matrix).  */
int fps_recon_planes_build(fps_ctx* ctx, float* A, int64_t n, int64_t ld, void* xop, int derive_grid, void* planes,
                           void* stream);
/* Replaces the GEMVs of solver.py:327-330 for B right-hand sides from cached recon planes: out = A W + bias.          */
int fps_reconstruct_x(fps_ctx* ctx, const void* xop, const void* planes, int64_t n, const double* W64, int64_t ldw,
                      const double* bias64, int B, float* out, int64_t ldo, void* stream);

/* ---- precision = torch.float64 (tests/test_solverclass.py:65-68 in the reference) ---------------
 * Same contracts as the fp32 entry points above with every feature matrix, right-hand side and
 * output in float64; features are evaluated entirely in fp64 on the DMMA tensor path.             */
int fps_features_f64(fps_ctx* ctx, const double* x, const double* y, int64_t n,
                     double* H, double* DH, int64_t ld, void* stream);
int fps_gram_accum_f64(fps_ctx* ctx, const double* A, int64_t n, int64_t ld, double* G64, void* stream);
int fps_colsum_accum_f64(fps_ctx* ctx, const double* A, int64_t n, int64_t ld, double* s64, void* stream);
int fps_project_rhs_f64(fps_ctx* ctx, const double* DH, int64_t n, int64_t ld, const double* F, int64_t ldf,
                        int B, double* R64, int64_t ldr, void* stream);
int fps_reconstruct_f64(fps_ctx* ctx, const double* A, int64_t n, int64_t ld, const double* W64, int64_t ldw,
                        const double* bias64, int B, double* out, int64_t ldo, void* stream);

/* ---- on-device synthetic sources (data_utils/source_functions.py:24-73, the 'sin' family) ---------
 * F[r, j] = sum_t s * sin(k1*pi*(x_r - o1)) * (use_cos ? cos : sin)(k2*pi*(y_r - o2)),  j < B, t < n_terms
 * params: device (B, n_terms, 6) float64 = {k1, o1, k2, o2, s, use_cos}; terms with s == 0 are skipped.
 * Writes the (n, B) right-hand-side block that fps_project_rhs consumes; x, y are device float64.   */
int fps_sine_sources(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params,
                     int n_terms, int B, float* F, int64_t ldf, void* stream);
int fps_sine_sources_f64(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params,
                         int n_terms, int B, double* F, int64_t ldf, void* stream);
/* The 'perlin' family (data_utils/source_functions.py:173-227): params device (B, 6, 3) float64, rows 0..4 additive
 * layers {octaves, amplitude, seed} (amplitude 0 = unused), row 5 the multiplicative modulation layer (:213-220;
 * amplitude 0 = none).  Lattice gradients come from an integer hash (fast_poisson_solver_b200/sources.py mirrors it).   */
int fps_perlin_sources(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params, int B,
                       float* F, int64_t ldf, void* stream);
int fps_perlin_sources_f64(fps_ctx* ctx, const double* x, const double* y, int64_t n, const double* params, int B,
                           double* F, int64_t ldf, void* stream);
/* Problem geometry of the reference's generator (data.py:192-217): interior G x G grid (x fastest) and the 4G + 4
 * boundary ring [bottom, top, left, right]; [first, first + n) selects a rank's shard.  float64 device outputs.       */
int fps_grid_points(fps_ctx* ctx, int G, int64_t first_pde, int64_t n_pde, double* x_pde, double* y_pde,
                    int64_t first_bc, int64_t n_bc, double* x_bc, double* y_bc, void* stream);

/* ---- finite-difference accuracy oracle (numeric_solver.py:104-151 + diff_matrices.py:22-59) ------------------------
 * Solves the reference's 5-point Dirichlet system on a Gx x Gy interior grid (spacings hx, hy) for B right-hand sides
 * exactly, by type-I sine transforms evaluated as fp64 GEMMs (csrc/fd_dst.cu).  F, U: (Gx*Gy, ld) with grid point
 * (iy, ix) in row iy*Gx + ix (data.py:192-198 ordering); bc (B) constant boundary values or NULL (= 0; general boundary
 * data is folded into F by the caller, see fast_poisson_solver_b200/numeric.py).  work: fps_fd_workspace_bytes(Gx, Gy,
 * batch) bytes of device scratch; batch < B processes the right-hand sides in groups.                                  */
size_t fps_fd_workspace_bytes(int Gx, int Gy, int batch);
int fps_fd_solve(fps_ctx* ctx, int Gx, int Gy, double hx, double hy, const double* F, int64_t ldf, const double* bc,
                 int B, double* U, int64_t ldu, void* work, size_t work_bytes, void* stream);
int fps_fd_solve_f32(fps_ctx* ctx, int Gx, int Gy, double hx, double hy, const float* F, int64_t ldf, const double* bc,
                     int B, double* U, int64_t ldu, void* work, size_t work_bytes, void* stream);

/* ---- multi-GPU: the path's one exchange step -------------------------------------------------------
 * Collocation points shard across GPUs (one process and one context per GPU); every sum over points is an all-reduce:
 * the packed buffer [G | Gbc | s | N | N_bc] (FPS_PACK_DOUBLES float64) once per precompute, the projected right-hand
 * sides (FPS_FP x B float64, plus the boundary sums) once per solve.  The reference has no collective (it runs on one
 * device, solver.py:162-199).  The communicator is NCCL, resolved with dlopen at the first call; it is an object of its
 * own (one per process / GPU serves any number of contexts):
 *   rank 0:       fps_comm_unique_id(id)        -> 128 bytes, handed to the other ranks by ANY means (file, MPI, ...)
 *   every rank:   fps_comm_init(&comm, device, id, rank, world)            (collective: ncclCommInitRank)
 *   then:         fps_allreduce_f64(comm, buf, count, stream)              in place, SUM, enqueued on `stream`
 *   at the end:   fps_comm_destroy(comm)   COLLECTIVE (every rank of the node), or fps_comm_abort(comm) (local)      */
#define FPS_COMM_ID_BYTES 128
typedef struct fps_comm fps_comm;
int fps_comm_unique_id(void* h_id128);
int fps_comm_init(fps_comm** out, int device, const void* h_id128, int rank, int world);
int fps_allreduce_f64(fps_comm* comm, double* buf, int64_t count, void* stream);
int fps_comm_destroy(fps_comm* comm);
int fps_comm_abort(fps_comm* comm);

/* ---- instrumentation (bench.py / profiles) ----------------------------------------------------
 * fps_launch_count: number of kernels this library has launched in this process (all contexts).
 * fps_profile_enable(ctx, 1): bracket every kernel launch of ctx with CUDA events on its stream.
 * fps_profile_read: synchronises the recorded events and returns, for kernel class `kind`
 * (FPS_PROF_*), the summed device time in ms and the number of launches since the last
 * fps_profile_reset.                                                                            */
#define FPS_PROF_FOURIER 0
#define FPS_PROF_LAYER0 1      /* mapping.linear, K = 400                                         */
#define FPS_PROF_HIDDEN 2      /* hidden 700x700 layers writing the next layer's planes           */
#define FPS_PROF_LAST 3        /* last hidden layer writing H / DH                                */
#define FPS_PROF_GRAM 4
#define FPS_PROF_FACTOR 5
#define FPS_PROF_PROJECT 6
#define FPS_PROF_TRISOLVE 7
#define FPS_PROF_RECONSTRUCT 8
#define FPS_PROF_KINDS 9
int64_t fps_launch_count(void);
int fps_profile_enable(fps_ctx* ctx, int on);
int fps_profile_reset(fps_ctx* ctx);
int fps_profile_read(fps_ctx* ctx, int kind, double* ms, int64_t* launches);
/* With FPS_TC_STATS=1 in the environment the tcgen05 layer kernel accumulates SM-clock cycle counters
 * (operand waits, TMEM waits, promotion, epilogue, per-launch loop time); this copies the 16 counters
 * to the host array and optionally clears them.  Developer instrumentation (tools/time_features.py).   */
int fps_debug_tc_stats(unsigned long long* h_out16, int reset);

#ifdef __cplusplus
}
#endif
#endif /* FPS_SM100_H */
