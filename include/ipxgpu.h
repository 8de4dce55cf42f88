/* ipxgpu.h -- C ABI of the B200-native IPX KKT-solve path (libipxgpu.so).
 *
 * Plain C types only: pointers, sizes, int64 indices (IPX's ipxint,
 * reference include/ipx_config.h:5). No CUDA or torch types cross this
 * boundary; device pointers are passed as void*.
 *
 * Each entry point replaces one interface of the reference (file:line cited
 * per function). IPX has no FFI/plugin seam for this path; the reference-side
 * binding is link-time substitution of seven translation units, shown in
 * INTEGRATION.md, whose replacements (ipx_b200/host/*.cc) call exactly these
 * functions.
 *
 * Conventions
 *  - every function returns 0 on success or an IPXGPU_ERR_* code and never
 *    throws; ipxgpu_last_error() returns a thread-local message.
 *  - "host" pointers are ordinary host memory owned by the caller; the context
 *    copies. "_dev" entry points take device pointers on the context's device
 *    and run asynchronously on the context's stream.
 *  - there is NO CPU fallback: without a CUDA device ipxgpu_create fails.
 */
#ifndef IPXGPU_H_
#define IPXGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPXGPU_OK 0
#define IPXGPU_ERR_ARGUMENT 1
#define IPXGPU_ERR_CUDA 2
#define IPXGPU_ERR_OUT_OF_MEMORY 3
#define IPXGPU_ERR_STATE 4
#define IPXGPU_ERR_NCCL 5
#define IPXGPU_ERR_UNSUPPORTED 6

/* CR error flags, identical to reference include/ipx_status.h:31-35,47. */
#define IPXGPU_CR_ITER_LIMIT 201
#define IPXGPU_CR_MATRIX_NOT_POSDEF 202
#define IPXGPU_CR_PRECOND_NOT_POSDEF 203
#define IPXGPU_CR_NO_PROGRESS 204
#define IPXGPU_CR_INF_OR_NAN 205

typedef struct ipxgpu_ctx ipxgpu_ctx;

typedef struct ipxgpu_options {
    int32_t device;       /* CUDA ordinal; -1: env IPXGPU_DEVICE or current  */
    int32_t rank;         /* column shard held by this context ...            */
    int32_t nranks;       /* ... out of nranks (1 = whole matrix)             */
    int64_t col_begin;    /* structural columns [col_begin, col_end) of this  */
    int64_t col_end;      /* shard; -1/-1: split [0,n) evenly by nonzeros     */
    int64_t panel_cols;   /* max columns per L2-resident panel; 0 = auto      */
    void* stream;         /* cudaStream_t to run on; NULL: context-owned      */
} ipxgpu_options;

void ipxgpu_default_options(ipxgpu_options* opt);
const char* ipxgpu_last_error(void);
int ipxgpu_device_count(int* count);

/* Starts creating the CUDA runtime and the context of `device` (-1: env
 * IPXGPU_DEVICE, else driver initialisation only) on a helper thread and
 * returns at once; ipxgpu_create joins it after its host-side layout work.
 * Context creation takes 0.7-5 s in a fresh process: calling this when the
 * library is loaded hides it behind model loading (the drop-in build of IPX
 * does, see ipx_b200/host/gpu_bridge.cc). Idempotent. No reference
 * counterpart (the reference needs no device). */
int ipxgpu_warmup(int32_t device);

/* Creates a context for the solver-form matrix AI = [A I] (m rows, n+m
 * columns, CSC, sorted row indices), i.e. Model::AI() (reference
 * src/model.h:61). Uploads this shard's structural columns in CSC and CSR
 * (int32 indices on device). Replaces the `const Model&` captured by the
 * constructors of NormalMatrix / DiagonalPrecond / SplittedNormalMatrix
 * (src/normal_matrix.cc:24, src/diagonal_precond.cc:12,
 * src/splitted_normal_matrix.cc:11). */
int ipxgpu_create(ipxgpu_ctx** ctx, int64_t m, int64_t n, const int64_t* AIp,
                  const int64_t* AIi, const double* AIx,
                  const ipxgpu_options* opt);
void ipxgpu_destroy(ipxgpu_ctx* ctx);

/* Dimensions and layout facts (for tests/benchmarks):
 * out = [m, n, nnz_local, col_begin, col_end, num_panels, csc_tiles,
 *        csr_tiles]. */
int ipxgpu_get_layout(ipxgpu_ctx* ctx, int64_t out[8]);
/* Tiling of the two sweeps of the normal-matrix apply, 8 values per sweep:
 * [enabled, VB, SB, NVB, NSB, K, nparts, nitems] (sweep 1, then sweep 2).
 * enabled: 0 = generic sweep, 1 = banded sweep, 2 = banded sweep plus a
 * generic sweep over the dense segments (rows / columns) it leaves out. */
int ipxgpu_get_tiling(ipxgpu_ctx* ctx, int64_t out[16]);
int ipxgpu_synchronize(ipxgpu_ctx* ctx);

/* ---- multi-GPU: one context per rank, NCCL allreduce of the m-vector ---- */

/* Contiguous column ranges balanced by nonzeros: rank r owns structural
 * columns [bounds[r], bounds[r+1]). Pure host arithmetic (no device needed);
 * ipxgpu_create uses the same rule when col_begin/col_end are -1. */
int ipxgpu_partition_columns(int64_t n, const int64_t* AIp, int32_t nranks,
                             int64_t* bounds);
int ipxgpu_comm_unique_id(char id[128]);
int ipxgpu_comm_init(ipxgpu_ctx* ctx, const char id[128]);
/* Peer exchange over NVLink (optional; without it a sharded CR solve runs one
 * launch per stage with an NCCL allreduce per iteration). Every rank exports
 * the IPC handle of its exchange buffer (ipxgpu_peer_export), the caller
 * gathers the nranks handles in rank order (e.g. with
 * torch.distributed.all_gather) and every rank imports them
 * (ipxgpu_peer_import). The ranks' partial products then cross NVLink as
 * self-validating 16-byte records written straight into the peers' buffers
 * and summed in rank order by the receiver - inside the persistent CR kernel,
 * or by a small kernel of its own in the launch-per-stage loop - instead of a
 * collective (env IPXGPU_XCHG=pull|one|two|nccl selects
 * other exchange shapes). One process per GPU, all GPUs of one NVLink domain.
 * Replaces nothing in the reference (it has no multi-device path); it is the
 * one exchange step of the column-sharded A*D^2*A' (SURVEY.md section 8e). */
int ipxgpu_peer_export(ipxgpu_ctx* ctx, char handle[64]);
int ipxgpu_peer_import(ipxgpu_ctx* ctx, const char* handles);

/* One process, several GPUs: creates ngpus column-sharded contexts (nnz-balanced, one per
 * device; devices == NULL: ordinals 0..ngpus-1), their NCCL communicator and their NVLink peer
 * exchange (cudaDeviceEnablePeerAccess), and returns ONE handle. Calls on the handle block and
 * run the ranks on one host thread each. The handle accepts the entry points KKTSolverDiag
 * needs - ipxgpu_kktdiag_factorize, ipxgpu_kktdiag_solve (and ipxgpu_get_layout / _tiling of
 * rank 0, ipxgpu_launch_count, ipxgpu_destroy) - so that an LpSolver user gets the
 * column-sharded A*D^2*A' of SURVEY.md section 8e without managing processes: the drop-in
 * build creates such a group when IPXGPU_NGPUS > 1 (ipx_b200/host/gpu_bridge.cc). Other entry
 * points return IPXGPU_ERR_UNSUPPORTED on a group handle. the rank, column range and stream
 * fields of opt are ignored. The interrupt callback of a solve is polled by rank 0, all ranks follow it. */
int ipxgpu_create_group(ipxgpu_ctx** ctx, int64_t m, int64_t n, const int64_t* AIp,
                        const int64_t* AIi, const double* AIx, const ipxgpu_options* opt,
                        int32_t ngpus, const int32_t* devices);

/* ---- NormalMatrix (reference src/normal_matrix.h) ---- */

/* NormalMatrix::Prepare (src/normal_matrix.cc:32-35). W: host, n+m entries, or
 * NULL (W = 1 on structurals, 0 on slacks). The weights are copied to the
 * device here (the reference captures the pointer; its only caller rebuilds W
 * right before, src/kkt_solver_diag.cc:59). */
int ipxgpu_normal_prepare(ipxgpu_ctx* ctx, const double* W);
/* Same with W already on the device (full n+m vector; the shard reads its own
 * columns). */
int ipxgpu_normal_prepare_dev(ipxgpu_ctx* ctx, const void* W_dev);

/* NormalMatrix::_Apply (src/normal_matrix.cc:45-126): lhs = AI*W*AI'*rhs,
 * *rhs_dot_lhs = rhs'lhs if not NULL. Host vectors of m entries. With
 * nranks > 1 the result is allreduced, every rank returns the full product. */
int ipxgpu_normal_apply(ipxgpu_ctx* ctx, const double* rhs, double* lhs,
                        double* rhs_dot_lhs);
/* Device-resident variant: rhs_dev m doubles, lhs_dev m+1 doubles (entry m
 * receives rhs'lhs). Asynchronous on the context's stream. */
int ipxgpu_normal_apply_dev(ipxgpu_ctx* ctx, const void* rhs_dev,
                            void* lhs_dev);

/* ---- DiagonalPrecond (reference src/diagonal_precond.h) ---- */

/* DiagonalPrecond::Factorize without dense columns
 * (src/diagonal_precond.cc:28-46): diag = W[n:] + sum_j W[j]*a_ij^2.
 * W: host n+m or NULL. use_prepared != 0: reuse the device weights of the last
 * ipxgpu_normal_prepare (W is ignored). */
int ipxgpu_diag_factorize(ipxgpu_ctx* ctx, const double* W, int use_prepared);
int ipxgpu_diag_get(ipxgpu_ctx* ctx, double* diag);
/* Installs a diagonal computed elsewhere (m host doubles). */
int ipxgpu_diag_set(ipxgpu_ctx* ctx, const double* diag);
/* DiagonalPrecond::_Apply (src/diagonal_precond.cc:150-157). */
int ipxgpu_diag_apply(ipxgpu_ctx* ctx, const double* rhs, double* lhs,
                      double* rhs_dot_lhs);

/* Dense-column part of the preconditioner (reference src/diagonal_precond.cc:48-102
 * Factorize, :133-149 _Apply; at most 1000 columns, src/model.cc:34-56).
 *
 * ipxgpu_diag_factorize_masked: the diagonal E of AI*W*AI' WITHOUT the nd columns listed in
 * dense_cols (ascending, structural) - built directly with their weights zeroed, so nothing
 * is subtracted afterwards (:28-36 never adds them). W / use_prepared as above.
 * ipxgpu_smw_load: installs Ad = AI[:, dense_cols] (CSC: Adp[nd+1], Adi, Adx, rows ascending
 * per column) and the Cholesky factor L (nd x nd, column-major, lower triangle, as dpotrf
 * 'L' leaves it) of the Schur complement S = inv(Wd) + Ad' inv(E) Ad. From then on
 * ipxgpu_diag_apply and the preconditioned CR solves apply
 * inv(E) - inv(E) Ad inv(S) Ad' inv(E) entirely on the device.
 * ipxgpu_smw_clear: back to the pure diagonal (also done by ipxgpu_diag_factorize). */
int ipxgpu_diag_factorize_masked(ipxgpu_ctx* ctx, const double* W, int use_prepared,
                                 int64_t nd, const int64_t* dense_cols);
int ipxgpu_smw_load(ipxgpu_ctx* ctx, int64_t nd, const int64_t* Adp, const int64_t* Adi,
                    const double* Adx, const double* L);
int ipxgpu_smw_clear(ipxgpu_ctx* ctx);

/* ---- ConjugateResiduals (reference src/conjugate_residuals.h) ---- */

typedef struct ipxgpu_cr_result {
    int64_t errflag;   /* 0 or IPXGPU_CR_* / the interrupt callback's value */
    int64_t iter;
    double time;       /* host wall clock of the whole solve, seconds       */
    double time_op;    /* device time in C.Apply (AAt, or NNt+B+Bt)         */
    double time_pre;   /* device time in the kernels holding P.Apply         */
    double time_B;     /* split operator only: L,U solves                   */
    double time_Bt;    /* split operator only: U',L' solves                 */
    double time_NNt;   /* split operator only: N N' product                 */
    double resnorm;    /* last tested residual norm                         */
} ipxgpu_cr_result;

/* Polled between iteration batches; nonzero return aborts the solve and is
 * returned as errflag (reference Control::InterruptCheck,
 * src/conjugate_residuals.cc:84,209). May be NULL. */
typedef int64_t (*ipxgpu_interrupt_fn)(void* user);

/* ConjugateResiduals::Solve(C, P, rhs, tol, resscale, maxiter, lhs) with
 * C = the prepared normal matrix, P = the factorized diagonal
 * (src/conjugate_residuals.cc:90-213). lhs: initial iterate in, solution out.
 * resscale may be NULL. maxiter < 0 => m+100. resnorm_hist (may be NULL)
 * receives the residual norm tested at the top of each pass, up to hist_cap. */
int ipxgpu_pcr_solve(ipxgpu_ctx* ctx, const double* rhs, double tol,
                     const double* resscale, int64_t maxiter, double* lhs,
                     ipxgpu_cr_result* result, ipxgpu_interrupt_fn interrupt,
                     void* user, double* resnorm_hist, int64_t hist_cap);

/* Same solve with every vector already on the device (m doubles each;
 * resscale_dev may be NULL). zero_start != 0: lhs_dev is output only and the
 * iteration starts from 0. No host<->device vector traffic. */
int ipxgpu_pcr_solve_dev(ipxgpu_ctx* ctx, const void* rhs_dev, double tol,
                         const void* resscale_dev, int64_t maxiter,
                         void* lhs_dev, int zero_start,
                         ipxgpu_cr_result* result);

/* Unpreconditioned overload (src/conjugate_residuals.cc:14-88).
 * op: 0 = normal matrix, 1 = split (basis-preconditioned) operator. */
int ipxgpu_cr_solve(ipxgpu_ctx* ctx, int op, const double* rhs, double tol,
                    const double* resscale, int64_t maxiter, double* lhs,
                    ipxgpu_cr_result* result, ipxgpu_interrupt_fn interrupt,
                    void* user, double* resnorm_hist, int64_t hist_cap);

/* ---- KKTSolverDiag (reference src/kkt_solver_diag.cc) ---- */

/* _Factorize (:18-65): builds W and resscale on the device from the iterate
 * (xl,xu,zl,zu: host n+m each; all NULL => W = 1), prepares the normal matrix
 * and the diagonal. W_out/resscale_out (host, may be NULL) receive copies. */
int ipxgpu_kktdiag_factorize(ipxgpu_ctx* ctx, const double* xl,
                             const double* xu, const double* zl,
                             const double* zu, double mu, double* W_out,
                             double* resscale_out);
/* _Solve (:82-118): rhs = -b + AI*(W.*a), y = 0, PCR, recovery of x.
 * a, x: host n+m; b, y: host m. */
int ipxgpu_kktdiag_solve(ipxgpu_ctx* ctx, const double* a, const double* b,
                         double tol, int64_t maxiter, double* x, double* y,
                         ipxgpu_cr_result* result,
                         ipxgpu_interrupt_fn interrupt, void* user);

/* ---- context options ---- */

/* "tri_reference_order" (0 / 1; default 0, or 1 when the environment has
 * IPXGPU_TRI_ORDER=reference at context creation): summation order inside a
 * row of the triangular solves. 1: the reference's order throughout; all four
 * solves are bit-identical to the reference's loops. 0 deviates in two places,
 * both rounding-level differences (the same terms in another order):
 *  - L' solve (which = 3 below): the reference sums a column of L from the
 *    diagonal downwards (src/sparse_matrix.cc:290-303) while the solve runs
 *    upwards, so the FIRST summand of a row is the dependency that resolves
 *    LAST and every addition of the row has to wait for it; with 0 the row is
 *    summed in the order its dependencies resolve (bottom up);
 *  - rows with more than 2048 entries (any system): a serial chain of fp64
 *    additions advances by one entry every ~14 cycles on the device, so all
 *    entries but the last 2048 (in summation order) are summed as 32
 *    interleaved partial sums that are then folded in lane order.
 * Rows of L, U and U' with up to 2048 entries are bit-identical either way. */
int ipxgpu_set_option(ipxgpu_ctx* ctx, const char* name, int64_t value);

/* ---- sparse triangular solves (reference src/sparse_matrix.cc:224-311) ---- */

/* Uploads L (strict lower, unit diagonal not stored) and U (upper, diagonal
 * LAST in each column), both CSC dim x dim with sorted indices, builds their
 * row-wise copies and the level schedules of the four solves. Replaces the
 * factor export consumed by SplittedNormalMatrix::Prepare
 * (src/splitted_normal_matrix.cc:26). out_levels (may be NULL) receives the
 * level counts of [L, U, U', L'] solves. */
int ipxgpu_lu_load(ipxgpu_ctx* ctx, int64_t dim, const int64_t* Lp,
                   const int64_t* Li, const double* Lx, const int64_t* Up,
                   const int64_t* Ui, const double* Ux, int64_t out_levels[4]);
/* TriangularSolve on the loaded factors, in place on a host vector.
 * which: 0 = L ('n', unit), 1 = U ('n'), 2 = U' ('t'), 3 = L' ('t', unit);
 * 4 = ForwardSolve (L then U), 5 = BackwardSolve (U' then L'). */
int ipxgpu_tri_solve(ipxgpu_ctx* ctx, int which, double* x);

/* ---- SplittedNormalMatrix (reference src/splitted_normal_matrix.cc) ---- */

/* Prepare (:18-66) after ipxgpu_lu_load. The U passed to ipxgpu_lu_load must
 * already be column-scaled (:30-39). N = AI[:, NONBASIC] is formed on the
 * device from the resident AI: nonbasic_scale holds colscale[j] for NONBASIC
 * j and 0 for every other column (n+m host doubles, finite); rows are mapped
 * through rowperm_inv (m host int64). free_positions: pivot positions of
 * BASIC_FREE variables. */
int ipxgpu_split_prepare(ipxgpu_ctx* ctx, const double* nonbasic_scale,
                         const int64_t* rowperm_inv, int64_t num_free,
                         const int64_t* free_positions);
/* _Apply (:90-117). */
int ipxgpu_split_apply(ipxgpu_ctx* ctx, const double* rhs, double* lhs,
                       double* rhs_dot_lhs);

/* ---- KKTSolverBasis (reference src/kkt_solver_basis.cc) ---- */

/* What _Solve needs on top of ipxgpu_lu_load + ipxgpu_split_prepare (i.e. after
 * SplittedNormalMatrix::Prepare, src/kkt_solver_basis.cc:64), m entries each:
 * basic_var[k] = basis[colperm[k]], the variable at pivot position k;
 * colperm[k] = its basis position (src/basis.h:109-118);
 * basic_scale[k] = colscale[basic_var[k]] for BASIC variables and 1 for
 * BASIC_FREE ones (the positions listed as free in ipxgpu_split_prepare).
 * Unsharded contexts only. */
int ipxgpu_kktbasis_prepare(ipxgpu_ctx* ctx, const int64_t* basic_var,
                            const int64_t* colperm, const double* basic_scale);
/* Basis::SolveDense on the fresh factorization (src/basis.cc:168-170, reached
 * from src/kkt_solver_basis.cc:97,121,124,175,191): lhs = inverse(B) rhs
 * (trans 'N': rhs indexed by row, lhs by basis position) or inverse(B') rhs
 * ('T': rhs by basis position, lhs by row), as permutations + the triangular
 * solves on the loaded factors. Host vectors of m entries. */
int ipxgpu_basis_solve(ipxgpu_ctx* ctx, char trans, const double* rhs,
                       double* lhs);
/* _Solve (src/kkt_solver_basis.cc:75-194): right-hand side through the masked
 * sweeps over the NONBASIC columns, SolveDense steps, unpreconditioned CR on
 * the split operator from y = 0, recovery of (x, y). a, x: host n+m; b, y:
 * host m. result->time_NNt/_B/_Bt carry the device times of the CR loop. */
int ipxgpu_kktbasis_solve(ipxgpu_ctx* ctx, const double* a, const double* b,
                          double tol, int64_t maxiter, double* x, double* y,
                          ipxgpu_cr_result* result,
                          ipxgpu_interrupt_fn interrupt, void* user);

/* ---- Maxvolume column sweeps (reference src/maxvolume.cc:170-320) ---- */

/* What FindLargest (src/maxvolume.cc:170-200) returns over the resident
 * weights: the column of the largest |weight| (the first one among equals)
 * and of the second largest, 0 with weight 0 where there is none. */
typedef struct ipxgpu_maxvol_top {
    int64_t jmax, jmax2;
    double wmax, wmax2; /* |colweights[jmax]|, |colweights[jmax2]| */
} ipxgpu_maxvol_top;

/* Opening of a Maxvolume::Driver call (src/maxvolume.cc:220-231):
 * colweights[j] = (AI[:,j]'work) * colscale[j] where colscale[j] != 0, else 0,
 * over all n+m columns of the resident AI; then FindLargest. colscale (n+m host
 * doubles) is uploaded and stays resident with the weights until
 * ipxgpu_maxvol_release; NULL keeps the resident factors of the previous call
 * (the next slice of the same run). work: m host doubles. Unsharded contexts. */
int ipxgpu_maxvol_weights(ipxgpu_ctx* ctx, const double* colscale,
                          const double* work, ipxgpu_maxvol_top* top);
/* A column the heuristic gives up on (:262-265): colweights[j] = colscale[j] =
 * 0 (j = -1: no column). top != NULL: FindLargest over the resident weights. */
int ipxgpu_maxvol_skip(ipxgpu_ctx* ctx, int64_t j, ipxgpu_maxvol_top* top);
/* One basis update (:280-309), after jb left and jn entered the basis:
 * colscale[jb] = colscale_jb, colscale[jn] = 0; the tableau row of jb over the
 * nonbasic columns, row[j] = AI[:,j]'btran (Basis::TableauRow's dense branch,
 * src/basis.cc:266-279), folded into colweights[j] += alpha * row[j] *
 * colscale[j]; colweights[jb] = colweight_jb, colweights[jn] = 0; FindLargest.
 * btran: m host doubles (row of inverse(B) of the leaving variable). */
int ipxgpu_maxvol_update(ipxgpu_ctx* ctx, const double* btran, double alpha,
                         int64_t jb, double colscale_jb, double colweight_jb,
                         int64_t jn, ipxgpu_maxvol_top* top);
/* Copies the resident factors / weights to the host (either may be NULL). */
int ipxgpu_maxvol_get(ipxgpu_ctx* ctx, double* colscale, double* colweights);
/* Frees the resident factors and weights (end of RunHeuristic). */
int ipxgpu_maxvol_release(ipxgpu_ctx* ctx);

/* ---- products with AI outside the KKT solve ---- */

/* ipx::MultiplyAdd on the resident AI = [A I] (reference
 * src/sparse_matrix.cc:194-209): trans 'N': lhs(m) += alpha * AI * rhs(n+m);
 * trans 't'/'T': lhs(n+m) += alpha * AI' * rhs(m). Host vectors. These are the
 * residuals b - AI*x and c - AI'*y of Iterate::ComputeResiduals
 * (src/iterate.cc:543-551) and the starting point's AI'*y (src/ipm.cc:191).
 * Every sum is taken in the order of the reference's loops (ScatterColumn over
 * ascending columns / DotColumn from zero, src/sparse_matrix.h:136-152) with
 * products and additions rounded separately: the result is bit-identical to
 * the reference's. Unsharded contexts. */
int ipxgpu_multiply_add(ipxgpu_ctx* ctx, const double* rhs, double alpha,
                        double* lhs, char trans);

/* ---- measurement helpers ---- */

/* Runs `reps` device-resident normal-matrix applies on resident vectors and
 * returns the mean device time per apply in ms (CUDA events on the context's
 * stream), split into the two sweeps. flush_l2 != 0 writes a 256 MB buffer
 * between repetitions (outside the timed intervals). out_ms = [apply, sweep1
 * (A'x), sweep2 (A t)]. */
int ipxgpu_time_normal_apply(ipxgpu_ctx* ctx, int reps, int flush_l2,
                             double out_ms[3]);
/* Mean device time (CUDA events on the context's stream) of `reps` solves
 * with system `which` (0..3 as in ipxgpu_tri_solve) on a resident copy of x. */
int ipxgpu_time_tri_solve(ipxgpu_ctx* ctx, int which, int reps, const double* x,
                          double* out_ms);
/* Tuning only: with option "tri_trace" = 1 every row of a triangular solve
 * stamps %globaltimer when its warp starts it and when its value is final;
 * out receives 2*m stamps [start, finish] of the last solve, by row. */
int ipxgpu_tri_trace(ipxgpu_ctx* ctx, uint64_t* out);
/* Host-only check of the banded sweep layout (no device needed): re-tiles the
 * structural columns of AI for both sweeps of the normal-matrix apply with the
 * planner's choice (force != 0: accept any plan that fits), walks the row
 * streams on the host exactly as the kernel does and compares
 * t = A'x and y = A t with a plain CSC computation. out = [sweep-1 planned,
 * sweep-1 max abs error, sweep-1 padding fraction, sweep-2 planned, sweep-2
 * max abs error, sweep-2 padding fraction]. */
int ipxgpu_band_selftest(int64_t m, int64_t n, const int64_t* AIp, const int64_t* AIi,
                         const double* AIx, const double* x, int32_t force, double out[6]);
/* Number of kernels launched by this context since creation. */
int ipxgpu_launch_count(ipxgpu_ctx* ctx, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* IPXGPU_H_ */
