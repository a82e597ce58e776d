/*
 * lqpb.h -- C ABI of the B200-native batched ADMM box-QP solver.
 *
 * This is the drop-in boundary underneath the reference's Python API.  The reference
 * (ipo-lab/lqp_py) has no FFI of its own: its hot path is the pair of torch functions
 *
 *   torch_solve_box_qp      lqp_py/solve_box_qp_admm_torch.py:108-333   (forward ADMM solve)
 *   torch_solve_box_qp_grad lqp_py/solve_box_qp_admm_torch.py:349-432   (fixed-point backward)
 *   TorchLULayer            lqp_py/lu_layer.py:18-58                    (cached-factor LU solve)
 *
 * and every entry point below replaces one of them (see the comment on each).  All pointers
 * are plain device pointers borrowed for the duration of the call (row-major, contiguous,
 * exactly the torch layouts of the reference: Q (B,n,n), p (B,n,1), A (B,m,n), b (B,m,1),
 * lb/ub (B,n,1)); `stream` is a cudaStream_t passed as void*; the workspace is allocated by the
 * caller (torch) with the size the *_workspace_bytes function returns.  No global solver state:
 * calls on different streams with different workspaces are independent.  Return value 0 = OK,
 * otherwise an LQPB_E_* code; lqpb_last_error() gives a message for the calling thread.
 *
 * One function set per dtype: suffix _f32 (float) and _f64 (double).
 */
#ifndef LQPB_H
#define LQPB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQPB_ABI_VERSION 1

enum {
  LQPB_OK = 0,
  LQPB_E_ARG = 1,        /* bad argument (shape, null pointer, unsupported size) */
  LQPB_E_WORKSPACE = 2,  /* workspace too small */
  LQPB_E_CUDA = 3,       /* CUDA runtime error, see lqpb_last_error() */
  LQPB_E_NOT_BLACKWELL = 4, /* device is not sm_100 */
  LQPB_E_TAPE = 5           /* lqpb_unroll_forward_* in first-pass mode: the solve does not fit the tape (see there) */
};

/* status written to lqpb_info.status.  BREAKDOWN: an iterate became NaN / inf (singular or indefinite KKT system,
 * non-finite input) -- the loop stops at the stop check that sees it instead of running on to max_iters. */
enum { LQPB_STATUS_CONVERGED = 1, LQPB_STATUS_MAX_ITERS = 2, LQPB_STATUS_BREAKDOWN = 4 };

/* Derived settings of solve_box_qp_admm_torch.py:134-154 (the keys the solver *reads*),
 * flattened by the Python adapter (lqp_py_b200/solve_box_qp_admm_torch.py). */
typedef struct lqpb_config {
  int32_t max_iters;             /* :134 */
  int32_t check_solved;          /* :139 derived: max(round(sqrt(n)/10)*10, 1) unless control['check_solved'] */
  int32_t adaptive_rho;          /* :143 */
  int32_t adaptive_rho_iter;     /* :145-147 already rounded to a multiple of check_solved */
  int32_t adaptive_rho_max_iter; /* :148 ('adaptive_max_iter', default 1000) */
  int32_t scale;                 /* :152 */
  int32_t rho_auto;              /* 1 when control['rho'] is None (:200) */
  int32_t beta_auto;             /* 1 when control['beta'] is None (:171) */
  int32_t verbose;               /* :151 -- per-check residual log is returned in lqpb_info */
  int32_t keep_operators;        /* 1: leave K11 / K21 / K22 / Q~ of the solve in the workspace (the recording pass of the
                                    unrolled mode re-reads them); 0 lets small problems take the one-launch forward */
  double eps_abs;                /* :135-136 clamped to >= 1e-12 */
  double eps_rel;                /* :137-138 */
  double rho;                    /* user rho when rho_auto == 0 */
  double rho_min, rho_max;       /* :141-142 */
  double adaptive_rho_tol;       /* :144 */
  double adaptive_rho_threshold; /* :149-150 (value after rounding to the default dtype) */
  double beta;                   /* user beta when beta_auto == 0 */
  double zero_clamp;             /* :229-230 (1e-16 rounded to the default dtype) */
} lqpb_config;

#define LQPB_LOG_CAP 64

/* Host-side result record of a forward solve. */
typedef struct lqpb_info {
  int32_t iter;            /* value of the loop index at exit (:331 'iter') */
  int32_t status;          /* LQPB_STATUS_* */
  int32_t n_factor;        /* factorisations done (1 + adaptive-rho updates) */
  int32_t any_lb, any_ub;  /* :129-130 evaluated on device over the whole batch */
  int32_t n_log;           /* number of valid entries in the log (verbose) */
  int32_t log_iter[LQPB_LOG_CAP];
  double log_primal[LQPB_LOG_CAP]; /* max over batch of the primal residual at each check (:290) */
  double log_dual[LQPB_LOG_CAP];
} lqpb_info;

/* Per-phase device times (ms) of the last call on this thread when profiling is on. */
typedef struct lqpb_profile {
  float scale_ms, factor_ms, iterate_ms, finalize_ms; /* forward */
  float bwd_factor_ms, bwd_solve_ms, bwd_grad_ms;     /* backward */
  int32_t iterate_launches, factor_launches;
  int32_t kernel_launches; /* all kernels launched by the last forward/backward call */
} lqpb_profile;

int lqpb_abi_version(void);
const char* lqpb_last_error(void);
void lqpb_profile_enable(int on);
void lqpb_profile_get(lqpb_profile* out); /* blocks on the recorded events */

/* ---- forward: replaces torch_solve_box_qp (solve_box_qp_admm_torch.py:108-333) ------------
 * A/b may be NULL with m == 0.  Outputs: x,z,u (B,n), lams (B,2n), nus (B,m) (ignored if m==0),
 * rho_out (B) (rho per problem after the solve).  `info` is host memory.  The call enqueues all
 * kernels on `stream` and synchronises the stream once per segment (a segment ends at
 * convergence, at max_iters, or at an adaptive-rho update): never once per iteration. */
size_t lqpb_forward_workspace_bytes_f32(int B, int n, int m);
size_t lqpb_forward_workspace_bytes_f64(int B, int n, int m);
int lqpb_forward_f32(const lqpb_config* cfg, int B, int n, int m, const float* Q, const float* p,
                     const float* A, const float* b, const float* lb, const float* ub, float* x,
                     float* z, float* u, float* lams, float* nus, float* rho_out, lqpb_info* info,
                     void* workspace, size_t workspace_bytes, void* stream);
int lqpb_forward_f64(const lqpb_config* cfg, int B, int n, int m, const double* Q, const double* p,
                     const double* A, const double* b, const double* lb, const double* ub,
                     double* x, double* z, double* u, double* lams, double* nus, double* rho_out,
                     lqpb_info* info, void* workspace, size_t workspace_bytes, void* stream);

/* ---- additions to the reference's behaviour (the reference always starts from x = z = u = 0,
 * solve_box_qp_admm_torch.py:221-223, and keeps no record of how a solve ended, :235, :331) ---------------
 * forward_warm:     lqpb_forward_* started from the caller's z0, u0 (B,n): the UNSCALED z and u an earlier solve of a
 *                   nearby problem returned (both or neither; NULL, NULL = lqpb_forward_*).  rho0 (B) or NULL: a
 *                   per-problem rho given by the caller (the reference broadcasts a (B,1,1) tensor in control['rho'],
 *                   :200, e.g. sol['rho'] fed back); used when cfg->rho_auto == 0 instead of cfg->rho.
 * solution_status:  per-problem record of the LAST stop check of the solve whose workspace is passed (call it right
 *                   after lqpb_forward_* on the same stream): residuals (B,4) = [primal residual, dual residual,
 *                   primal tolerance, dual tolerance] (:286-304) and converged (B) = that problem's own stop test
 *                   (:307-309).  The global stop needs ALL problems converged (:312); after LQPB_STATUS_MAX_ITERS
 *                   this tells which ones were not. */
int lqpb_forward_warm_f32(const lqpb_config* cfg, int B, int n, int m, const float* Q, const float* p,
                          const float* A, const float* b, const float* lb, const float* ub, const float* z0,
                          const float* u0, const float* rho0, float* x, float* z, float* u, float* lams, float* nus,
                          float* rho_out, lqpb_info* info, void* workspace, size_t workspace_bytes, void* stream);
int lqpb_forward_warm_f64(const lqpb_config* cfg, int B, int n, int m, const double* Q, const double* p,
                          const double* A, const double* b, const double* lb, const double* ub, const double* z0,
                          const double* u0, const double* rho0, double* x, double* z, double* u, double* lams,
                          double* nus, double* rho_out, lqpb_info* info, void* workspace, size_t workspace_bytes,
                          void* stream);
/* forward_async: lqpb_forward_warm_* without the end-of-solve synchronisation, for callers that only need x (the autograd
 * layer, :26-53).  When the one-launch forward applies (small problems; adaptive-rho refactorisations happen on the device
 * there, so nothing about the solve needs the host) the call returns as soon as the bound flags are known: info->any_lb /
 * any_ub are valid, *deferred = 1, and everything else is written into `pinned_ctrl` (lqpb_ctrl_bytes() bytes of
 * page-locked host memory, owned by the caller) when the stream reaches the end of the solve; lqpb_forward_collect decodes
 * it into info after the caller has synchronised the stream.  Otherwise (*deferred = 0) the call behaved exactly like
 * lqpb_forward_warm_* and info is complete. */
size_t lqpb_ctrl_bytes(void);
int lqpb_forward_async_f32(const lqpb_config* cfg, int B, int n, int m, const float* Q, const float* p,
                           const float* A, const float* b, const float* lb, const float* ub, const float* z0,
                           const float* u0, float* x, float* z, float* u, float* lams, float* nus, float* rho_out,
                           void* pinned_ctrl, lqpb_info* info, void* workspace, size_t workspace_bytes, void* stream,
                           int32_t* deferred);
int lqpb_forward_async_f64(const lqpb_config* cfg, int B, int n, int m, const double* Q, const double* p,
                           const double* A, const double* b, const double* lb, const double* ub, const double* z0,
                           const double* u0, double* x, double* z, double* u, double* lams, double* nus,
                           double* rho_out, void* pinned_ctrl, lqpb_info* info, void* workspace,
                           size_t workspace_bytes, void* stream, int32_t* deferred);
int lqpb_forward_collect(const void* pinned_ctrl, const lqpb_config* cfg, lqpb_info* info);

/* Which iteration kernel a forward solve of this shape takes on the current device (SURVEY App. C regimes; needs a CUDA
 * device): the x-update operator streamed from L2 / HBM every iteration, held in shared memory as packed tiles, held in
 * shared memory as dense matrices, or the latter inside the one-launch forward.  -1: no sm_100 device. */
enum { LQPB_REGIME_STREAM = 0, LQPB_REGIME_PACKED_RESIDENT = 1, LQPB_REGIME_ROWS = 2, LQPB_REGIME_FUSED_ROWS = 3,
       LQPB_REGIME_STREAM_SPLIT = 4 /* streamed, every problem split over a cluster of 2 or 4 CTAs (small batches) */ };
int lqpb_iterate_regime_f32(const lqpb_config* cfg, int B, int n, int m);
int lqpb_iterate_regime_f64(const lqpb_config* cfg, int B, int n, int m);
int lqpb_solution_status_f32(const lqpb_config* cfg, int B, int n, int m, void* workspace, size_t workspace_bytes,
                             float* residuals, int32_t* converged, void* stream);
int lqpb_solution_status_f64(const lqpb_config* cfg, int B, int n, int m, void* workspace, size_t workspace_bytes,
                             double* residuals, int32_t* converged, void* stream);

/* ---- backward: replaces torch_solve_box_qp_grad (solve_box_qp_admm_torch.py:349-432) ------
 * rho_dev: per-problem rho (B) or NULL, in which case rho_scalar is used (:356-357, :379-382).
 * Any of dQ (B,n,n), dp (B,n), dA (B,m,n), db (B,m), dlb (B,n), dub (B,n) may be NULL: that
 * gradient is skipped (ctx.needs_input_grad).  Fully asynchronous on `stream`. */
size_t lqpb_backward_workspace_bytes_f32(int B, int n, int m);
size_t lqpb_backward_workspace_bytes_f64(int B, int n, int m);
int lqpb_backward_f32(int B, int n, int m, const float* dl_dz, const float* x, const float* u,
                      const float* lams, const float* nus, const float* Q, const float* A,
                      const float* lb, const float* ub, const float* rho_dev, double rho_scalar,
                      float* dQ, float* dp, float* dA, float* db, float* dlb, float* dub,
                      void* workspace, size_t workspace_bytes, void* stream);
int lqpb_backward_f64(int B, int n, int m, const double* dl_dz, const double* x, const double* u,
                      const double* lams, const double* nus, const double* Q, const double* A,
                      const double* lb, const double* ub, const double* rho_dev, double rho_scalar,
                      double* dQ, double* dp, double* dA, double* db, double* dlb, double* dub,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- KKT backward: replaces torch_solve_box_qp_grad_kkt (solve_box_qp_admm_torch.py:435-584), the
 * backward='kkt' mode of SolveBoxQPLayer.backward (:63-64).  Same outputs and NULL-skipping as
 * lqpb_backward_*; workspace of lqpb_backward_workspace_bytes_*.  The 2n inequality rows of the
 * reference's (3n+m) system are eliminated in closed form, leaving the symmetric
 * [[Q + diag(lam_lo/s_lo + lam_hi/s_hi), A^T], [A, 0]] with the clamps of :450-451.
 * any_bounds: optional HOST int[2] receiving any_lb, any_ub (:439-440) -- the reference returns
 * dlb = None / dub = None without them (:572-579); passing it synchronises the stream once, NULL keeps
 * the call fully asynchronous (the layer already knows the flags from the forward solve). */
int lqpb_backward_kkt_f32(int B, int n, int m, const float* dl_dz, const float* x, const float* lams,
                          const float* nus, const float* Q, const float* A, const float* lb,
                          const float* ub, float* dQ, float* dp, float* dA, float* db, float* dlb,
                          float* dub, int32_t* any_bounds, void* workspace, size_t workspace_bytes,
                          void* stream);
int lqpb_backward_kkt_f64(int B, int n, int m, const double* dl_dz, const double* x, const double* lams,
                          const double* nus, const double* Q, const double* A, const double* lb,
                          const double* ub, double* dQ, double* dp, double* dA, double* db,
                          double* dlb, double* dub, int32_t* any_bounds, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- forward solve that also prepares the backward.  The adjoint system of either backward mode depends on the
 * solution only, not on the upstream gradient: forward_prep queues its mask / KKT diagonal, assembly and block LDL^T
 * into `bwd_workspace` (lqpb_backward_workspace_bytes_*) right behind the solve and returns as soon as the SOLVE has
 * finished, so the GPU keeps working while the caller travels from .forward() to .backward() (reference :26-67: the
 * same Function, same saved tensors).  *prepared = 1 when that was done (tensor-core factorisation path: fp32,
 * n + m > 128), 0 otherwise (call lqpb_backward_* / lqpb_backward_kkt_* as usual).  backward_finish then runs only the
 * substitution with dl_dz and the gradient assembly on the prepared workspace; arguments as lqpb_backward_*, kkt != 0
 * for the KKT mode (u, rho_dev unused).  x, u, lams, nus, Q, A, lb, ub must be the buffers of the forward_prep call,
 * unchanged, and no other call may have used bwd_workspace in between. */
int lqpb_forward_prep_f32(const lqpb_config* cfg, int B, int n, int m, const float* Q, const float* p,
                          const float* A, const float* b, const float* lb, const float* ub, float* x, float* z,
                          float* u, float* lams, float* nus, float* rho_out, lqpb_info* info, void* workspace,
                          size_t workspace_bytes, void* bwd_workspace, size_t bwd_workspace_bytes, int kkt,
                          int32_t* prepared, void* stream);
int lqpb_forward_prep_f64(const lqpb_config* cfg, int B, int n, int m, const double* Q, const double* p,
                          const double* A, const double* b, const double* lb, const double* ub, double* x,
                          double* z, double* u, double* lams, double* nus, double* rho_out, lqpb_info* info,
                          void* workspace, size_t workspace_bytes, void* bwd_workspace,
                          size_t bwd_workspace_bytes, int kkt, int32_t* prepared, void* stream);
int lqpb_backward_finish_f32(int B, int n, int m, int kkt, const float* dl_dz, const float* x, const float* u,
                             const float* lams, const float* nus, const float* Q, const float* A,
                             const float* lb, const float* ub, const float* rho_dev, double rho_scalar, float* dQ,
                             float* dp, float* dA, float* db, float* dlb, float* dub, void* workspace,
                             size_t workspace_bytes, void* stream);
int lqpb_backward_finish_f64(int B, int n, int m, int kkt, const double* dl_dz, const double* x, const double* u,
                             const double* lams, const double* nus, const double* Q, const double* A,
                             const double* lb, const double* ub, const double* rho_dev, double rho_scalar,
                             double* dQ, double* dp, double* dA, double* db, double* dlb, double* dub,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- host-buffer forms of the two calls above: what a caller of the reference holds are CPU tensors
 * (experiments/experiment_1.py:58-77 builds Q, p, ... on the host, calls .forward and .backward and reads
 * x and the .grad fields on the host).  h* pointers are HOST memory (page-locked memory keeps the copies
 * asynchronous; pageable memory works, slower); the unprefixed pointers are DEVICE buffers of the same
 * shapes supplied by the caller: forward_host fills Q, p, A, b, lb, ub with the device copies (keep them
 * for the backward, like ctx.save_for_backward in :48) and x, z, u, lams, nus, rho_out with the results;
 * hx (may be NULL) receives x.  backward_host uploads h_dl_dz into dl_dz, leaves every requested gradient
 * in its device buffer and copies it to the matching h* pointer (NULL = not wanted on the host); kkt != 0
 * selects the KKT backward (u, rho_dev unused).  The batch is cut into `chunks` (0 = choose) slices of
 * whole problems: a copy stream uploads Q slice c + 1 while slice c is scaled and factorised, and returns
 * the dQ rows of slice c while slice c + 1 is differentiated.  Both calls return after `stream` and the
 * copy stream have drained (the host buffers are valid on return).  forward_host can prepare the backward like
 * lqpb_forward_prep_* (bwd_workspace may be NULL: no preparation; *prepared tells whether it happened), and
 * backward_host with prepared != 0 then runs only the substitution and the gradient assembly per chunk on that
 * workspace, so the first dQ rows start their trip to the host without waiting for a factorisation.
 * Overlapping the NEXT batch's upload with this batch's gradient download (PCIe is full duplex): h_dl_dz may also be a
 * DEVICE pointer (it is copied with cudaMemcpyDefault).  The H2D copy engine serves its queue in submission order, so a
 * caller that wants the overlap uploads dl_dz itself, THEN enqueues the next batch's cudaMemcpyAsync(s) on a stream of
 * its own, THEN calls backward_host with the device copy of dl_dz, and gives the next lqpb_forward_prep_* the device
 * buffers once its copy event has completed (lqp_py_b200.solve_box_qp_admm_torch.prefetch_inputs does exactly this).
 * h_dl_dz == dl_dz means "dl_dz is already in place" (no copy at all: e.g. brought up with lqpb_copy_mapped). */
int lqpb_forward_host_f32(const lqpb_config* cfg, int B, int n, int m, const float* hQ, const float* hp,
                          const float* hA, const float* hb, const float* hlb, const float* hub, float* Q,
                          float* p, float* A, float* b, float* lb, float* ub, float* x, float* z, float* u,
                          float* lams, float* nus, float* rho_out, float* hx, lqpb_info* info,
                          void* workspace, size_t workspace_bytes, void* stream, int chunks, void* bwd_workspace,
                          size_t bwd_workspace_bytes, int bwd_kkt, int32_t* prepared);
int lqpb_forward_host_f64(const lqpb_config* cfg, int B, int n, int m, const double* hQ, const double* hp,
                          const double* hA, const double* hb, const double* hlb, const double* hub,
                          double* Q, double* p, double* A, double* b, double* lb, double* ub, double* x,
                          double* z, double* u, double* lams, double* nus, double* rho_out, double* hx,
                          lqpb_info* info, void* workspace, size_t workspace_bytes, void* stream,
                          int chunks, void* bwd_workspace, size_t bwd_workspace_bytes, int bwd_kkt,
                          int32_t* prepared);
int lqpb_backward_host_f32(int B, int n, int m, int kkt, const float* h_dl_dz, float* dl_dz, const float* x,
                           const float* u, const float* lams, const float* nus, const float* Q,
                           const float* A, const float* lb, const float* ub, const float* rho_dev,
                           double rho_scalar, float* dQ, float* dp, float* dA, float* db, float* dlb,
                           float* dub, float* hdQ, float* hdp, float* hdA, float* hdb, float* hdlb,
                           float* hdub, int32_t* any_bounds, void* workspace, size_t workspace_bytes,
                           void* stream, int chunks, int prepared);
int lqpb_backward_host_f64(int B, int n, int m, int kkt, const double* h_dl_dz, double* dl_dz,
                           const double* x, const double* u, const double* lams, const double* nus,
                           const double* Q, const double* A, const double* lb, const double* ub,
                           const double* rho_dev, double rho_scalar, double* dQ, double* dp, double* dA,
                           double* db, double* dlb, double* dub, double* hdQ, double* hdp, double* hdA,
                           double* hdb, double* hdlb, double* hdub, int32_t* any_bounds, void* workspace,
                           size_t workspace_bytes, void* stream, int chunks, int prepared);

/* ---- unrolled mode: control['unroll'] = True (solve_box_qp_admm_torch.py:13-15) lets autograd differentiate
 * every ADMM iteration (:259-282), each KKT solve through TorchLULayer (lu_layer.py:18-58: dx = M^-1 (-g),
 * dl_dA = dx xv^T, dl_db = -dx).  Here the loop is recorded on a tape and swept backwards by a kernel.  All calls
 * follow a plain lqpb_forward_* solve of the same problem, which tells n_iter = info.iter + 1 and
 * n_seg = info.n_factor (operator segments: 1 + adaptive-rho updates, :246-256).
 *
 * tapes: tape_x / tape_z / tape_u (B, n_iter, n) hold the scaled iterate x~_k, z_k, u_k after iteration k,
 *        tape_nu (B, n_iter, m) the equality part of the KKT solve of iteration k (unused when m == 0).
 * unroll_record:   n_seg == 1 only.  Re-runs the n_iter iterations from z = u = 0 on the operators still held in
 *                  the forward call's `workspace` (identical arithmetic) and fills the tapes.
 * unroll_forward:  any n_seg.  A complete recording solve (same arguments / outputs as lqpb_forward_*) that also
 *                  keeps, per operator segment s, a snapshot of (K11, K21, K22, c, rho) in `snapshots`
 *                  (n_seg x lqpb_unroll_snapshot_bytes_*), writes the first iteration of every segment and n_iter
 *                  into the HOST array seg_start[n_seg + 1], and the do_rho_update flags (:310-311) each update
 *                  applied into the DEVICE array wants (n_seg - 1, B).
 *                  Both recording calls synchronise the stream (they verify the run ended at iteration n_iter - 1).
 *                  First-pass mode (snapshots == NULL, n_seg == 1): unroll_forward IS the solve -- n_iter is then only
 *                  the capacity of the tapes (their row stride); the call succeeds when the solve converges (or hits
 *                  cfg->max_iters) within n_iter iterations without an adaptive-rho refactorisation, info->iter + 1
 *                  rows are valid and the reverse sweep reads the operators from `workspace`; otherwise it returns
 *                  LQPB_E_TAPE and the caller falls back to a plain solve followed by a recording call.
 * unroll_backward: reverse sweep over the iterations k_hi .. k_lo (inclusive) of ONE operator segment: operators from
 *                  `snapshot` (one snapshot of unroll_forward) or, when NULL, from `workspace`.  In: adjoints of
 *                  x~_{k_hi}, z_{k_hi}, u_{k_hi}, z_{k_hi - 1} (g_x, g_z, g_u, g_zprev, (B, n) each, NULL = zero).
 *                  Out: the adjoints of the SCALED problem data accumulated over the range -- gp (B,n), gb (B,m), glb,
 *                  gub (B,n), grho (B) and, when not NULL, gQ (B,n,n) = -sum_k w_k x_k^T (not symmetric, like the
 *                  reference's) and gA (B,m,n) -- and those of the state the range started from, gz_in / gu_in
 *                  (z_{k_lo - 1}, u_{k_lo - 1}; may be NULL).  tape_w (B,n_iter,n) / tape_wnu (B,n_iter,m) are
 *                  scratch for the adjoint solves.  One symmetric GEMV with the cached K11 per iteration -- the
 *                  operator stream of the forward loop.  Fully asynchronous on `stream`.
 * The adapter (lqp_py_b200/solve_box_qp_admm_torch.py) chains the ranges around each adaptive-rho update and maps
 * the scaled adjoints to the caller's Q, p, A, b, lb, ub through the scaling (:161-203). */
size_t lqpb_unroll_snapshot_bytes_f32(int B, int n, int m);
size_t lqpb_unroll_snapshot_bytes_f64(int B, int n, int m);
int lqpb_unroll_record_f32(const lqpb_config* cfg, int B, int n, int m, int n_iter, void* workspace,
                           size_t workspace_bytes, float* tape_x, float* tape_z, float* tape_u, float* tape_nu,
                           void* stream);
int lqpb_unroll_record_f64(const lqpb_config* cfg, int B, int n, int m, int n_iter, void* workspace,
                           size_t workspace_bytes, double* tape_x, double* tape_z, double* tape_u, double* tape_nu,
                           void* stream);
int lqpb_unroll_forward_f32(const lqpb_config* cfg, int B, int n, int m, int n_iter, int n_seg, const float* Q,
                            const float* p, const float* A, const float* b, const float* lb, const float* ub,
                            float* x, float* z, float* u, float* lams, float* nus, float* rho_out, float* tape_x,
                            float* tape_z, float* tape_u, float* tape_nu, void* snapshots, size_t snapshot_bytes,
                            int32_t* seg_start, int32_t* wants, lqpb_info* info, void* workspace,
                            size_t workspace_bytes, void* stream);
int lqpb_unroll_forward_f64(const lqpb_config* cfg, int B, int n, int m, int n_iter, int n_seg, const double* Q,
                            const double* p, const double* A, const double* b, const double* lb, const double* ub,
                            double* x, double* z, double* u, double* lams, double* nus, double* rho_out,
                            double* tape_x, double* tape_z, double* tape_u, double* tape_nu, void* snapshots,
                            size_t snapshot_bytes, int32_t* seg_start, int32_t* wants, lqpb_info* info,
                            void* workspace, size_t workspace_bytes, void* stream);
int lqpb_unroll_backward_f32(int B, int n, int m, int n_iter, int k_lo, int k_hi, void* workspace,
                             size_t workspace_bytes, const void* snapshot, const float* g_x, const float* g_z,
                             const float* g_u, const float* g_zprev, const float* tape_x, const float* tape_z,
                             const float* tape_u, const float* tape_nu, float* tape_w, float* tape_wnu, float* gQ,
                             float* gp, float* gA, float* gb, float* glb, float* gub, float* grho, float* gz_in,
                             float* gu_in, void* stream);
int lqpb_unroll_backward_f64(int B, int n, int m, int n_iter, int k_lo, int k_hi, void* workspace,
                             size_t workspace_bytes, const void* snapshot, const double* g_x, const double* g_z,
                             const double* g_u, const double* g_zprev, const double* tape_x, const double* tape_z,
                             const double* tape_u, const double* tape_nu, double* tape_w, double* tape_wnu,
                             double* gQ, double* gp, double* gA, double* gb, double* glb, double* gub, double* grho,
                             double* gz_in, double* gu_in, void* stream);

/* unroll_scale_grad: adjoint of the two O(n^2) steps that lead from the caller's Q to the scaled problem,
 * Q~ = D Q D (:176) and rho = clamp(||Q~||_F / sqrt(n)) (:200-203), in one pass.  In: G (B,n,n) = adjoint of Q~ (gQ of
 * unroll_backward), Q (B,n,n), D (B,n) or NULL when the solve ran without scaling, coef (B) = grho / (n rho) (zero where
 * rho was clamped; NULL when rho was given by the caller).  Out: G overwritten with the adjoint of Q for fixed D,
 * gD (B,n) the adjoint of D through Q~ and rho (NULL iff D is NULL).  scratch: *_scratch_elems(B, n) elements. */
size_t lqpb_unroll_scale_grad_scratch_elems_f32(int B, int n);
size_t lqpb_unroll_scale_grad_scratch_elems_f64(int B, int n);
int lqpb_unroll_scale_grad_f32(int B, int n, float* G, const float* Q, const float* D, const float* coef, float* gD,
                               float* scratch, void* stream);
int lqpb_unroll_scale_grad_f64(int B, int n, double* G, const double* Q, const double* D, const double* coef,
                               double* gD, double* scratch, void* stream);

/* The O(B n) part of the same scaling map (:161-197): (column inf-norms of Q, p, A, b, lb, ub) -> (D, p~ = D p, A~ = E (A D),
 * b~ = E b, lb~ = lb / D, ub~ = ub / D) with D = blend(sqrt(1 / guard(norms))) (:163-175) and E = 1 / guard(||A D||_inf)
 * (:180-188).  scaled_vectors copies the VALUES out of the workspace of the recording solve (D, pt, lbt, ubt (B,n); At (B,m,n);
 * bt, E (B,m)); scale_vec_grad is the adjoint of the whole map in one kernel per problem, with torch's subgradient rules
 * (inf-norms split evenly among exact ties, torch.quantile sends (1 - w, w) to the two order statistics it interpolates,
 * guarded entries route to the mean).  g* inputs may be NULL (= zero); use_lb / use_ub = the bound vector enters the loop
 * (any_lb / any_ub); beta_auto = control['beta'] is None. */
int lqpb_unroll_scaled_vectors_f32(int B, int n, int m, void* workspace, size_t workspace_bytes, float* D, float* pt,
                                   float* At, float* bt, float* lbt, float* ubt, float* E, void* stream);
int lqpb_unroll_scaled_vectors_f64(int B, int n, int m, void* workspace, size_t workspace_bytes, double* D, double* pt,
                                   double* At, double* bt, double* lbt, double* ubt, double* E, void* stream);
int lqpb_unroll_scale_vec_grad_f32(int B, int n, int m, int beta_auto, double beta, int use_lb, int use_ub,
                                   const float* colmax, const float* p, const float* A, const float* b, const float* lb,
                                   const float* ub, const float* D, const float* E, const float* gD, const float* gD2,
                                   const float* gpt, const float* gAt, const float* gbt, const float* glbt,
                                   const float* gubt, float* gcolmax, float* gp, float* gA, float* gb, float* glb,
                                   float* gub, void* stream);
int lqpb_unroll_scale_vec_grad_f64(int B, int n, int m, int beta_auto, double beta, int use_lb, int use_ub,
                                   const double* colmax, const double* p, const double* A, const double* b,
                                   const double* lb, const double* ub, const double* D, const double* E, const double* gD,
                                   const double* gD2, const double* gpt, const double* gAt, const double* gbt,
                                   const double* glbt, const double* gubt, double* gcolmax, double* gp, double* gA,
                                   double* gb, double* glb, double* gub, void* stream);
/* The column inf-norms of Q (:163: torch.linalg.norm(Q, inf, dim=1)) and their adjoint: colmax (B,n); colmax_grad adds
 * sign(Q_ij) gcolmax_j / (number of maximisers of column j) IN PLACE to G (B,n,n) on every maximiser of every column --
 * torch's inf-norm backward, exact ties included (gD and gD2 of scale_vec_grad are both added to the adjoint of D). */
int lqpb_unroll_colmax_f32(int B, int n, const float* Q, float* colmax, void* stream);
int lqpb_unroll_colmax_f64(int B, int n, const double* Q, double* colmax, void* stream);
int lqpb_unroll_colmax_grad_f32(int B, int n, const float* Q, const float* colmax, const float* gcolmax, float* G,
                                void* stream);
int lqpb_unroll_colmax_grad_f64(int B, int n, const double* Q, const double* colmax, const double* gcolmax, double* G,
                                void* stream);

/* ---- lu_layer: replaces TorchLU / TorchLULayer (lu_layer.py:5-58) -------------------------
 * lu_factor: partial-pivoting LU of B general N x N matrices (torch.linalg.lu_factor, :10,:30);
 *            LU (B,N,N) packed L\U, piv (B,N) 1-based row swaps like LAPACK getrf.
 * lu_solve:  X = A^-1 RHS from cached factors (torch.linalg.lu_solve, :33,:52); RHS/X (B,N,nrhs);
 *            `negate_rhs` solves with -RHS (:52).
 * outer:     C (B,N,M) = a (B,N) b(B,M)^T, the dl_dA = dx x^T of :53. */
int lqpb_lu_factor_f32(int B, int N, const float* A, float* LU, int32_t* piv, void* stream);
int lqpb_lu_factor_f64(int B, int N, const double* A, double* LU, int32_t* piv, void* stream);
int lqpb_lu_solve_f32(int B, int N, int nrhs, const float* LU, const int32_t* piv, const float* rhs,
                      float* x, int negate_rhs, void* stream);
int lqpb_lu_solve_f64(int B, int N, int nrhs, const double* LU, const int32_t* piv,
                      const double* rhs, double* x, int negate_rhs, void* stream);
int lqpb_outer_f32(int B, int N, int M, const float* a, const float* b, float* C, void* stream);
int lqpb_outer_f64(int B, int N, int M, const double* a, const double* b, double* C, void* stream);

/* ---- developer / diagnostic entry points (not part of the reference-facing surface) ----------
 * Inverse of B symmetric (quasi-definite) N x N fp32 matrices, N a multiple of 128, through the
 * tensor-core blocked sweep that lqpb_forward_f32 uses for n + m > 128 (csrc/tcfactor.cu); lets
 * tools/tc_check.py and the GPU tests measure the tcgen05 3xTF32 path in isolation. */
size_t lqpb_dev_tc_inverse_work_bytes(int B, int N);
int lqpb_dev_tc_inverse_f32(int B, int N, const float* A, float* Ainv, void* work, void* stream);
/* Streaming read of `bytes` bytes of device memory, `reps` passes, 16-byte loads that bypass L1: bench.py times it on a
 * buffer smaller than the L2 (and on one far larger) to measure the read-bandwidth ceilings the iteration kernel's
 * roofline is quoted against.  sink: 4 bytes of device memory (never written for real data). */
int lqpb_dev_stream_read(const void* buf, size_t bytes, int reps, void* sink, void* stream);

/* ---- small transfers beside the copy engines ----------------------------------------------------
 * Copies `bytes` (a multiple of 4; both pointers 4-byte aligned) between a device buffer and PAGE-LOCKED host memory
 * (cudaMallocHost / cudaHostAlloc / torch pin_memory: mapped into the device's address space under unified addressing)
 * with a kernel on `stream` -- SM loads / stores over PCIe -- instead of a copy engine.  A copy engine finishes the
 * transfer it has started before it serves the next one, so a caller that keeps both engines busy with 128 MB batches
 * (next batch up, last gradients down: SolveBoxQP.solve_ahead) moves dl_dz, x and the control block this way; the
 * library itself reads the control block at the end of every forward segment like this.  The host side of the data is
 * valid once `stream` has reached the point after the call (event / stream synchronisation), as with cudaMemcpyAsync. */
int lqpb_copy_mapped(void* dst, const void* src, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LQPB_H */
