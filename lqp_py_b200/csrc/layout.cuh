// Workspace layout (HBM) of one forward / backward call and the kernel launch prototypes.
// All matrices are row-major; `ld` is the padded row stride of the compact n x n matrices
// (16-byte multiple so that rows are legal bulk-TMA sources), `np` the dimension padded to the
// 32 x 32 tiles of the Gauss-Jordan inversion.
#pragma once
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "../../include/lqpb.h"

namespace lqpb {

constexpr int kTile = 32;      // Gauss-Jordan tile edge
constexpr int kMacro = 64;     // macro tile edge of the trailing update (np is a multiple of it)
constexpr int kMaxM = 256;     // max equality rows (they live in the padding of the factorisation; per-CTA scratch is sized by it)
constexpr int kTcBlock = 128;  // block edge of the tensor-core factorisation (tcfactor.cu)

// Problems with n + m > 128 are factorised by the blocked sweep with 128 x 128 tile products on the tensor cores: fp32 in
// tcfactor.cu / tcfused.cu (tcgen05, 3xTF32), fp64 in f64block.cu (DMMA); small problems use the Gauss-Jordan kernel of
// factor.cu.  LQPB_FACTOR=gj in the environment forces the latter for every size (A/B measurements).
template <typename T> inline bool tc_factor_enabled(int n, int m) {
  const char* e = getenv("LQPB_FACTOR");
  if (e && !strcmp(e, "gj")) return false;
  return n + m > kTcBlock;
}

// ------------------------------------------------------------------ packed symmetric storage
// The two O(n^2) operands of the iteration kernel -- the x-update operator K11 and the scaled Q~ -- are
// symmetric, so only the lower triangle is stored and streamed: half the bytes per ADMM iteration.
// Layout: tiles of 32 rows x TC columns (TC = 32 floats / 16 doubles: 4 KB either way), ordered by
// block column Jc and, inside a column, by block row I >= Jc / R (R = 32 / TC) -- a contiguous tile
// sequence that is cut into per-warp runs.  Inside a tile the data is CHUNK-MAJOR: the 16-byte chunk k (columns
// k VN .. k VN + VN - 1) of all 32 rows is one contiguous 512-byte segment, row l at offset 16 l inside it.  The
// access every consumer makes -- "lane l reads chunk k of row l" -- is then one fully coalesced 512-byte request
// when the tile is read straight from global memory / L2 (streaming iteration kernel) and a bank-conflict-free
// LDS.128 when the tile sits in shared memory (resident kernels, TMA-staged reverse sweep).  Entries above the
// diagonal (inside diagonal tiles) and in the padding rows/columns are zero and the diagonal is stored HALVED: every
// tile can then be applied symmetrically ( x_I += T v_J  and  x_J += T^T v_I ) without a special case for diagonal
// tiles.  in_tile() is the ONE place that encodes the order inside a tile.
constexpr int kPackRows = 32;
template <typename T>
struct Pack {
  static constexpr int VN = Vec<T>::N;          // elements per 16-byte chunk
  static constexpr int TC = 8 * VN;             // tile columns
  static constexpr int R = kPackRows / TC;      // block columns per block row (1 or 2)
  static constexpr int TILE = kPackRows * TC;   // elements per tile (4096 bytes)
  // position of (row l, 16-byte chunk k, element e of the chunk) inside a tile
  __host__ __device__ static int in_tile(int l, int k, int e = 0) { return (k * kPackRows + l) * VN + e; }
  __host__ __device__ static int nt(int n) { return (n + kPackRows - 1) / kPackRows; }
  __host__ __device__ static int nbc(int n) { return nt(n) * R; }
  __host__ __device__ static int col_start(int Jc, int ntv) {     // tiles in the block columns before Jc
    const int q = Jc / R, rem = Jc % R;
    return Jc * ntv - (R * (q * (q - 1) / 2) + rem * q);
  }
  __host__ __device__ static int ntiles(int n) { return col_start(nbc(n), nt(n)); }
  __host__ __device__ static size_t elems(int n) { return (size_t)ntiles(n) * TILE; }
  // offset of element (i, j), j <= i
  __host__ __device__ static size_t offset(int i, int j, int ntv) {
    const int Jc = j / TC, I = i / kPackRows, l = i % kPackRows, c = j % TC, k = c / VN, e = c % VN;
    return (size_t)(col_start(Jc, ntv) + I - Jc / R) * TILE + in_tile(l, k, e);
  }
};

template <typename T>
struct FwdWs {
  int B, n, m, ld, np;
  int tc, nb;   // tensor-core factorisation: W = block-lower tiles, Vg / Wg / Pb = nb tiles per problem
  T* Pb;        // B*nb*128*128  pivot-block inverses (tc only)
  T* Qp;        // B*Pack::elems(n)  scaled Q~ = D Q D, packed lower triangle (read at every check: Q~ x~;
                //                   also the source of the factorisation)
  T* Kp;        // B*Pack::elems(n)  x-update operator K11, packed lower triangle -- streamed every iteration
  T* W;         // B*np*np  Gauss-Jordan work matrix (lower triangle)
  T* Vg;        // B*np*kTile   column panel before the sweep step
  T* Wg;        // B*np*kTile   column panel after the sweep step
  T *D, *pt, *lbt, *ubt, *c, *z, *u, *xs;   // B*ld each: scaling, p~, lb~, ub~, c = K12 b~, ADMM state, last x~
  T *At, *Gt;   // B*m*ld   A~ and K21 (the (2,1) block of the KKT inverse)
  T* Sinv;      // B*m*m    K22 (the (2,2) block of the KKT inverse = -(A~ H^-1 A~^T)^-1)
  T *bt, *E;    // B*m
  T *rho, *rho_cand, *pnorm, *ratio;   // B
  double* fro_part;  // B*n_fro  per-CTA partial sums of ||Q~||_F^2 (scale_pack_kernel), added in order by select_rho
  int n_fro;
  T* chk;       // B*4  [primal, dual, tol_primal_rel, tol_dual_rel] of the last check
  int* wants;   // B    do_rho_update of the last check
  Ctrl* ctrl;
  const T *z0, *u0;   // optional warm start (caller's UNSCALED z, u of an earlier solve, (B, n)); null = zero start (:221-223)
  const T* rho_in;    // optional per-problem rho (B) given by the caller (a (B,1,1) tensor in control['rho']); null = cfg.rho
  size_t bytes;
};

template <typename T>
inline FwdWs<T> carve_fwd(void* base, int B, int n, int m) {
  FwdWs<T> w;
  w.B = B; w.n = n; w.m = m;
  w.z0 = w.u0 = nullptr;
  w.rho_in = nullptr;
  w.ld = round_up(n, Vec<T>::N);
  w.np = round_up(n + m, kMacro);
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t count, size_t elt) {
    void* r = p ? p + off : nullptr;
    off = round_up_sz(off + count * elt, 256);
    return r;
  };
  const size_t Bn = (size_t)B;
  w.tc = tc_factor_enabled<T>(n, m) ? 1 : 0;
  w.nb = (n + m + kTcBlock - 1) / kTcBlock;
  const size_t tile_e = (size_t)kTcBlock * kTcBlock;
  w.ctrl = (Ctrl*)take(1, sizeof(Ctrl));
  w.Qp = (T*)take(Bn * Pack<T>::elems(n), sizeof(T));
  w.Kp = (T*)take(Bn * Pack<T>::elems(n), sizeof(T));
  if (w.tc) {
    w.W = (T*)take(Bn * ((size_t)w.nb * (w.nb + 1) / 2) * tile_e, sizeof(T));
    w.Vg = (T*)take(Bn * w.nb * tile_e, sizeof(T));
    w.Wg = (T*)take(Bn * w.nb * tile_e, sizeof(T));
    w.Pb = (T*)take(Bn * w.nb * tile_e, sizeof(T));
  } else {
    w.W = (T*)take(Bn * w.np * w.np, sizeof(T));
    w.Vg = (T*)take(Bn * w.np * kTile, sizeof(T));
    w.Wg = (T*)take(Bn * w.np * kTile, sizeof(T));
    w.Pb = nullptr;
  }
  T** vecs[] = {&w.D, &w.pt, &w.lbt, &w.ubt, &w.c, &w.z, &w.u, &w.xs};
  for (auto v : vecs) *v = (T*)take(Bn * w.ld, sizeof(T));
  const size_t mm = m > 0 ? m : 1;
  w.At = (T*)take(Bn * mm * w.ld, sizeof(T));
  w.Gt = (T*)take(Bn * mm * w.ld, sizeof(T));
  w.Sinv = (T*)take(Bn * mm * mm, sizeof(T));
  w.bt = (T*)take(Bn * mm, sizeof(T));
  w.E = (T*)take(Bn * mm, sizeof(T));
  w.rho = (T*)take(Bn, sizeof(T));
  w.rho_cand = (T*)take(Bn, sizeof(T));
  w.pnorm = (T*)take(Bn, sizeof(T));
  w.ratio = (T*)take(Bn, sizeof(T));
  w.n_fro = (Pack<T>::ntiles(n) + 7) / 8;      // 8 packed tiles (one per warp) per scale_pack CTA
  w.fro_part = (double*)take(Bn * w.n_fro, sizeof(double));
  w.chk = (T*)take(Bn * 4, sizeof(T));
  w.wants = (int*)take(Bn, sizeof(int));
  w.bytes = off;
  return w;
}

template <typename T>
struct BwdWs {
  int B, n, m, ld, np;
  int tc, nb;   // see FwdWs
  T* Pb;
  T* W;         // B*np*np  LDL^T work matrix (lower triangle)
  T *Vg, *Wg;   // B*np*kTile
  T *mask, *dv; // B*ld
  T* dnu;       // B*m
  T* dvec;      // B*ld  KKT backward: lam_lo / s_lo + lam_hi / s_hi (diagonal of G^T diag(lam / s) G)
  int* flags;   // [any_lb, any_ub] of the KKT backward (OR over the batch)
  size_t bytes;
};

template <typename T>
inline BwdWs<T> carve_bwd(void* base, int B, int n, int m) {
  BwdWs<T> w;
  w.B = B; w.n = n; w.m = m;
  w.ld = round_up(n, Vec<T>::N);
  w.np = round_up(n + m, kMacro);
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t count, size_t elt) {
    void* r = p ? p + off : nullptr;
    off = round_up_sz(off + count * elt, 256);
    return r;
  };
  const size_t Bn = (size_t)B;
  w.tc = tc_factor_enabled<T>(n, m) ? 1 : 0;
  w.nb = (n + m + kTcBlock - 1) / kTcBlock;
  if (w.tc) {
    const size_t tile_e = (size_t)kTcBlock * kTcBlock;
    w.W = (T*)take(Bn * ((size_t)w.nb * (w.nb + 1) / 2) * tile_e, sizeof(T));
    w.Vg = (T*)take(Bn * w.nb * tile_e, sizeof(T));
    w.Wg = (T*)take(Bn * w.nb * tile_e, sizeof(T));
    w.Pb = (T*)take(Bn * w.nb * tile_e, sizeof(T));
  } else {
    w.W = (T*)take(Bn * w.np * w.np, sizeof(T));
    w.Vg = (T*)take(Bn * w.np * kTile, sizeof(T));
    w.Wg = (T*)take(Bn * w.np * kTile, sizeof(T));
    w.Pb = nullptr;
  }
  w.mask = (T*)take(Bn * w.ld, sizeof(T));
  w.dv = (T*)take(Bn * w.ld, sizeof(T));
  w.dnu = (T*)take(Bn * (m > 0 ? m : 1), sizeof(T));
  w.dvec = (T*)take(Bn * w.ld, sizeof(T));
  w.flags = (int*)take(4, sizeof(int));
  w.bytes = off;
  return w;
}

// ------------------------------------------------------------------ launchers (one per .cu)
// scale.cu  -- K1: Ruiz-style equilibration, rho candidate, bound flags, state reset
template <typename T>
cudaError_t launch_scale(const lqpb_config& cfg, const FwdWs<T>& w, const T* Q, const T* p, const T* A, const T* b,
                         const T* lb, const T* ub, cudaStream_t st);

template <typename T>
cudaError_t launch_bound_flags(const FwdWs<T>& w, const T* lb, const T* ub, cudaStream_t st);

// factor.cu -- K2: inverse of the KKT matrix [[H, A^T], [A, a_diag I]] by tiled symmetric Gauss-Jordan
template <typename T>
struct GjArgs {
  int n, m, np;
  const T* src; int lds;          // B matrices, row stride lds, lower triangle read; lds == 0: src is the packed
                                  // symmetric layout (Pack<T>, diagonal stored halved)
  const T* diag_shift;            // per problem, may be null
  T diag_const;                   // added to the kept diagonal entries of H
  const T* diag_vec;              // per element (row stride ldm), added to the kept diagonal entries; may be null
  const T* mask; int ldm;         // 1 = keep, 0 = replace row/col of H by identity (and zero that column of A); may be null
  const T* Arows; int lda;        // B*m rows (row stride lda); unused when m == 0
  T a_diag;                       // diagonal of the (2,2) block
  T *W, *Vg, *Wg;                 // work: B*np*np, B*np*32, B*np*32
  T* dst; int ldd;                // K11 in the packed symmetric layout (B*Pack::elems(n)); ldd = row stride of G21 / c_out
  T* G21;                         // K21 (B*m*ldd)
  T* K22;                         // K22 (B*m*m)
  const T* bt; T* c_out;          // optional: c = K21^T b~ (B*ldd)
  // LDL-solve mode only (backward): right-hand side [-mask*rhs_g; 0], solution [sol_x (B*ldd); sol_nu (B*m)]
  const T* rhs_g; T* sol_x; T* sol_nu;
};
template <typename T>
cudaError_t launch_gj_inverse(int B, const GjArgs<T>& a, cudaStream_t st);
template <typename T>
cudaError_t launch_ldl_solve(int B, const GjArgs<T>& a, cudaStream_t st);
template <typename T>
cudaError_t launch_select_rho(const lqpb_config& cfg, const FwdWs<T>& w, cudaStream_t st);

// tcfactor.cu -- K2 on the tensor cores (fp32): same arguments / outputs as launch_gj_inverse / launch_ldl_solve,
// a.W = block-lower work matrix, a.Vg / a.Wg = panel tile buffers, Pbuf = pivot-block inverses
// prebuilt: a.W already holds the H block of the KKT matrix without the diagonal shift (written by
// scale_pack_kernel); only the shift, the equality rows and the padding are added before the sweep
cudaError_t launch_tc_inverse(int B, const GjArgs<float>& a, float* Pbuf, int nb, bool prebuilt, cudaStream_t st,
                              int* launches);
cudaError_t launch_tc_ldl_solve(int B, const GjArgs<float>& a, float* Pbuf, int nb, cudaStream_t st, int* launches,
                                int stage = 0);
cudaError_t launch_tc_dev_inverse(int B, int N, const float* A, float* Ainv, float* work, cudaStream_t st);
// f64block.cu -- the same blocked sweep in fp64 (DMMA tile products); same arguments / outputs
cudaError_t launch_tc_inverse(int B, const GjArgs<double>& a, double* Pbuf, int nb, bool prebuilt, cudaStream_t st,
                              int* launches);
cudaError_t launch_tc_ldl_solve(int B, const GjArgs<double>& a, double* Pbuf, int nb, cudaStream_t st, int* launches,
                                int stage);

// Tape of the unrolled mode: per problem and ADMM iteration k the scaled iterate x~_k, z_k, u_k ((B, n_iter, n),
// unpadded rows) and the equality part nu_k of the KKT solve ((B, n_iter, m), before the E un-scaling).
template <typename T>
struct Tape {
  int n_iter;
  T *x, *z, *u, *nu;
};

// iterate.cu -- K3 (+K4): persistent ADMM loop and finalisation (tape != nullptr: recording pass of the unrolled mode)
template <typename T>
cudaError_t launch_iterate(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                           int* launches, cudaStream_t st, const Tape<T>* tape = nullptr);

// iterate_res.cu -- K3, resident regime: the same loop with the operators held in shared memory for the whole solve
// (problems whose packed K11 fits the shared memory of the SMs the batch can use).  *taken = false: does not apply.
template <typename T>
cudaError_t launch_iterate_resident(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                                    int* launches, cudaStream_t st, bool* taken);

// iterate_row.cu -- K3, small-problem regime: dense operators in shared memory, one row per lane group
template <typename T>
cudaError_t launch_iterate_rows(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                                int* launches, cudaStream_t st, bool* taken);

// iterate_row.cu, FUSED kernel: scaling + rho + KKT inverse + ADMM loop + finalisation of small problems in one launch
// (adaptive-rho refactorisations on the device).  Needs a zeroed control block and bound_flags_kernel before it.
template <typename T>
cudaError_t launch_forward_fused(const lqpb_config& cfg, const FwdWs<T>& w, const T* Q, const T* p, const T* A, const T* b,
                                 const T* lb, const T* ub, T* x, T* z, T* u, T* lams, T* nus, T* rho_out, int* launches,
                                 cudaStream_t st, bool* taken);

template <typename T>
bool forward_fused_applies(const lqpb_config& cfg, const FwdWs<T>& w);
template <typename T>
bool iterate_rows_applies(const lqpb_config& cfg, const FwdWs<T>& w);
template <typename T>
bool iterate_resident_applies(const lqpb_config& cfg, const FwdWs<T>& w);

// unroll.cu -- reverse sweep of the unrolled mode and the rank-n_iter products that form dQ~ and dA~
template <typename T>
struct UnrollGrads {
  int k_lo, k_hi;           // iterations swept (inclusive), all with the operators / rho of `w`
  const T* gx;              // (B, n)  adjoint of x~_{k_hi} (may be null)
  const T *gz_last, *gu_last, *gzprev_last;   // (B, n) adjoints of z_{k_hi}, u_{k_hi}, z_{k_hi - 1} (may be null)
  T *gz_in, *gu_in;         // (B, n) out: adjoints of z_{k_lo - 1}, u_{k_lo - 1} (may be null)
  T *tw, *twnu;             // (B, n_iter, n), (B, n_iter, max(m,1))  adjoint solves w_k = K11 gx_k, K21 gx_k (scratch)
  T *gQ, *gp, *gA, *gb, *glb, *gub, *grho;   // adjoints of Q~, p~, A~, b~, lb~, ub~, rho (gQ / gA may be null)
};
template <typename T>
cudaError_t launch_unroll_reverse(const FwdWs<T>& w, const Tape<T>& tape, const UnrollGrads<T>& g, int* launches,
                                  cudaStream_t st);
// adjoint of Q~ = D Q D and rho = ||Q~||_F / sqrt(n): G (B,n,n) in/out, gD (B,n) out, part = (B, ceil(n/32) + 1, n) scratch
template <typename T>
cudaError_t launch_scale_grad(int B, int n, T* G, const T* Q, const T* D, const T* coef, T* gD, T* part, cudaStream_t st);
// the O(B n) part of the scaling map (:161-197) in the unrolled mode: values out of the workspace, adjoint in one kernel
template <typename T>
struct ScaleVecGrad {
  int n, m, beta_auto, use_lb, use_ub;
  T beta;
  const T *colmax, *p, *A, *b, *lb, *ub, *D, *E;          // forward inputs / outputs
  const T *gD, *gD2, *gpt, *gAt, *gbt, *glbt, *gubt;       // upstream adjoints (any may be null = zero; gD + gD2 reach D)
  T *gcolmax, *gp, *gA, *gb, *glb, *gub;                   // outputs (gA / gb null when m == 0)
};

template <typename T>
cudaError_t launch_scaled_vectors(const FwdWs<T>& w, T* D, T* pt, T* At, T* bt, T* lbt, T* ubt, T* E, cudaStream_t st);
template <typename T>
cudaError_t launch_scale_vec_grad(int B, const ScaleVecGrad<T>& a, cudaStream_t st);
template <typename T>
cudaError_t launch_colmax_plain(int B, int n, const T* Q, T* out, cudaStream_t st);
template <typename T>
cudaError_t launch_colmax_scatter(int B, int n, const T* Q, const T* colmax, const T* g, T* G, cudaStream_t st);
template <typename T>
cudaError_t launch_finalize(const FwdWs<T>& w, T* x, T* z, T* u, T* lams, T* rho_out, cudaStream_t st);

template <typename T>
cudaError_t launch_status(const lqpb_config& cfg, const FwdWs<T>& w, T* out, int* converged, cudaStream_t st);

// backward.cu -- K5/K6
template <typename T>
cudaError_t launch_bwd_mask(const BwdWs<T>& w, const T* x, const T* u, const T* lb, const T* ub, cudaStream_t st);
template <typename T>
cudaError_t launch_bwd_kkt_prep(const BwdWs<T>& w, const T* x, const T* lams, const T* lb, const T* ub, cudaStream_t st);
template <typename T>
cudaError_t launch_bwd_grads(const BwdWs<T>& w, const T* dl_dz, const T* x, const T* u, const T* lams, const T* nus,
                             const T* Q, const T* A, const T* rho_dev, double rho_scalar, T* dQ, T* dp, T* dA, T* db,
                             T* dlb, T* dub, cudaStream_t st, const T* lb_kkt = nullptr, const T* ub_kkt = nullptr);

}  // namespace lqpb
