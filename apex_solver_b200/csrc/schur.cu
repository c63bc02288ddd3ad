// schur.cu — K4 implicit Schur operator, reduced gradient, back-substitution, K6 block-PCG with
// device-resident control.
//
// sm_100a equivalents of IterativeSchurSolver (src/linalg/sparse/implicit_schur.rs):
//   apply_schur_operator_fast :163-251  -> hcc_apply_kernel + schur_tile_kernel<DC, MODE_MATVEC>
//   reduced gradient          :863-880  -> schur_tile_kernel<DC, MODE_RHS>
//   back-substitution         :923-932  -> schur_tile_kernel<DC, MODE_BACKSUB>
//   apply_preconditioner      :409-443, solve_pcg_block :577-679 -> pcg_init_kernel / pcg_step_kernel
//
// The operator is applied in ONE pass over the Jacobian planes: a CTA owns a tile of whole landmarks,
// so  t_p = sum_o Jp_o^T (Jc_o x_c(o))  is a segmented sum in shared memory, w_p = Hpp_p^-1 t_p stays in
// shared memory, and each observation then sends  -Jc_o^T (Jp_o w_p)  to y_c with FP64 reductions into L2
// (red.global.add.f64). H_cp = Jc^T Jp is never materialised: 2*(dc+3) doubles per observation are read
// instead of 3*dc.
#include "apex_ctx.h"
#include "ba_device.cuh"
#include "kernels_common.cuh"

namespace apex {

struct SchurArgs {
  const TileDesc* tiles;
  const uint32_t* slot_cam;
  const uint16_t* slot_lp;
  const uint32_t* pt_slot0;
  const uint32_t* pt_cnt;
  const double* J;
  const double* hinv;
  const double* gp;
  const double* x;
  double* y;
  double* step_pt;
  uint32_t npl;
  int check_done;
  const DevState* st;
};

template <int DC>
__device__ __forceinline__ void load_slot_jacobian(const double* Jt, double* jc, double* jp) {
#pragma unroll
  for (int k = 0; k < 2 * DC; ++k) jc[k] = ld_stream(Jt + (size_t)k * TILE);
#pragma unroll
  for (int k = 0; k < 6; ++k) jp[k] = ld_stream(Jt + (size_t)(2 * DC + k) * TILE);
}

// u = Jp^T (Jc x_c)
template <int DC>
__device__ __forceinline__ void obs_forward(const double* jc, const double* jp, const double* __restrict__ xc, double u[3]) {
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int k = 0; k < DC; ++k) {
    const double xv = __ldg(xc + k);
    a0 = fma(jc[k], xv, a0);
    a1 = fma(jc[DC + k], xv, a1);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) u[k] = fma(jp[k], a0, jp[3 + k] * a1);
}

// y_c -= Jc^T (Jp w)
template <int DC>
__device__ __forceinline__ void obs_backward(const double* jc, const double* jp, const double w[3], double* yc) {
  const double b0 = fma(jp[0], w[0], fma(jp[1], w[1], jp[2] * w[2]));
  const double b1 = fma(jp[3], w[0], fma(jp[4], w[1], jp[5] * w[2]));
#pragma unroll
  for (int k = 0; k < DC; ++k) red_add(yc + k, -fma(jc[k], b0, jc[DC + k] * b1));
}

// per-landmark middle step; returns w (MATVEC / RHS) or writes the landmark step (BACKSUB)
template <int MODE>
__device__ __forceinline__ void landmark_middle(const SchurArgs& a, uint32_t lp, const double t[3], double w[3]) {
  const size_t n = a.npl;
  const double h00 = a.hinv[0 * n + lp], h01 = a.hinv[1 * n + lp], h02 = a.hinv[2 * n + lp];
  const double h11 = a.hinv[3 * n + lp], h12 = a.hinv[4 * n + lp], h22 = a.hinv[5 * n + lp];
  double v[3];
  if (MODE == MODE_MATVEC) { v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; }
  else {
    const double g0 = -a.gp[0 * n + lp], g1 = -a.gp[1 * n + lp], g2 = -a.gp[2 * n + lp];
    if (MODE == MODE_RHS) { v[0] = g0; v[1] = g1; v[2] = g2; }
    else { v[0] = g0 - t[0]; v[1] = g1 - t[1]; v[2] = g2 - t[2]; }
  }
  w[0] = h00 * v[0] + h01 * v[1] + h02 * v[2];
  w[1] = h01 * v[0] + h11 * v[1] + h12 * v[2];
  w[2] = h02 * v[0] + h12 * v[1] + h22 * v[2];
  if (MODE == MODE_BACKSUB) {
    a.step_pt[3 * (size_t)lp] = w[0]; a.step_pt[3 * (size_t)lp + 1] = w[1]; a.step_pt[3 * (size_t)lp + 2] = w[2];
  }
}

template <int DC, int MODE>
__global__ void __launch_bounds__(TILE) schur_tile_kernel(SchurArgs a) {
  constexpr int NP = 2 * (DC + 3);
  if (a.check_done && a.st->pcg_done) return;
  __shared__ double sh[3][TILE];
  __shared__ double shw[3][TILE];
  const TileDesc td = a.tiles[blockIdx.x];
  const int tid = threadIdx.x;
  if (td.nchunks == 1) {
    const size_t chunk = td.chunk0;
    const size_t slot = chunk * TILE + tid;
    const uint32_t cam = a.slot_cam[slot];
    double jc[2 * DC], jp[6];
    if (cam != PAD_CAM) load_slot_jacobian<DC>(a.J + chunk * NP * TILE + tid, jc, jp);
    if (MODE != MODE_RHS) {
      double u[3] = {0.0, 0.0, 0.0};
      if (cam != PAD_CAM) obs_forward<DC>(jc, jp, a.x + (size_t)cam * DC, u);
      sh[0][tid] = u[0]; sh[1][tid] = u[1]; sh[2][tid] = u[2];
      __syncthreads();
    }
    if ((uint32_t)tid < td.npt) {
      const uint32_t lp = td.pt0 + tid;
      double t[3] = {0.0, 0.0, 0.0}, w[3];
      if (MODE != MODE_RHS) {
        const uint32_t off = a.pt_slot0[lp] - td.chunk0 * TILE, cnt = a.pt_cnt[lp];
        for (uint32_t q = 0; q < cnt; ++q) { t[0] += sh[0][off + q]; t[1] += sh[1][off + q]; t[2] += sh[2][off + q]; }
      }
      landmark_middle<MODE>(a, lp, t, w);
      if (MODE != MODE_BACKSUB) { shw[0][tid] = w[0]; shw[1][tid] = w[1]; shw[2][tid] = w[2]; }
    }
    if (MODE == MODE_BACKSUB) return;
    __syncthreads();
    if (cam != PAD_CAM) {
      const uint32_t li = a.slot_lp[slot];
      const double w[3] = {shw[0][li], shw[1][li], shw[2][li]};
      obs_backward<DC>(jc, jp, w, a.y + (size_t)cam * DC);
    }
  } else {
    // one landmark spread over several chunks: accumulate per thread, reduce in fixed order, re-read J
    double u[3] = {0.0, 0.0, 0.0};
    if (MODE != MODE_RHS) {
      for (uint32_t ch = 0; ch < td.nchunks; ++ch) {
        const size_t chunk = (size_t)td.chunk0 + ch;
        const uint32_t cam = a.slot_cam[chunk * TILE + tid];
        if (cam == PAD_CAM) continue;
        double jc[2 * DC], jp[6], uu[3];
        load_slot_jacobian<DC>(a.J + chunk * NP * TILE + tid, jc, jp);
        obs_forward<DC>(jc, jp, a.x + (size_t)cam * DC, uu);
        u[0] += uu[0]; u[1] += uu[1]; u[2] += uu[2];
      }
    }
    double t[3];
    t[0] = block_reduce_sum(u[0], &sh[0][0]);
    t[1] = block_reduce_sum(u[1], &sh[0][0]);
    t[2] = block_reduce_sum(u[2], &sh[0][0]);
    if (tid == 0) {
      double w[3];
      landmark_middle<MODE>(a, td.pt0, t, w);
      shw[0][0] = w[0]; shw[1][0] = w[1]; shw[2][0] = w[2];
    }
    if (MODE == MODE_BACKSUB) return;
    __syncthreads();
    const double w[3] = {shw[0][0], shw[1][0], shw[2][0]};
    for (uint32_t ch = 0; ch < td.nchunks; ++ch) {
      const size_t chunk = (size_t)td.chunk0 + ch;
      const uint32_t cam = a.slot_cam[chunk * TILE + tid];
      if (cam == PAD_CAM) continue;
      double jc[2 * DC], jp[6];
      load_slot_jacobian<DC>(a.J + chunk * NP * TILE + tid, jc, jp);
      obs_backward<DC>(jc, jp, w, a.y + (size_t)cam * DC);
    }
  }
}

// y = (H_cc + lambda I) x on the block diagonal (first half of apply_schur_operator_fast); sign = -1 with
// x = g_c gives the start value -g_c of the reduced gradient when `hcc` is null.
__global__ void hcc_apply_kernel(const double* __restrict__ hcc, const double* __restrict__ x, double* __restrict__ y, const DevState* st,
                                 uint32_t n, int dc, int check_done) {
  if (check_done && st->pcg_done) return;
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const uint32_t cam = row / dc, a = row % dc;
  const double* H = hcc + ((size_t)cam * dc + a) * dc;
  const double* xc = x + (size_t)cam * dc;
  double s = st->damping * xc[a];
  for (int b = 0; b < dc; ++b) s += H[b] * xc[b];
  y[row] = s;
}

__global__ void negate_kernel(const double* __restrict__ x, double* __restrict__ y, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = -x[i];
}

// z = M^-1 r, block diagonal per camera variable (apply_preconditioner, implicit_schur.rs:409-443)
__device__ __forceinline__ double precond_row(const double* __restrict__ pinv, const double* __restrict__ r, uint32_t row, int dc, int K) {
  const uint32_t cam = row / dc, a = row % dc;
  const double* P = pinv + (size_t)cam * (36 + K * K);
  const double* rc = r + (size_t)cam * dc;
  double s = 0.0;
  if (a < 6) {
    for (int b = 0; b < 6; ++b) s += P[a * 6 + b] * rc[b];
  } else {
    const double* Q = P + 36 + (a - 6) * K;
    for (int b = 0; b < K; ++b) s += Q[b] * rc[6 + b];
  }
  return s;
}

// solve_pcg_block set-up (implicit_schur.rs:577-600): x0 = 0, r = b, z = M^-1 r, p = z
__global__ void __launch_bounds__(1024) pcg_init_kernel(const double* __restrict__ b, const double* __restrict__ pinv, double* x, double* r,
                                                        double* z, double* p, DevState* st, uint32_t n, int dc, int K, int max_it, double cg_tol) {
  __shared__ double sh[1024];
  double bb = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) { const double v = b[i]; r[i] = v; x[i] = 0.0; bb += v * v; }
  __syncthreads();
  double rz = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) { const double zv = precond_row(pinv, b, i, dc, K); z[i] = zv; p[i] = zv; rz += b[i] * zv; }
  bb = block_reduce_sum(bb, sh);
  rz = block_reduce_sum(rz, sh);
  if (threadIdx.x == 0) {
    const double b_norm = sqrt(bb);
    st->b_norm = b_norm;
    st->pcg_tol = cg_tol * dmax(b_norm, 1.0);
    st->rz_old = rz;
    st->r_norm = b_norm;
    st->pcg_iters = 0;
    st->pcg_max = max_it;
    st->pcg_done = max_it <= 0 ? 1 : 0;
  }
}

// one PCG iteration after ap = S p is complete (implicit_schur.rs:604-676), single CTA, deterministic
__global__ void __launch_bounds__(1024) pcg_step_kernel(const double* __restrict__ ap, const double* __restrict__ pinv, double* x, double* r,
                                                        double* z, double* p, DevState* st, uint32_t n, int dc, int K) {
  __shared__ double sh[1024];
  if (st->pcg_done) return;
  const int iters = st->pcg_iters + 1;
  const double rz_old = st->rz_old, tol = st->pcg_tol;
  const int max_it = st->pcg_max;
  double v = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) v = fma(p[i], ap[i], v);
  const double p_ap = block_reduce_sum(v, sh);
  if (fabs(p_ap) < 1e-20) {
    if (threadIdx.x == 0) { st->pcg_iters = iters; st->pcg_done = 1; }
    return;
  }
  const double alpha = rz_old / p_ap;
  v = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) {
    x[i] += alpha * p[i];
    const double rv = r[i] - alpha * ap[i];
    r[i] = rv;
    v = fma(rv, rv, v);
  }
  const double r_norm = sqrt(block_reduce_sum(v, sh));  // also orders the r writes before the z reads
  if (r_norm < tol) {
    if (threadIdx.x == 0) { st->pcg_iters = iters; st->pcg_done = 1; st->r_norm = r_norm; }
    return;
  }
  v = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) { const double zv = precond_row(pinv, r, i, dc, K); z[i] = zv; v = fma(r[i], zv, v); }
  const double rz_new = block_reduce_sum(v, sh);
  if (fabs(rz_old) < 1e-30) {
    if (threadIdx.x == 0) { st->pcg_iters = iters; st->pcg_done = 1; st->r_norm = r_norm; }
    return;
  }
  const double beta = rz_new / rz_old;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) p[i] = z[i] + beta * p[i];
  if (threadIdx.x == 0) {
    st->rz_old = rz_new;
    st->r_norm = r_norm;
    st->pcg_iters = iters;
    if (iters >= max_it) st->pcg_done = 1;
  }
}

// ----------------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------------
static SchurArgs make_schur_args(Ctx& c, const double* x, double* y, int check_done) {
  SchurArgs a;
  a.tiles = c.tiles.p; a.slot_cam = c.slot_cam.p; a.slot_lp = c.slot_lp.p; a.pt_slot0 = c.pt_slot0.p; a.pt_cnt = c.pt_cnt.p;
  a.J = c.J.p; a.hinv = c.hinv.p; a.gp = c.gp.p; a.x = x; a.y = y; a.step_pt = c.step_pt.p;
  a.npl = c.npl; a.check_done = check_done; a.st = c.state.p;
  return a;
}

template <int DC>
static void launch_tiles_dc(Ctx& c, int mode, const SchurArgs& a) {
  switch (mode) {
    case MODE_MATVEC: schur_tile_kernel<DC, MODE_MATVEC><<<c.ntiles, TILE, 0, c.stream>>>(a); break;
    case MODE_RHS: schur_tile_kernel<DC, MODE_RHS><<<c.ntiles, TILE, 0, c.stream>>>(a); break;
    default: schur_tile_kernel<DC, MODE_BACKSUB><<<c.ntiles, TILE, 0, c.stream>>>(a); break;
  }
}

apex_status launch_schur_tiles(Ctx& c, int mode, const double* x, double* y, int check_done) {
  if (c.ntiles == 0) return APEX_OK;
  SchurArgs a = make_schur_args(c, x, y, check_done);
  switch (c.dc) {
    case 6: launch_tiles_dc<6>(c, mode, a); break;
    case 9: launch_tiles_dc<9>(c, mode, a); break;
    case 10: launch_tiles_dc<10>(c, mode, a); break;
    case 12: launch_tiles_dc<12>(c, mode, a); break;
    case 14: launch_tiles_dc<14>(c, mode, a); break;
    default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
  }
  c.launches++;
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

// y = (H_cc + lambda I) x on rank 0, 0 elsewhere (the all-reduce that follows the tile kernel adds it once)
apex_status launch_hcc_apply(Ctx& c, const double* x, double* y, int check_done) {
  const uint32_t n = c.ncam * c.dc;
  if (c.rank == 0) {
    hcc_apply_kernel<<<(n + 255) / 256, 256, 0, c.stream>>>(c.hcc.p, x, y, c.state.p, n, c.dc, check_done);
    c.launches++;
    APEX_CUDA_TRY(c, cudaGetLastError());
  } else {
    APEX_CUDA_TRY(c, cudaMemsetAsync(y, 0, (size_t)n * sizeof(double), c.stream));
  }
  return APEX_OK;
}

// full operator y = S x (all ranks end with the same y)
apex_status schur_operator(Ctx& c, const double* x, double* y, int check_done) {
  APEX_TRY(launch_hcc_apply(c, x, y, check_done));
  APEX_TRY(launch_schur_tiles(c, MODE_MATVEC, x, y, check_done));
  APEX_TRY(allreduce_sum(c, y, (size_t)c.ncam * c.dc));
  return APEX_OK;
}

// b = -g_c - H_cp Hpp^-1 (-g_p)   (implicit_schur.rs:863-880 with g = -J^T r; explicit_schur.rs:928-977)
apex_status launch_reduced_gradient(Ctx& c, double* b) {
  const uint32_t n = c.ncam * c.dc;
  if (c.rank == 0) {
    negate_kernel<<<(n + 255) / 256, 256, 0, c.stream>>>(c.gc, b, n);
    c.launches++;
    APEX_CUDA_TRY(c, cudaGetLastError());
  } else {
    APEX_CUDA_TRY(c, cudaMemsetAsync(b, 0, (size_t)n * sizeof(double), c.stream));
  }
  APEX_TRY(launch_schur_tiles(c, MODE_RHS, nullptr, b, 0));
  APEX_TRY(allreduce_sum(c, b, n));
  return APEX_OK;
}

// IterativeSchurSolver::solve_with_cached_hessian (implicit_schur.rs:835-946) on the current linearization.
// Leaves the camera step in c.step_cam and the landmark step in c.step_pt.
apex_status solve_implicit(Ctx& c, int precond, int cg_max_it, double cg_tol) {
  const uint32_t n = c.ncam * c.dc;
  cudaStream_t s = c.stream;
  APEX_TRY(launch_reduced_gradient(c, c.vb.p));
  APEX_TRY(launch_schur_jacobi_blocks(c, precond));
  pcg_init_kernel<<<1, 1024, 0, s>>>(c.vb.p, c.pinv.p, c.step_cam.p, c.vr.p, c.vz.p, c.vp.p, c.state.p, n, c.dc, c.K, cg_max_it, cg_tol);
  c.launches++;
  APEX_CUDA_TRY(c, cudaGetLastError());
  // PCG: iterations are enqueued in batches; every kernel of an iteration is a no-op once the device-side
  // `pcg_done` flag is set, so the host only polls the flag between batches.
  const int BATCH = 10;
  int enq = 0;
  while (enq < cg_max_it) {
    const int nb = std::min(BATCH, cg_max_it - enq);
    for (int i = 0; i < nb; ++i) {
      APEX_TRY(schur_operator(c, c.vp.p, c.vy.p, 1));
      pcg_step_kernel<<<1, 1024, 0, s>>>(c.vy.p, c.pinv.p, c.step_cam.p, c.vr.p, c.vz.p, c.vp.p, c.state.p, n, c.dc, c.K);
      c.launches++;
    }
    APEX_CUDA_TRY(c, cudaGetLastError());
    enq += nb;
    APEX_TRY(sync_state(c));
    if (c.h_state->pcg_done) break;
  }
  if (cg_max_it <= 0) APEX_TRY(sync_state(c));
  c.last_pcg_iters = c.h_state->pcg_iters;
  if (c.h_state->singular_landmark) { c.err = "Landmark block singular"; return APEX_ERR_SINGULAR_MATRIX; }
  APEX_TRY(launch_schur_tiles(c, MODE_BACKSUB, c.step_cam.p, nullptr, 0));
  return APEX_OK;
}

}  // namespace apex
