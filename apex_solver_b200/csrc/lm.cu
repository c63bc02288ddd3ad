// lm.cu — K9 manifold update / revert, K10 Levenberg–Marquardt bookkeeping on the device, and the loop.
//
// sm_100a equivalents of: apply_parameter_step / apply_negative_parameter_step (src/optimizer/mod.rs:309-356 ->
// src/core/problem.rs:185-289), compute_parameter_norm (mod.rs:458-467), compute_predicted_reduction /
// update_damping (src/optimizer/levenberg_marquardt.rs:721-727, 702-717), compute_step_quality and
// check_convergence (mod.rs:668-675, 591-658), optimize_with_mode (levenberg_marquardt.rs:823-1028).
// Damping, nu, costs, rho, the accept flag and the termination status live in DevState; the host only
// enqueues kernels and reads the status word once per LM iteration.
#include <chrono>
#include <cmath>

#include "apex_ctx.h"
#include "ba_device.cuh"
#include "kernels_common.cuh"

namespace apex {

// ---- norms of gradient and step ------------------------------------------------------------------------
// camera side: one CTA (ncam*dc elements, replicated on all ranks)
__global__ void __launch_bounds__(1024) cam_norms_kernel(const double* __restrict__ gc, const double* __restrict__ sc, uint32_t n, DevState* st) {
  __shared__ double sh[1024];
  double g2 = 0.0, s2 = 0.0, sg = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) { const double g = gc[i], s = sc[i]; g2 += g * g; s2 += s * s; sg += s * g; }
  g2 = block_reduce_sum(g2, sh); s2 = block_reduce_sum(s2, sh); sg = block_reduce_sum(sg, sh);
  if (threadIdx.x == 0) { st->g2_cam = g2; st->s2_cam = s2; st->sg_cam = sg; }
}

// landmark side: per-CTA partials over the local landmarks, then a fixed-order final sum
__global__ void __launch_bounds__(256) pt_norms_kernel(const double* __restrict__ gp, const double* __restrict__ sp, uint32_t npl, double* partial,
                                                       uint32_t nblocks) {
  __shared__ double sh[256];
  const uint32_t lp = blockIdx.x * 256 + threadIdx.x;
  double g2 = 0.0, s2 = 0.0, sg = 0.0;
  if (lp < npl) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double g = gp[(size_t)k * npl + lp], s = sp[3 * (size_t)lp + k]; g2 += g * g; s2 += s * s; sg += s * g; }
  }
  g2 = block_reduce_sum(g2, sh); s2 = block_reduce_sum(s2, sh); sg = block_reduce_sum(sg, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x] = g2; partial[nblocks + blockIdx.x] = s2; partial[2 * (size_t)nblocks + blockIdx.x] = sg; }
}

__global__ void __launch_bounds__(1024) sum3_kernel(const double* __restrict__ partial, uint32_t nblocks, double* out3) {
  __shared__ double sh[1024];
  for (int k = 0; k < 3; ++k) {
    double v = 0.0;
    for (uint32_t i = threadIdx.x; i < nblocks; i += 1024) v += partial[(size_t)k * nblocks + i];
    v = block_reduce_sum(v, sh);
    if (threadIdx.x == 0) out3[k] = v;
  }
}

apex_status launch_step_norms(Ctx& c) {
  cudaStream_t s = c.stream;
  // (shared intrinsics: gradient and step of the reduced layout [poses | one intrinsics block], left there by solve_explicit)
  if (c.shared_intr) { const uint32_t nsh = 6 * c.ncam + (uint32_t)c.K; cam_norms_kernel<<<1, 1024, 0, s>>>(c.sh_vec.p, c.sh_vec.p + nsh, nsh, c.state.p); }
  else cam_norms_kernel<<<1, 1024, 0, s>>>(c.gc, c.step_cam.p, c.ncam * c.dc, c.state.p);
  const uint32_t nb = (c.npl + 255) / 256;
  if (nb) pt_norms_kernel<<<nb, 256, 0, s>>>(c.gp.p, c.step_pt.p, c.npl, c.red_scratch.p, nb);
  sum3_kernel<<<1, 1024, 0, s>>>(c.red_scratch.p, nb, &c.state.p->g2_pt);
  c.launches += 2 + (nb ? 1 : 0);
  APEX_CUDA_TRY(c, cudaGetLastError());
  APEX_TRY(allreduce_sum(c, &c.state.p->g2_pt, 3));
  return APEX_OK;
}

// ---- manifold update ----------------------------------------------------------------------------------
// x (+) sign*step with the fixed tangent indices zeroed (problem.rs:185-289); SE3 right-plus, Rn addition
__global__ void apply_step_cam_kernel(double* pose, double* intr, const double* __restrict__ step, const uint8_t* __restrict__ pose_fixed,
                                      const uint16_t* __restrict__ intr_fixed, uint32_t ncam, int dc, int K, int opt_intr, double sign,
                                      const DevState* st, int only_if_rejected) {
  if (only_if_rejected && st->accepted) return;
  const uint32_t cam = blockIdx.x * blockDim.x + threadIdx.x;
  if (cam >= ncam) return;
  const uint8_t fx = pose_fixed[cam];
  double tau[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) tau[a] = ((fx >> a) & 1) ? 0.0 : sign * step[(size_t)cam * dc + a];
  double* P = pose + 7 * (size_t)cam;
  Pose p;  // the stored quaternion is not renormalised between updates
  p.t = {P[0], P[1], P[2]};
  p.q = {P[3], P[4], P[5], P[6]};
  const Pose np = pose_plus(p, tau);
  P[0] = np.t.x; P[1] = np.t.y; P[2] = np.t.z; P[3] = np.q.w; P[4] = np.q.i; P[5] = np.q.j; P[6] = np.q.k;
  if (opt_intr) {
    const uint16_t fi = intr_fixed[cam];
    for (int a = 0; a < K; ++a) {
      const double d = ((fi >> a) & 1) ? 0.0 : sign * step[(size_t)cam * dc + 6 + a];
      intr[(size_t)cam * K + a] += d;
    }
  }
}

__global__ void apply_step_pt_kernel(double* pt, const double* __restrict__ step, const uint8_t* __restrict__ pt_fixed, uint32_t npl, double sign,
                                     const DevState* st, int only_if_rejected) {
  if (only_if_rejected && st->accepted) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * (size_t)npl) return;
  const uint32_t lp = (uint32_t)(i / 3);
  const int a = (int)(i % 3);
  const double d = ((pt_fixed[lp] >> a) & 1) ? 0.0 : sign * step[i];
  pt[i] += d;
}

apex_status launch_apply_step(Ctx& c, double sign, bool only_if_rejected) {
  cudaStream_t s = c.stream;
  apply_step_cam_kernel<<<(c.ncam + 127) / 128, 128, 0, s>>>(c.pose.p, c.intr.p, c.step_cam.p, c.pose_fixed.p, c.intr_fixed.p, c.ncam, c.dc, c.K,
                                                             c.opt_intr ? 1 : 0, sign, c.state.p, only_if_rejected ? 1 : 0);
  c.launches++;
  const size_t n3 = 3 * (size_t)c.npl;
  if (n3) {
    apply_step_pt_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, s>>>(c.pt.p, c.step_pt.p, c.pt_fixed.p, c.npl, sign, c.state.p,
                                                                        only_if_rejected ? 1 : 0);
    c.launches++;
  }
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

// ---- compute_parameter_norm (mod.rs:458-467): SE3 contributes its 7-vector ----------------------------
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const double* __restrict__ v, size_t n, double* partial) {
  __shared__ double sh[256];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s += v[i] * v[i];
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) cam_param_norm_kernel(const double* __restrict__ pose, const double* __restrict__ intr, uint32_t ncam, int K,
                                                              int with_intr, DevState* st) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (size_t i = threadIdx.x; i < 7 * (size_t)ncam; i += 1024) s += pose[i] * pose[i];
  if (with_intr)   // with_intr == 2: one shared intrinsics variable (every camera holds a copy of it)
    for (size_t i = threadIdx.x; i < (size_t)K * (with_intr == 2 ? 1 : ncam); i += 1024) s += intr[i] * intr[i];
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) st->pn2_cam = s;
}

apex_status launch_param_norm(Ctx& c) {
  cudaStream_t s = c.stream;
  cam_param_norm_kernel<<<1, 1024, 0, s>>>(c.pose.p, c.intr.p, c.ncam, c.K, c.shared_intr ? 2 : ((c.opt_intr || c.intr_vars) ? 1 : 0), c.state.p);
  const size_t n3 = 3 * (size_t)c.npl;
  const unsigned nb = (unsigned)std::min<size_t>((n3 + 255) / 256, 1024);
  if (nb) sumsq_partial_kernel<<<nb, 256, 0, s>>>(c.pt.p, n3, c.red_scratch.p);
  reduce_sum_kernel<<<1, 1024, 0, s>>>(c.red_scratch.p, (size_t)nb, &c.state.p->pn2_pt);
  c.launches += 2 + (nb ? 1 : 0);
  APEX_CUDA_TRY(c, cudaGetLastError());
  APEX_TRY(allreduce_sum(c, &c.state.p->pn2_pt, 1));
  return APEX_OK;
}

// ---- LM bookkeeping (single thread; every scalar stays on the device) ---------------------------------
struct LmParams {
  int32_t max_iterations;
  double cost_tolerance, parameter_tolerance, gradient_tolerance;
  double timeout_seconds, damping_min, damping_max;
  double min_cost_threshold, trust_region_radius, min_trust_region_radius;
};

__global__ void lm_init_kernel(DevState* st, double damping, double nu) {
  const double cn = sqrt(st->cost2_local);  // compute_cost = 0.5 * norm_l2(r)^2 (mod.rs:358-361)
  const double c0 = 0.5 * cn * cn;
  st->damping = damping; st->nu = nu;
  st->current_cost = c0; st->previous_cost = c0; st->new_cost = c0;
  st->accepted = 0; st->status = -1; st->iteration = 0;
  st->singular_landmark = 0; st->chol_fail = 0;
}

// after the trial step and its cost: predicted reduction, rho, damping update, accept/reject
__global__ void lm_eval_kernel(DevState* st, LmParams p) {
  const double g2 = st->g2_cam + st->g2_pt, s2 = st->s2_cam + st->s2_pt, sg = st->sg_cam + st->sg_pt;
  const double grad_norm = sqrt(g2), step_norm = sqrt(s2);
  // compute_predicted_reduction (levenberg_marquardt.rs:721-727): 0.5 * step^T (damping*step - gradient)
  const double predicted = 0.5 * (st->damping * step_norm * step_norm - sg);
  const double cn = sqrt(st->cost2_local);
  const double new_cost = 0.5 * cn * cn;
  // compute_step_quality (mod.rs:668-675)
  const double actual = st->current_cost - new_cost;
  double rho;
  if (fabs(predicted) < 1e-15) rho = actual > 0.0 ? 1.0 : 0.0; else rho = actual / predicted;
  // update_damping (levenberg_marquardt.rs:702-717)
  int accepted;
  if (rho > 0.0) {
    const double coff = 2.0 * rho - 1.0;
    double d = st->damping * dmax(1.0 / 3.0, 1.0 - coff * coff * coff);
    d = dmax(d, p.damping_min);
    st->damping = d; st->nu = 2.0; accepted = 1;
  } else {
    double d = st->damping * st->nu;
    st->nu = st->nu * 2.0;
    d = dmin(d, p.damping_max);
    st->damping = d; accepted = 0;
  }
  st->grad_norm = grad_norm; st->step_norm = step_norm; st->step_dot_grad = sg;
  st->predicted = predicted; st->rho = rho; st->new_cost = new_cost; st->accepted = accepted;
}

// after accept/revert and the parameter norm: trace row + check_convergence (mod.rs:591-658)
__global__ void lm_converge_kernel(DevState* st, LmParams p, apex_iter_trace* trace, int trace_cap, double elapsed, int pcg_iters_hint) {
  if (elapsed < 0.0) elapsed = st->elapsed;  // nranks > 1: rank 0's clock, all-reduced, so that every rank takes the same decision
  const int iteration = st->iteration;
  const int accepted = st->accepted;
  const double new_cost = st->new_cost;
  double cost_reduction = 0.0;
  double current = st->current_cost;
  if (accepted) { cost_reduction = current - new_cost; current = new_cost; }
  const double pnorm = sqrt(st->pn2_cam + st->pn2_pt);
  st->param_norm = pnorm;
  if (trace && iteration < trace_cap) {
    apex_iter_trace& t = trace[iteration];
    t.iteration = iteration; t.accepted = accepted; t.ls_iter = pcg_iters_hint >= 0 ? pcg_iters_hint : st->pcg_iters; t.reserved = 0;
    t.cost = current; t.cost_change = st->previous_cost - current; t.gradient_norm = st->grad_norm; t.step_norm = st->step_norm;
    t.tr_ratio = st->rho; t.tr_radius = st->damping; t.new_cost = new_cost; t.predicted_reduction = st->predicted;
    t.parameter_norm = pnorm; t.iter_time_ms = 0.0;
  }
  st->previous_cost = current;
  st->current_cost = current;
  const double cost_before = accepted ? current + cost_reduction : current;  // levenberg_marquardt.rs:946-950
  // check_convergence: new_cost argument = state.current_cost (levenberg_marquardt.rs:952-969)
  int status = -1;
  const double cc = cost_before, nc = current, snorm = st->step_norm, gnorm = st->grad_norm;
  if (!isfinite(nc) || !isfinite(snorm) || !isfinite(gnorm)) status = APEX_STATUS_INVALID_NUMERICAL_VALUES;
  else if (p.timeout_seconds > 0.0 && elapsed >= p.timeout_seconds) status = APEX_STATUS_TIMEOUT;
  else if (iteration >= p.max_iterations) status = APEX_STATUS_MAX_ITERATIONS_REACHED;
  else if (accepted) {
    if (gnorm < p.gradient_tolerance) status = APEX_STATUS_GRADIENT_TOLERANCE_REACHED;
    if (status < 0 && iteration > 0) {
      const double rel_step_tol = p.parameter_tolerance * (pnorm + p.parameter_tolerance);
      if (snorm <= rel_step_tol) status = APEX_STATUS_PARAMETER_TOLERANCE_REACHED;
      else {
        const double cost_change = fabs(cc - nc);
        const double rel = cost_change / dmax(cc, 1e-10);
        if (rel < p.cost_tolerance) status = APEX_STATUS_COST_TOLERANCE_REACHED;
      }
    }
    if (status < 0 && !isnan(p.min_cost_threshold) && nc < p.min_cost_threshold) status = APEX_STATUS_MIN_COST_THRESHOLD_REACHED;
    if (status < 0 && p.trust_region_radius < p.min_trust_region_radius) status = APEX_STATUS_TRUST_REGION_RADIUS_TOO_SMALL;
  }
  st->status = status;
  st->iteration = iteration + 1;
}

// Jacobi column scaling (process_jacobian_generic, src/optimizer/mod.rs:749-763): the column norms of J are the square roots of
// the diagonals of the camera / landmark blocks of J^T J of the UNSCALED linearisation; scaling = 1 / (1 + norm)
__global__ void jacobi_scale_kernel(const double* __restrict__ hcc, const double* __restrict__ hpp, double* __restrict__ sc, double* __restrict__ sp,
                                    uint32_t ncam, int dc, uint32_t npl) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t ncd = (size_t)ncam * dc;
  if (i < ncd) sc[i] = 1.0 / (1.0 + sqrt(hcc[i * dc + i % dc]));
  else if (i < ncd + (size_t)npl * 3) {
    const size_t j = i - ncd, lp = j / 3;
    const int k = (int)(j % 3);
    const int plane = k == 0 ? 0 : (k == 1 ? 3 : 5);   // H_pp planes: 00 01 02 11 12 22
    sp[j] = 1.0 / (1.0 + sqrt(hpp[(size_t)plane * npl + lp]));
  }
}
// step = scaled_step .* scaling (apply_inverse_scaling, src/linearizer/mod.rs:255-261)
__global__ void jacobi_unscale_step_kernel(double* __restrict__ step_cam, double* __restrict__ step_pt, const double* __restrict__ sc,
                                           const double* __restrict__ sp, size_t ncd, size_t np3) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncd) step_cam[i] *= sc[i];
  else if (i < ncd + np3) step_pt[i - ncd] *= sp[i - ncd];
}

static double now_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// LevenbergMarquardt::optimize -> optimize_with_mode (levenberg_marquardt.rs:1034-1083, 823-1028)
apex_status lm_solve(Ctx& c, const apex_lm_config* cfg, apex_lm_result* res, apex_iter_trace* trace, int trace_cap) {
  if (!c.have_problem) { c.err = "no problem uploaded"; return APEX_ERR_INVALID_STATE; }
  if (c.nobs == 0) { c.err = "no residual blocks"; return APEX_ERR_NO_RESIDUAL_BLOCKS; }
  // compute_covariances: accepted and without effect, like the reference on this path - the Schur solvers do not implement
  // LinearSolver::compute_covariance_matrix (default None, src/linalg/mod.rs:170-172), so SolverResult::covariances is None
  c.jacobi_on = false;
  struct ScalingOff { Ctx& c; ~ScalingOff() { if (c.jacobi_on) { c.jacobi_on = false; c.linearized = false; } } } scaling_off{c};  // standalone entry points stay unscaled
  if (cfg->schur_variant < APEX_SCHUR_EXPLICIT || cfg->schur_variant > APEX_SCHUR_EXPLICIT_PCG) { c.err = "bad schur_variant"; return APEX_ERR_INVALID_PARAMETERS; }
  if (c.shared_intr && (cfg->schur_variant != APEX_SCHUR_EXPLICIT || cfg->use_jacobi_scaling)) { c.err = "shared intrinsics: explicit Schur + Cholesky without column scaling only"; return APEX_ERR_UNSUPPORTED; }
  cudaStream_t s = c.stream;
  const double t0 = now_seconds();
  LmParams p{cfg->max_iterations, cfg->cost_tolerance, cfg->parameter_tolerance, cfg->gradient_tolerance, cfg->timeout_seconds,
             cfg->damping_min, cfg->damping_max, cfg->min_cost_threshold, cfg->trust_region_radius, cfg->min_trust_region_radius};
  if (trace_cap < 0) trace_cap = 0;
  APEX_CUDA_TRY(c, c.trace.alloc((size_t)std::max(trace_cap, 1)));
  std::vector<double> iter_ms;

  if (!c.ev_lm0) { cudaEventCreate(&c.ev_lm0); cudaEventCreate(&c.ev_lm1); }
  cudaEventRecord(c.ev_lm0, s);
  c.lm_timed = false;
  // Variable::new(SE3::from(DVector)) normalises the initial quaternions (src/core/problem.rs:743-757)
  APEX_TRY(launch_normalize_poses(c));
  APEX_TRY(launch_cost(c, nullptr));  // initial cost (mod.rs:550-552)
  lm_init_kernel<<<1, 1, 0, s>>>(c.state.p, cfg->damping, cfg->damping_nu);
  c.launches++;
  APEX_TRY(sync_state(c));
  const double initial_cost = c.h_state->current_cost;
  int cost_evals = 1, jac_evals = 0, ok_steps = 0, bad_steps = 0;
  int64_t lin_iters = 0;

  for (int iteration = 0;; ++iteration) {
    const double it0 = now_seconds();
    if (cfg->use_jacobi_scaling && iteration == 0) {
      // the scaling comes from the unscaled Jacobian of the first iterate and stays fixed (mod.rs:754-758); the linearisation
      // is then repeated with it (the once-per-solve price of never storing an unscaled copy)
      const size_t ncd = (size_t)c.ncam * c.dc, np3 = (size_t)c.npl * 3;
      APEX_CUDA_TRY(c, c.scale_cam.alloc(ncd));
      APEX_CUDA_TRY(c, c.scale_pt.alloc(std::max<size_t>(np3, 1)));
      APEX_TRY(launch_linearize(c));
      jacobi_scale_kernel<<<(unsigned)((ncd + np3 + 255) / 256), 256, 0, s>>>(c.hcc.p, c.hpp.p, c.scale_cam.p, c.scale_pt.p, c.ncam, c.dc, c.npl);
      c.launches++;
      APEX_CUDA_TRY(c, cudaGetLastError());
      c.jacobi_on = true;
    }
    APEX_TRY(launch_linearize(c));
    c.linearized = true;
    jac_evals++;
    apex_status st;
    if (cfg->schur_variant == APEX_SCHUR_IMPLICIT) st = solve_implicit(c, cfg->schur_preconditioner, cfg->cg_max_iterations, cfg->cg_tolerance);
    else st = solve_explicit(c, cfg->schur_variant == APEX_SCHUR_EXPLICIT_PCG, cfg->cg_max_iterations, cfg->cg_tolerance);
    if (st != APEX_OK) return (st == APEX_ERR_SINGULAR_MATRIX || st == APEX_ERR_FACTORIZATION_FAILED) ? APEX_ERR_LINEAR_SOLVE_FAILED : st;
    if (c.jacobi_on) {
      // the predicted reduction pairs the UNSCALED step with the SCALED gradient (compute_step_generic, levenberg_marquardt.rs:738-761)
      const size_t ncd = (size_t)c.ncam * c.dc, np3 = (size_t)c.npl * 3;
      jacobi_unscale_step_kernel<<<(unsigned)((ncd + np3 + 255) / 256), 256, 0, s>>>(c.step_cam.p, c.step_pt.p, c.scale_cam.p, c.scale_pt.p, ncd, np3);
      c.launches++;
    }
    c.have_step = true;
    const int pcg_it = (int)c.last_pcg_iters;
    lin_iters += pcg_it;
    APEX_TRY(launch_step_norms(c));
    APEX_TRY(launch_apply_step(c, +1.0, false));  // evaluate_and_apply_step (:770-817)
    APEX_TRY(launch_cost(c, nullptr));
    cost_evals++;
    lm_eval_kernel<<<1, 1, 0, s>>>(c.state.p, p);
    c.launches++;
    APEX_TRY(launch_apply_step(c, -1.0, true));   // apply_negative_parameter_step when rejected (:804-808)
    APEX_TRY(launch_param_norm(c));
    double elapsed = now_seconds() - t0;
    if (c.nranks > 1 && p.timeout_seconds > 0.0) {
      // the TIMEOUT test must not depend on each rank's own clock: a rank that returned TIMEOUT while its peers enter the
      // next linearisation would leave them blocked in ncclAllReduce. Everybody uses rank 0's elapsed time.
      const double mine = c.rank == 0 ? elapsed : 0.0;
      APEX_CUDA_TRY(c, cudaMemcpyAsync(&c.state.p->elapsed, &mine, sizeof(double), cudaMemcpyHostToDevice, s));
      APEX_CUDA_TRY(c, cudaStreamSynchronize(s));  // `mine` is a stack variable
      APEX_TRY(allreduce_sum(c, &c.state.p->elapsed, 1));
      elapsed = -1.0;
    }
    lm_converge_kernel<<<1, 1, 0, s>>>(c.state.p, p, trace ? c.trace.p : nullptr, trace_cap, elapsed, pcg_it);
    c.launches++;
    APEX_CUDA_TRY(c, cudaGetLastError());
    APEX_TRY(sync_state(c));
    c.linearized = false;  // the variables moved (or lambda changed): the cached linearization is stale
    iter_ms.push_back((now_seconds() - it0) * 1e3);
    if (c.h_state->accepted) ok_steps++; else bad_steps++;
    if (!c.observers.empty()) {
      // notify_observers_generic (src/optimizer/mod.rs:728-743; levenberg_marquardt.rs:930-940): the metric tuple of the iteration
      // just finished, after the accept / reject decision. The stream is idle here (sync_state), so a callback may read the
      // variables with apex_params_download.
      apex_observer_metrics m;
      m.iteration = iteration; m.accepted = c.h_state->accepted; m.cost = c.h_state->current_cost; m.gradient_norm = c.h_state->grad_norm;
      m.damping = c.h_state->damping; m.step_norm = c.h_state->step_norm; m.step_quality = c.h_state->rho;
      for (size_t i = 0; i < c.observers.size(); ++i) { const apex_observer o = c.observers[i]; if (o.on_step) o.on_step(o.user, c.self, &m); }
    }
    if (c.h_state->status >= 0) {
      cudaEventRecord(c.ev_lm1, s);
      c.lm_timed = true;
      res->status = c.h_state->status;
      res->iterations = iteration + 1;
      res->initial_cost = initial_cost;
      res->final_cost = c.h_state->current_cost;
      res->elapsed_seconds = now_seconds() - t0;
      res->final_gradient_norm = c.h_state->grad_norm;
      res->final_parameter_update_norm = c.h_state->step_norm;
      res->cost_evaluations = cost_evals;
      res->jacobian_evaluations = jac_evals;
      res->successful_steps = ok_steps;
      res->unsuccessful_steps = bad_steps;
      res->final_damping = c.h_state->damping;
      res->final_damping_nu = c.h_state->nu;
      res->linear_iterations = lin_iters;
      if (trace && trace_cap > 0) {
        const int nrow = std::min(iteration + 1, trace_cap);
        APEX_CUDA_TRY(c, cudaMemcpyAsync(trace, c.trace.p, (size_t)nrow * sizeof(apex_iter_trace), cudaMemcpyDeviceToHost, s));
        APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
        for (int i = 0; i < nrow; ++i) trace[i].iter_time_ms = iter_ms[i];
      }
      // notify_complete(&final_parameters, iteration + 1) (levenberg_marquardt.rs:1010-1011)
      for (size_t i = 0; i < c.observers.size(); ++i) { const apex_observer o = c.observers[i]; if (o.on_optimization_complete) o.on_optimization_complete(o.user, c.self, iteration + 1); }
      return APEX_OK;
    }
  }
}

}  // namespace apex
