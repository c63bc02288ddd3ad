// linearize.cu — K1 linearize (+ landmark-side J^T J / J^T r and the damped 3x3 inverses), K1' cost,
// K2 camera-side block accumulation, K5 Schur-Jacobi / block-diagonal preconditioner blocks.
//
// Replaces, on sm_100a: AssemblyBackend::assemble (src/linearizer/mod.rs:216-227 ->
// src/linearizer/cpu/sparse.rs:119-184 -> linearize_block src/linearizer/mod.rs:117-175), the block
// part of H = J^T J, g = J^T r (src/linalg/sparse/explicit_schur.rs:1146-1166),
// invert_landmark_blocks (explicit_schur.rs:365-442 / implicit_schur.rs:685-778),
// compute_residual_sparse + compute_cost (src/core/problem.rs:864-899, src/optimizer/mod.rs:358-361) and
// compute_schur_jacobi_preconditioner / compute_block_preconditioner (implicit_schur.rs:456-573, 352-404).
#include <cstdlib>

#include "apex_ctx.h"
#include "ba_device.cuh"
#include "kernels_common.cuh"

namespace apex {

// ----------------------------------------------------------------------------------------------------
// SE3::from(DVector) at variable creation (src/core/problem.rs:743-757): normalise the quaternions once
// ----------------------------------------------------------------------------------------------------
__global__ void normalize_poses_kernel(double* pose, uint32_t ncam) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncam) return;
  Pose p = pose_from7(pose + 7 * (size_t)i);
  pose[7 * (size_t)i + 3] = p.q.w; pose[7 * (size_t)i + 4] = p.q.i; pose[7 * (size_t)i + 5] = p.q.j; pose[7 * (size_t)i + 6] = p.q.k;
}

apex_status launch_normalize_poses(Ctx& c) {
  normalize_poses_kernel<<<(c.ncam + 127) / 128, 128, 0, c.stream>>>(c.pose.p, c.ncam);
  c.launches++;
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

// ----------------------------------------------------------------------------------------------------
// K1: one CTA per tile, one thread per observation slot
// ----------------------------------------------------------------------------------------------------
struct LinArgs {
  const TileDesc* tiles;
  const uint32_t* slot_cam;
  const uint16_t* slot_lp;
  const double* slot_uv;
  const uint32_t* pt_slot0;
  const uint32_t* pt_cnt;
  const uint2* cslot_meta;   // normal chunks: [15:8] of .y = camera-sorted lane of the observation at this point-major lane
  const double* pose;
  const double* intr;
  const double* pt;
  double* J;
  double* R;
  double* hpp;
  double* gp;
  double* hinv;
  double* tile_partial;
  uint32_t npl;
  LossSpec loss;
  const uint8_t* slot_loss;   // per-block loss functions: index into loss_tab per slot, or null (uniform `loss`)
  const LossSpec* loss_tab;
  const double* scale_cam;    // Jacobi column scaling [ncam][dc] / [npl][3], or null (off)
  const double* scale_pt;
  DevState* st;
};

// J * diag(scaling) for one block (apply_column_scaling, src/linearizer/mod.rs:240-252): camera columns by the camera's
// scaling row, landmark columns by the landmark's
template <int DC>
__device__ __forceinline__ void scale_block(double* jc, double* jp, const double* __restrict__ sc, const double* __restrict__ sp) {
#pragma unroll
  for (int k = 0; k < DC; ++k) { const double f = sc[k]; jc[k] *= f; jc[DC + k] *= f; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { const double f = sp[k]; jp[k] *= f; jp[3 + k] *= f; }
}

// landmark block: store H_pp, g_p; damp with lambda*I (explicit_schur.rs:1208-1212, implicit_schur.rs:1053-1068);
// guarded inverse (explicit_schur.rs:377-442)
__device__ __forceinline__ void finish_landmark(const LinArgs& a, uint32_t lp, const double s[9], double lambda) {
  const size_t n = a.npl;
#pragma unroll
  for (int k = 0; k < 6; ++k) a.hpp[k * n + lp] = s[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) a.gp[k * n + lp] = s[6 + k];
  double blk[9] = {s[0] + lambda, s[1], s[2], s[1], s[3] + lambda, s[4], s[2], s[4], s[5] + lambda};
  double inv[9];
  if (!invert_landmark_block(blk, inv)) {
    atomicExch(&a.st->singular_landmark, 1);
    for (int k = 0; k < 9; ++k) inv[k] = 0.0;
  }
  a.hinv[0 * n + lp] = inv[0]; a.hinv[1 * n + lp] = inv[1]; a.hinv[2 * n + lp] = inv[2];
  a.hinv[3 * n + lp] = inv[4]; a.hinv[4 * n + lp] = inv[5]; a.hinv[5 * n + lp] = inv[8];
}

// 3 CTAs per SM (80 registers, ~300 B of spills on the BAL model) instead of the 128 registers the compiler takes when left
// alone: the kernel is latency bound (gathers behind the slot metadata, FP64 dependency chains), so 24 instead of 16 resident
// warps more than pay for the spills - the linearisation group went from 2.38 to 1.12 ms on the Venice shape.
template <int MODEL, bool OPT_INTR>
__global__ void __launch_bounds__(TILE, 3) linearize_tile_kernel(LinArgs a) {
  constexpr int K = CamK<MODEL>::K;
  constexpr int DC = 6 + (OPT_INTR ? K : 0);
  constexpr int NP = 2 * (DC + 3);
  __shared__ double sh[9][TILE];
  const TileDesc td = a.tiles[blockIdx.x];
  const int tid = threadIdx.x;
  const double lambda = a.st->damping;
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;

  for (uint32_t ch = 0; ch < td.nchunks; ++ch) {
    const size_t chunk = (size_t)td.chunk0 + ch;
    const size_t slot = chunk * TILE + tid;
    const uint32_t cam = a.slot_cam[slot];
    double r[2] = {0.0, 0.0}, jall[NP];
    double* jc = jall;
    double* jp = jall + 2 * DC;
    if (cam != PAD_CAM) {
      const uint32_t lp = td.pt0 + a.slot_lp[slot];
      const Pose pose = pose_from7(a.pose + 7 * (size_t)cam);
      double in[K];
#pragma unroll
      for (int k = 0; k < K; ++k) in[k] = a.intr[(size_t)cam * K + k];
      const V3 pw{a.pt[3 * (size_t)lp], a.pt[3 * (size_t)lp + 1], a.pt[3 * (size_t)lp + 2]};
      const double u = a.slot_uv[(chunk * 2 + 0) * TILE + tid], v = a.slot_uv[(chunk * 2 + 1) * TILE + tid];
      linearize_obs<MODEL, OPT_INTR, true>(a.slot_loss ? a.loss_tab[a.slot_loss[slot]] : a.loss, pose, in, pw, u, v, r, jc, jp);
      if (a.scale_cam) scale_block<DC>(jc, jp, a.scale_cam + (size_t)cam * DC, a.scale_pt + 3 * (size_t)lp);
    } else {
#pragma unroll
      for (int k = 0; k < 2 * DC; ++k) jc[k] = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) jp[k] = 0.0;
    }
    // split slot order (apex_ctx.h): the landmark half stays at this lane, the camera half goes to the observation's
    // camera-sorted lane inside the chunk (padding lanes are the same set in both orders); a landmark with more than 256
    // observations keeps both halves at its own lane
    const int clane = (td.nchunks == 1 && cam != PAD_CAM) ? (int)((__ldg(&a.cslot_meta[slot].y) >> 8) & 0xFFu) : tid;
    store_jacobian_split<DC>(a.J, chunk, clane, tid, jall);
    a.R[(chunk * 2 + 0) * TILE + tid] = r[0];
    a.R[(chunk * 2 + 1) * TILE + tid] = r[1];
    // landmark-side contributions: upper triangle of Jp^T Jp, then Jp^T r
    double cv[9];
    cv[0] = jp[0] * jp[0] + jp[3] * jp[3];
    cv[1] = jp[0] * jp[1] + jp[3] * jp[4];
    cv[2] = jp[0] * jp[2] + jp[3] * jp[5];
    cv[3] = jp[1] * jp[1] + jp[4] * jp[4];
    cv[4] = jp[1] * jp[2] + jp[4] * jp[5];
    cv[5] = jp[2] * jp[2] + jp[5] * jp[5];
    cv[6] = jp[0] * r[0] + jp[3] * r[1];
    cv[7] = jp[1] * r[0] + jp[4] * r[1];
    cv[8] = jp[2] * r[0] + jp[5] * r[1];
    if (td.nchunks == 1) {
#pragma unroll
      for (int k = 0; k < 9; ++k) sh[k][tid] = cv[k];
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[k] += cv[k];
    }
  }
  if (td.nchunks == 1) {
    __syncthreads();
    if ((uint32_t)tid < td.npt) {
      const uint32_t lp = td.pt0 + tid;
      const uint32_t off = a.pt_slot0[lp] - td.chunk0 * TILE, cnt = a.pt_cnt[lp];
      double s[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        double v = 0.0;
        for (uint32_t q = 0; q < cnt; ++q) v += sh[k][off + q];  // insertion order, as the CPU path
        s[k] = v;
      }
      finish_landmark(a, lp, s, lambda);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) sh[k][tid] = acc[k];
    __syncthreads();
    for (int s = TILE / 2; s > 0; s >>= 1) {
      if (tid < s) {
#pragma unroll
        for (int k = 0; k < 9; ++k) sh[k][tid] += sh[k][tid + s];
      }
      __syncthreads();
    }
    if (tid == 0) {
      double s[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) s[k] = sh[k][0];
      finish_landmark(a, td.pt0, s, lambda);
    }
  }
}

// K1': residual only -> per-tile sum of r~^2
template <int MODEL, bool OPT_INTR>
__global__ void __launch_bounds__(TILE) cost_tile_kernel(LinArgs a) {
  constexpr int K = CamK<MODEL>::K;
  __shared__ double sh[TILE];
  const TileDesc td = a.tiles[blockIdx.x];
  const int tid = threadIdx.x;
  double acc = 0.0;
  for (uint32_t ch = 0; ch < td.nchunks; ++ch) {
    const size_t chunk = (size_t)td.chunk0 + ch;
    const size_t slot = chunk * TILE + tid;
    const uint32_t cam = a.slot_cam[slot];
    if (cam != PAD_CAM) {
      const uint32_t lp = td.pt0 + a.slot_lp[slot];
      const Pose pose = pose_from7(a.pose + 7 * (size_t)cam);
      double in[K];
#pragma unroll
      for (int k = 0; k < K; ++k) in[k] = a.intr[(size_t)cam * K + k];
      const V3 pw{a.pt[3 * (size_t)lp], a.pt[3 * (size_t)lp + 1], a.pt[3 * (size_t)lp + 2]};
      const double u = a.slot_uv[(chunk * 2 + 0) * TILE + tid], v = a.slot_uv[(chunk * 2 + 1) * TILE + tid];
      double r[2];
      linearize_obs<MODEL, OPT_INTR, false>(a.slot_loss ? a.loss_tab[a.slot_loss[slot]] : a.loss, pose, in, pw, u, v, r, nullptr, nullptr);
      acc += r[0] * r[0] + r[1] * r[1];
    }
  }
  sh[tid] = acc;
  __syncthreads();
  for (int s = TILE / 2; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  if (tid == 0) a.tile_partial[blockIdx.x] = sh[0];
}

// ----------------------------------------------------------------------------------------------------
// K2 / K5: camera-major accumulation. One CTA per (camera, <=CAM_CHUNK observations) work item; every
// thread re-evaluates the projection of its observations and keeps the block sums in registers;
// fixed-order shuffle + shared-memory reduction; work-item partials are summed in item order by the
// finalize kernels => deterministic, no global atomics.
//   WHAT = 0: H_cc (upper triangle of Jc^T Jc) and g_c = Jc^T r       (explicit_schur.rs:1146-1160)
//   WHAT = 1: sum_j H_cp[i,j] H_pp[j]^-1 H_cp[i,j]^T per camera VARIABLE block (pose 6x6, intrinsics KxK)
//             (implicit_schur.rs:456-573)
// ----------------------------------------------------------------------------------------------------
static_assert(sizeof(LossSpec) == sizeof(LossSpecPod), "loss table entries are uploaded as LossSpecPod");
struct CamArgs {
  const CamItem* items;
  const uint32_t* cam_item_start;
  const double* cm_uv;
  const uint32_t* cm_lp;
  const double* pose;
  const double* intr;
  const double* pt;
  const double* hinv;
  double* partial;
  double* hcc;
  double* gc;
  double* sj;
  uint32_t npl;
  uint32_t ncam;
  uint64_t nobs_local;
  LossSpec loss;
  const uint8_t* cm_loss;     // per-block loss functions, camera-major order, or null
  const LossSpec* loss_tab;
  const double* scale_cam;    // Jacobi column scaling, or null
  const double* scale_pt;
};

template <int MODEL, bool OPT_INTR, int WHAT>
struct CamAcc {
  static constexpr int K = CamK<MODEL>::K;
  static constexpr int DC = 6 + (OPT_INTR ? K : 0);
  static constexpr int NH = DC * (DC + 1) / 2;
  static constexpr int NI = OPT_INTR ? K * (K + 1) / 2 : 0;
  static constexpr int NACC = WHAT == 0 ? NH + DC : 21 + NI;
};

template <int MODEL, bool OPT_INTR, int WHAT>
__global__ void __launch_bounds__(CAM_THREADS) camera_accum_kernel(CamArgs a) {
  using CA = CamAcc<MODEL, OPT_INTR, WHAT>;
  constexpr int K = CA::K, DC = CA::DC, NH = CA::NH, NACC = CA::NACC;
  __shared__ double shw[CAM_THREADS / 32][NACC];
  const CamItem item = a.items[blockIdx.x];
  const int tid = threadIdx.x;
  const Pose pose = pose_from7(a.pose + 7 * (size_t)item.cam);
  double in[K];
#pragma unroll
  for (int k = 0; k < K; ++k) in[k] = a.intr[(size_t)item.cam * K + k];
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  for (uint32_t o = item.begin + tid; o < item.end; o += CAM_THREADS) {
    const uint32_t lp = a.cm_lp[o];
    const double u = a.cm_uv[o], v = a.cm_uv[a.nobs_local + o];
    const V3 pw{a.pt[3 * (size_t)lp], a.pt[3 * (size_t)lp + 1], a.pt[3 * (size_t)lp + 2]};
    double r[2], jc[2 * DC], jp[6];
    linearize_obs<MODEL, OPT_INTR, true>(a.cm_loss ? a.loss_tab[a.cm_loss[o]] : a.loss, pose, in, pw, u, v, r, jc, jp);
    if (a.scale_cam) scale_block<DC>(jc, jp, a.scale_cam + (size_t)item.cam * DC, a.scale_pt + 3 * (size_t)lp);
    if constexpr (WHAT == 0) {
      int idx = 0;
#pragma unroll
      for (int p = 0; p < DC; ++p)
#pragma unroll
        for (int q = p; q < DC; ++q) { acc[idx] += jc[p] * jc[q] + jc[DC + p] * jc[DC + q]; ++idx; }
#pragma unroll
      for (int p = 0; p < DC; ++p) acc[NH + p] += jc[p] * r[0] + jc[DC + p] * r[1];
    } else {
      const size_t n = a.npl;
      const double h00 = a.hinv[0 * n + lp], h01 = a.hinv[1 * n + lp], h02 = a.hinv[2 * n + lp];
      const double h11 = a.hinv[3 * n + lp], h12 = a.hinv[4 * n + lp], h22 = a.hinv[5 * n + lp];
      double E[DC][3], T[DC][3];
#pragma unroll
      for (int p = 0; p < DC; ++p) {
#pragma unroll
        for (int k = 0; k < 3; ++k) E[p][k] = jc[p] * jp[k] + jc[DC + p] * jp[3 + k];
        T[p][0] = E[p][0] * h00 + E[p][1] * h01 + E[p][2] * h02;
        T[p][1] = E[p][0] * h01 + E[p][1] * h11 + E[p][2] * h12;
        T[p][2] = E[p][0] * h02 + E[p][1] * h12 + E[p][2] * h22;
      }
      int idx = 0;
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = p; q < 6; ++q) { acc[idx] += T[p][0] * E[q][0] + T[p][1] * E[q][1] + T[p][2] * E[q][2]; ++idx; }
      if constexpr (OPT_INTR) {
#pragma unroll
        for (int p = 0; p < K; ++p)
#pragma unroll
          for (int q = p; q < K; ++q) { acc[idx] += T[6 + p][0] * E[6 + q][0] + T[6 + p][1] * E[6 + q][1] + T[6 + p][2] * E[6 + q][2]; ++idx; }
      }
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    double v = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) shw[warp][k] = v;
  }
  __syncthreads();
  for (int k = tid; k < NACC; k += CAM_THREADS) {
    double v = shw[0][k];
#pragma unroll
    for (int w = 1; w < CAM_THREADS / 32; ++w) v += shw[w][k];
    a.partial[(size_t)blockIdx.x * NACC + k] = v;
  }
}

// one thread per (camera, accumulator): sum the work-item partials in item order, expand symmetric storage
template <int DC>
__global__ void camera_finalize_hcc_kernel(CamArgs a) {
  constexpr int NH = DC * (DC + 1) / 2, NACC = NH + DC;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)a.ncam * NACC) return;
  const uint32_t cam = (uint32_t)(gid / NACC);
  const int k = (int)(gid % NACC);
  double v = 0.0;
  for (uint32_t it = a.cam_item_start[cam]; it < a.cam_item_start[cam + 1]; ++it) v += a.partial[(size_t)it * NACC + k];
  if (k >= NH) { a.gc[(size_t)cam * DC + (k - NH)] = v; return; }
  int p = 0, rem = k;
  while (rem >= DC - p) { rem -= DC - p; ++p; }
  const int q = p + rem;
  a.hcc[((size_t)cam * DC + p) * DC + q] = v;
  a.hcc[((size_t)cam * DC + q) * DC + p] = v;
}

template <int K, bool OPT_INTR>
__global__ void camera_finalize_sj_kernel(CamArgs a) {
  constexpr int NI = OPT_INTR ? K * (K + 1) / 2 : 0, NACC = 21 + NI, PB = 36 + K * K;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)a.ncam * NACC) return;
  const uint32_t cam = (uint32_t)(gid / NACC);
  const int k = (int)(gid % NACC);
  double v = 0.0;
  for (uint32_t it = a.cam_item_start[cam]; it < a.cam_item_start[cam + 1]; ++it) v += a.partial[(size_t)it * NACC + k];
  if (k < 21) {
    int p = 0, rem = k;
    while (rem >= 6 - p) { rem -= 6 - p; ++p; }
    const int q = p + rem;
    a.sj[(size_t)cam * PB + p * 6 + q] = v;
    a.sj[(size_t)cam * PB + q * 6 + p] = v;
  } else {
    int p = 0, rem = k - 21;
    while (rem >= K - p) { rem -= K - p; ++p; }
    const int q = p + rem;
    a.sj[(size_t)cam * PB + 36 + p * K + q] = v;
    a.sj[(size_t)cam * PB + 36 + q * K + p] = v;
  }
}

// Preconditioner block inverses (implicit_schur.rs:352-404, 456-573): per camera variable block
// (H_cc[i,i] + lambda I [- Schur-Jacobi subtrahend])^-1; on a singular block add max(1e-6|trace|/n, 1e-8) and
// retry, else identity. kind NONE -> identity.
template <int K>
__global__ void precond_invert_kernel(const double* hcc, const double* sj, double* pinv, const DevState* st, uint32_t ncam, int dc,
                                      int kind, int opt_intr) {
  constexpr int PB = 36 + K * K;
  const uint32_t cam = blockIdx.x * blockDim.x + threadIdx.x;
  if (cam >= ncam) return;
  const double lambda = st->damping;
  const double* H = hcc + (size_t)cam * dc * dc;
  const double* sub = sj + (size_t)cam * PB;
  double* out = pinv + (size_t)cam * PB;
  double blk[64], inv[64];
  for (int which = 0; which < 2; ++which) {
    const int n = which == 0 ? 6 : K, o0 = which == 0 ? 0 : 6, base = which == 0 ? 0 : 36;
    if (which == 1 && !opt_intr) {
      for (int a = 0; a < K * K; ++a) out[36 + a] = 0.0;
      break;
    }
    if (kind == APEX_PRECOND_NONE) {
      for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) out[base + a * n + b] = a == b ? 1.0 : 0.0;
      continue;
    }
    for (int a = 0; a < n; ++a)
      for (int b = 0; b < n; ++b) {
        double v = H[(o0 + a) * dc + o0 + b] + (a == b ? lambda : 0.0);
        if (kind == APEX_PRECOND_SCHUR_JACOBI) v -= sub[base + a * n + b];
        blk[a * n + b] = v;
      }
    double work[64];
    for (int a = 0; a < n * n; ++a) work[a] = blk[a];
    bool ok = inverse_n<8>(n, work, inv);
    if (!ok) {
      double trace = 0.0;
      for (int a = 0; a < n; ++a) trace += blk[a * n + a];
      const double reg = dmax(1e-6 * fabs(trace) / (double)n, 1e-8);
      for (int a = 0; a < n * n; ++a) work[a] = blk[a];
      for (int a = 0; a < n; ++a) work[a * n + a] += reg;
      ok = inverse_n<8>(n, work, inv);
      if (!ok) for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) inv[a * n + b] = a == b ? 1.0 : 0.0;
    }
    for (int a = 0; a < n * n; ++a) out[base + a] = inv[a];
  }
}

// ----------------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------------
static LinArgs make_lin_args(Ctx& c) {
  LinArgs a;
  a.tiles = c.tiles.p; a.slot_cam = c.slot_cam.p; a.slot_lp = c.slot_lp.p; a.slot_uv = c.slot_uv.p;
  a.pt_slot0 = c.pt_slot0.p; a.pt_cnt = c.pt_cnt.p; a.cslot_meta = c.cslot_meta.p;
  a.pose = c.pose.p; a.intr = c.intr.p; a.pt = c.pt.p;
  a.J = c.J.p; a.R = c.R.p; a.hpp = c.hpp.p; a.gp = c.gp.p; a.hinv = c.hinv.p;
  a.tile_partial = c.red_scratch.p;
  a.npl = c.npl;
  a.loss = {c.loss_id, c.loss_p[0], c.loss_p[1]};
  a.slot_loss = c.per_obs_loss ? c.slot_loss.p : nullptr; a.loss_tab = reinterpret_cast<const LossSpec*>(c.loss_tab.p);
  a.scale_cam = c.jacobi_on ? c.scale_cam.p : nullptr; a.scale_pt = c.scale_pt.p;
  a.st = c.state.p;
  return a;
}

static CamArgs make_cam_args(Ctx& c) {
  CamArgs a;
  a.items = c.items.p; a.cam_item_start = c.cam_item_start.p; a.cm_uv = c.cm_uv.p; a.cm_lp = c.cm_lp.p;
  a.pose = c.pose.p; a.intr = c.intr.p; a.pt = c.pt.p; a.hinv = c.hinv.p;
  a.partial = c.partial.p; a.hcc = c.hcc.p; a.gc = c.gc; a.sj = c.sj.p;
  a.npl = c.npl; a.ncam = c.ncam; a.nobs_local = c.nobs_local;
  a.loss = {c.loss_id, c.loss_p[0], c.loss_p[1]};
  a.cm_loss = c.per_obs_loss ? c.cm_loss.p : nullptr; a.loss_tab = reinterpret_cast<const LossSpec*>(c.loss_tab.p);
  a.scale_cam = c.jacobi_on ? c.scale_cam.p : nullptr; a.scale_pt = c.scale_pt.p;
  return a;
}

#define APEX_DISPATCH_MODEL(c, CALL)                                                        \
  switch ((c).model) {                                                                      \
    case APEX_CAM_BAL: if ((c).opt_intr) { CALL(APEX_CAM_BAL, true); } else { CALL(APEX_CAM_BAL, false); } break; \
    case APEX_CAM_PINHOLE: if ((c).opt_intr) { CALL(APEX_CAM_PINHOLE, true); } else { CALL(APEX_CAM_PINHOLE, false); } break; \
    case APEX_CAM_KANNALA_BRANDT: if ((c).opt_intr) { CALL(APEX_CAM_KANNALA_BRANDT, true); } else { CALL(APEX_CAM_KANNALA_BRANDT, false); } break; \
    case APEX_CAM_DOUBLE_SPHERE: if ((c).opt_intr) { CALL(APEX_CAM_DOUBLE_SPHERE, true); } else { CALL(APEX_CAM_DOUBLE_SPHERE, false); } break; \
    case APEX_CAM_RADTAN: if ((c).opt_intr) { CALL(APEX_CAM_RADTAN, true); } else { CALL(APEX_CAM_RADTAN, false); } break; \
    case APEX_CAM_UCM: if ((c).opt_intr) { CALL(APEX_CAM_UCM, true); } else { CALL(APEX_CAM_UCM, false); } break; \
    case APEX_CAM_EUCM: if ((c).opt_intr) { CALL(APEX_CAM_EUCM, true); } else { CALL(APEX_CAM_EUCM, false); } break; \
    case APEX_CAM_FOV: if ((c).opt_intr) { CALL(APEX_CAM_FOV, true); } else { CALL(APEX_CAM_FOV, false); } break; \
    case APEX_CAM_FTHETA: if ((c).opt_intr) { CALL(APEX_CAM_FTHETA, true); } else { CALL(APEX_CAM_FTHETA, false); } break; \
    default: (c).err = "camera model not supported"; return APEX_ERR_UNSUPPORTED;           \
  }

apex_status launch_linearize(Ctx& c) {
  cudaStream_t s = c.stream;
  cudaEvent_t* evp = c.prof ? prof_pair(c.ev_lin, c.ev_lin_used++) : nullptr;
  if (evp) cudaEventRecord(evp[0], s);
  LinArgs la = make_lin_args(c);
  CamArgs ca = make_cam_args(c);
  if (c.ntiles) {
#define CALL(M, OI) linearize_tile_kernel<M, OI><<<c.ntiles, TILE, 0, s>>>(la)
    APEX_DISPATCH_MODEL(c, CALL)
#undef CALL
    c.launches++;
  }
  if (c.nitems) {
#define CALL(M, OI) camera_accum_kernel<M, OI, 0><<<c.nitems, CAM_THREADS, 0, s>>>(ca)
    APEX_DISPATCH_MODEL(c, CALL)
#undef CALL
    c.launches++;
  }
  {
    const int dc = c.dc;
    const size_t total = (size_t)c.ncam * (dc * (dc + 1) / 2 + dc);
    const unsigned grid = (unsigned)((total + 255) / 256);
    switch (dc) {
      case 6: camera_finalize_hcc_kernel<6><<<grid, 256, 0, s>>>(ca); break;
      case 9: camera_finalize_hcc_kernel<9><<<grid, 256, 0, s>>>(ca); break;
      case 10: camera_finalize_hcc_kernel<10><<<grid, 256, 0, s>>>(ca); break;
      case 12: camera_finalize_hcc_kernel<12><<<grid, 256, 0, s>>>(ca); break;
      case 11: camera_finalize_hcc_kernel<11><<<grid, 256, 0, s>>>(ca); break;
      case 14: camera_finalize_hcc_kernel<14><<<grid, 256, 0, s>>>(ca); break;
      case 15: camera_finalize_hcc_kernel<15><<<grid, 256, 0, s>>>(ca); break;
      default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
    }
    c.launches++;
  }
  if (evp) cudaEventRecord(evp[1], s);
  APEX_CUDA_TRY(c, cudaGetLastError());
  // camera-side blocks are sums over all ranks' observations
  APEX_TRY(allreduce_sum(c, c.hcc.p, (size_t)c.ncam * c.dc * (c.dc + 1)));
  APEX_TRY(agree_error_flags(c));  // a singular landmark block in one shard is everybody's error
  return APEX_OK;
}

apex_status launch_cost(Ctx& c, double*) {
  cudaStream_t s = c.stream;
  LinArgs la = make_lin_args(c);
  if (c.ntiles) {
#define CALL(M, OI) cost_tile_kernel<M, OI><<<c.ntiles, TILE, 0, s>>>(la)
    APEX_DISPATCH_MODEL(c, CALL)
#undef CALL
    c.launches++;
  }
  reduce_sum_kernel<<<1, 1024, 0, s>>>(c.red_scratch.p, (size_t)c.ntiles, &c.state.p->cost2_local);
  c.launches++;
  APEX_CUDA_TRY(c, cudaGetLastError());
  APEX_TRY(allreduce_sum(c, &c.state.p->cost2_local, 1));
  return APEX_OK;
}

apex_status launch_schur_jacobi_blocks(Ctx& c, int kind) {
  cudaStream_t s = c.stream;
  CamArgs ca = make_cam_args(c);
  const int K = c.K;
  if (kind == APEX_PRECOND_SCHUR_JACOBI) {
    if (c.nitems) {
#define CALL(M, OI) camera_accum_kernel<M, OI, 1><<<c.nitems, CAM_THREADS, 0, s>>>(ca)
      APEX_DISPATCH_MODEL(c, CALL)
#undef CALL
      c.launches++;
    }
    const size_t total = (size_t)c.ncam * (21 + (c.opt_intr ? K * (K + 1) / 2 : 0));
    const unsigned grid = (unsigned)((total + 255) / 256);
#define CALLF(KK)                                                                  \
  if (c.opt_intr) camera_finalize_sj_kernel<KK, true><<<grid, 256, 0, s>>>(ca);    \
  else camera_finalize_sj_kernel<KK, false><<<grid, 256, 0, s>>>(ca)
    switch (K) {
      case 3: CALLF(3); break;
      case 4: CALLF(4); break;
      case 6: CALLF(6); break;
      case 5: CALLF(5); break;
      case 8: CALLF(8); break;
      case 9: CALLF(9); break;
      default: c.err = "unsupported K"; return APEX_ERR_UNSUPPORTED;
    }
#undef CALLF
    c.launches++;
    APEX_CUDA_TRY(c, cudaGetLastError());
    APEX_TRY(allreduce_sum(c, c.sj.p, (size_t)c.ncam * (36 + K * K)));
  }
  const unsigned grid = (c.ncam + 63) / 64;
  switch (K) {
    case 3: precond_invert_kernel<3><<<grid, 64, 0, s>>>(c.hcc.p, c.sj.p, c.pinv.p, c.state.p, c.ncam, c.dc, kind, c.opt_intr); break;
    case 4: precond_invert_kernel<4><<<grid, 64, 0, s>>>(c.hcc.p, c.sj.p, c.pinv.p, c.state.p, c.ncam, c.dc, kind, c.opt_intr); break;
    case 6: precond_invert_kernel<6><<<grid, 64, 0, s>>>(c.hcc.p, c.sj.p, c.pinv.p, c.state.p, c.ncam, c.dc, kind, c.opt_intr); break;
    case 5: precond_invert_kernel<5><<<grid, 64, 0, s>>>(c.hcc.p, c.sj.p, c.pinv.p, c.state.p, c.ncam, c.dc, kind, c.opt_intr); break;
    case 8: precond_invert_kernel<8><<<grid, 64, 0, s>>>(c.hcc.p, c.sj.p, c.pinv.p, c.state.p, c.ncam, c.dc, kind, c.opt_intr); break;
    case 9: precond_invert_kernel<9><<<grid, 64, 0, s>>>(c.hcc.p, c.sj.p, c.pinv.p, c.state.p, c.ncam, c.dc, kind, c.opt_intr); break;
    default: c.err = "unsupported K"; return APEX_ERR_UNSUPPORTED;
  }
  c.launches++;
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

}  // namespace apex
