// schur.cu — K4 implicit Schur operator, reduced gradient, back-substitution, K6 block-PCG with
// device-resident control.
//
// sm_100a equivalents of IterativeSchurSolver (src/linalg/sparse/implicit_schur.rs):
//   apply_schur_operator_fast :163-251  -> schur_chunk_kernel<DC, MODE_MATVEC> (schur_tile_kernel for landmarks with more
//                                          than 256 observations)
//   reduced gradient          :863-880  -> schur_chunk_kernel<DC, MODE_RHS>
//   back-substitution         :923-932  -> schur_chunk_kernel<DC, MODE_BACKSUB>
//   apply_preconditioner      :409-443, solve_pcg_block :577-679 -> pcg_init / pcg_dir_hcc / pcg_pap / pcg_update kernels
//
// The operator is applied in ONE pass over the Jacobian planes: a CTA owns a chunk of whole landmarks,
// so  t_p = sum_o Jp_o^T (Jc_o x_c(o))  is a segmented sum inside the CTA, w_p = Hpp_p^-1 t_p stays in
// shared memory, and each observation then sends  -Jc_o^T (Jp_o w_p)  to y_c. H_cp = Jc^T Jp is never materialised:
// 2*(dc+3) doubles per observation are read instead of 3*dc.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
  int debug;       // development probe: 1 = suppress the reductions into y
  uint32_t ntiles;
  const DevState* st;
  const ChunkDesc* chunk_desc;
  const uint2* cslot_meta;
  const uint32_t* cpt_meta;
  const double* xpad;   // x at the padded per-camera stride (chunk kernel)
  // ranges / windows of the chunk kernel (apex_ctx.h)
  const uint16_t* cslot_widx;
  const WinDesc* win_desc;
  const uint32_t* range_win0;
  const uint32_t* win_cams;
  uint32_t window;                 // cameras per window the shared-memory rows were sized for
  uint32_t nnormal;                // normal chunks (the grid's ranges tile them)
  const uint32_t* win_dst;         // deterministic flush: row of partial[][DC] for every entry of win_cams (camera-major)
  double* partial;
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
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

// same with x stored at the padded stride xpad_stride(DC) (32-byte aligned camera blocks). The gather costs one L1
// wavefront per (lane, instruction) because every lane reads another camera's line, so the block is fetched with as
// few instructions as possible: 256-bit loads (LDG.E.ENL2.256) plus one 64/128-bit load for the remainder.
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// same through the coherent path (L1-cached): x of the PCG loop is rewritten by the tail kernel while CTAs of the next operator
// launch may already be resident (programmatic dependent launch), which the read-only contract of ld.global.nc does not allow
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.ca.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <int DC>
__device__ __forceinline__ void obs_forward_padded(const double* jc, const double* jp, const double* __restrict__ xc, double u[3]) {
  constexpr int N4 = DC / 4, REM = DC % 4;
  double xv[4 * N4 + 4];
#pragma unroll
  for (int m = 0; m < N4; ++m) ldg256(xc + 4 * m, xv[4 * m], xv[4 * m + 1], xv[4 * m + 2], xv[4 * m + 3]);
  if (REM == 1) xv[4 * N4] = __ldg(xc + 4 * N4);
  else if (REM == 2) { const double2 v = __ldg(reinterpret_cast<const double2*>(xc + 4 * N4)); xv[4 * N4] = v.x; xv[4 * N4 + 1] = v.y; }
  else if (REM == 3) ldg256(xc + 4 * N4, xv[4 * N4], xv[4 * N4 + 1], xv[4 * N4 + 2], xv[4 * N4 + 3]);
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int k = 0; k < DC; ++k) { a0 = fma(jc[k], xv[k], a0); a1 = fma(jc[DC + k], xv[k], a1); }
#pragma unroll
  for (int k = 0; k < 3; ++k) u[k] = fma(jp[k], a0, jp[3 + k] * a1);
}

// y_c -= Jc^T (Jp w)
template <int DC>
__device__ __forceinline__ void obs_backward(const double* jc, const double* jp, const double w[3], double* yc, int debug = 0) {
  const double b0 = fma(jp[0], w[0], fma(jp[1], w[1], jp[2] * w[2]));
  const double b1 = fma(jp[3], w[0], fma(jp[4], w[1], jp[5] * w[2]));
#pragma unroll
  for (int k = 0; k < DC; ++k) {
    const double v = -fma(jc[k], b0, jc[DC + k] * b1);
    if (debug != 1 || v == 1.2345e300) red_add(yc + k, v);
  }
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

// Tile kernel: the three modes for a landmark with MORE than 256 observations (its observations fill several chunks that
// hold nothing else; in these chunks the camera half of the Jacobian is stored at the observation's own lane, not in
// camera-sorted order). Accumulate per thread -> fixed-order block reduction -> re-read J for the camera side.
template <int DC, int MODE>
__global__ void __launch_bounds__(TILE, 3) schur_tile_kernel(SchurArgs a) {
  constexpr int NP = 2 * (DC + 3);
  if (a.check_done && a.st->pcg_done) return;
  __shared__ double sh[3][TILE];
  __shared__ double shw[3][TILE];
  const TileDesc td = a.tiles[blockIdx.x];
  const int tid = threadIdx.x;
  {
    // one landmark spread over several chunks: accumulate per thread, reduce in fixed order, re-read J
    double u[3] = {0.0, 0.0, 0.0};
    if (MODE != MODE_RHS) {
      for (uint32_t ch = 0; ch < td.nchunks; ++ch) {
        const size_t chunk = (size_t)td.chunk0 + ch;
        const uint32_t cam = a.slot_cam[chunk * TILE + tid];
        if (cam == PAD_CAM) continue;
        double jall[NP], uu[3];
        load_jacobian_planes<NP>(a.J, chunk, tid, jall);
        obs_forward<DC>(jall, jall + 2 * DC, a.x + (size_t)cam * DC, uu);
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
      double jall[NP];
      load_jacobian_planes<NP>(a.J, chunk, tid, jall);
      obs_backward<DC>(jall, jall + 2 * DC, w, a.y + (size_t)cam * DC);
    }
  }
}

// x (stride dc) -> x at an even stride (camera blocks 16-byte aligned for the 128-bit gathers)
__global__ void pad_x_kernel(const double* __restrict__ x, double* __restrict__ xpad, uint32_t ncam, int dc, int xs, const DevState* st, int check_done) {
  if (check_done && st->pcg_done) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncam * (uint32_t)xs) return;
  const uint32_t cam = i / xs, k = i % xs;
  xpad[i] = k < (uint32_t)dc ? x[(size_t)cam * dc + k] : 0.0;
}

// ----------------------------------------------------------------------------------------------------
// Chunk kernel: persistent 256-thread CTAs, 3 per SM; CTA r walks the chunks of range r (equal shares of the normal chunks),
// one chunk (whole landmarks, <= 256 observations) per loop iteration.
//
// SPLIT SLOT ORDER. The two halves of an observation's Jacobian live in DIFFERENT orders inside the chunk: the landmark
// half Jp (2x3) in point-major order (a landmark's observations are consecutive lanes), the camera half Jc (2xdc) in
// CAMERA-SORTED order (a camera's observations are consecutive lanes; the permutation is built at upload,
// cslot_meta). Thread t therefore holds Jp of point-major slot t and Jc of camera-sorted slot t, and the operator is
//     a_o = Jc_o x_c            camera-sorted lanes: x is gathered with neighbouring lanes reading the same / adjacent rows
//     u_o = Jp_o^T a_o          point-major lanes after a 2-value exchange through shared memory (a at its point-major slot)
//     t_p = sum_o u_o           segmented warp-shuffle sum over the landmark's run of lanes (+ shared memory across warps)
//     w_p = Hpp_p^-1 t_p        one thread per landmark
//     b_o = Jp_o w_p            point-major lanes; 2 values back to the camera-sorted slot
//     c_o = -Jc_o^T b_o         camera-sorted lanes; summed over each camera's run of lanes with segmented warp shuffles
// Only 2 + 2 values per observation cross lanes through shared memory (instead of dc), both exchanges are one STS + one
// stride-1 LDS per value. Every global read of a chunk is issued at the top of its iteration (registers / cp.async), so one
// memory latency is exposed per chunk and CTA, hidden by the two other CTAs of the SM.
//
// WINDOWS. FP64 reductions into L2 cost 1.3 cycles per lane on the SM's path to the crossbar and do not overlap with the
// loads that share it (measured: ~1 500 reductions per chunk added 1 850 cycles to the 3 000 the rest of a chunk takes), so
// the camera-side sums do not leave the SM per chunk: consecutive chunks of a locality-ordered reconstruction see the same
// few hundred cameras, the host cuts every range into windows of chunks touching <= W distinct cameras, and the CTA adds the
// run sums into its window's rows of y in shared memory (ywin, indexed by the window-local camera index cslot_widx): a run's
// first lane does a plain read-modify-write - runs of one warp name distinct cameras, and a run cut by a warp boundary
// ("continuation") parks its sum in scont and is added by its warp's lanes 0..dc-1 after the next barrier. The window is
// flushed once:
//   DET   as a contiguous block of rows partial[win.cam0 + i][dc]; det_reduce_kernel then adds each camera's rows in a fixed
//         order. Every sum has a fixed order => bitwise reproducible; costs rows*dc*16 bytes of extra traffic per application
//         (Venice-1778 shape: 2 %).
//   !DET  with one red.global.add.f64 per (camera of the window, dof) (a landmark order without camera locality makes the
//         windows short and the rows many; problem_upload picks the flush, APEX_DETERMINISTIC overrides).
// ----------------------------------------------------------------------------------------------------
template <int DC>
__device__ __forceinline__ void gather_x_padded(const double* __restrict__ xc, double* xv) {
  // one L1 wavefront per distinct 128-byte line and instruction: as few instructions as possible (256-bit loads)
  constexpr int N4 = DC / 4, REM = DC % 4;
#pragma unroll
  for (int m = 0; m < N4; ++m) ld256(xc + 4 * m, xv[4 * m], xv[4 * m + 1], xv[4 * m + 2], xv[4 * m + 3]);
  if (REM == 1) { asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(xv[4 * N4]) : "l"(xc + 4 * N4)); }
  else if (REM == 2) { asm volatile("ld.global.ca.v2.f64 {%0, %1}, [%2];" : "=d"(xv[4 * N4]), "=d"(xv[4 * N4 + 1]) : "l"(xc + 4 * N4)); }
  else if (REM == 3) ld256(xc + 4 * N4, xv[4 * N4], xv[4 * N4 + 1], xv[4 * N4 + 2], xv[4 * N4 + 3]);
}

// shared memory of the chunk kernel in doubles: [Jacobian stage] work arrays, continuation sums, mbarrier, window rows
template <int MODE>
__host__ __device__ constexpr int chunk_work_doubles() { return 2 * TILE + 3 * TILE + (3 + 6 + (MODE == MODE_MATVEC ? 0 : 3)) * MAX_TILE_PTS + MAX_TILE_PTS / 2; }
template <int DC, bool STAGED>
__host__ __device__ constexpr int chunk_stage_doubles() { return STAGED ? 2 * (DC + 3) * TILE : 0; }
__host__ __device__ constexpr int ywin_stride(int dc) { return (dc + 1) & ~1; }   // window rows padded to 16 bytes: 128-bit read-modify-write
template <int DC, int MODE, bool STAGED>
__host__ __device__ constexpr size_t chunk_smem_bytes(uint32_t W) {
  return sizeof(double) * ((size_t)chunk_stage_doubles<DC, STAGED>() + chunk_work_doubles<MODE>() + 8 * DC + 2 + (size_t)W * ywin_stride(DC));
}

// programmatic dependent launch (see the note in front of pcg_tail_kernel)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait on a phase parity: ~2 s of polling, then trap (an error the host sees) instead of a hang
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int DC, int MODE, bool DET, bool STAGED>
__global__ void __launch_bounds__(TILE, STAGED ? 2 : 3) schur_chunk_kernel(SchurArgs a) {
  constexpr int NPAIR = DC + 3;
  constexpr int XS = xpad_stride(DC);
  constexpr int DCP = ywin_stride(DC);
  constexpr int NGP = MODE == MODE_MATVEC ? 0 : 3;
  constexpr int WORK = chunk_work_doubles<MODE>();
  constexpr int STG = chunk_stage_doubles<DC, STAGED>();
  constexpr uint32_t JBYTES = 16u * NPAIR * TILE;   // one chunk's Jacobian planes, contiguous in HBM
  extern __shared__ __align__(128) double sm[];
  double* wk = sm + STG;
  double (*sab)[TILE] = reinterpret_cast<double (*)[TILE]>(wk);                         // [2] a: camera-sorted -> point-major; later b: back
  double (*su)[TILE] = reinterpret_cast<double (*)[TILE]>(wk + 2 * TILE);                // [3]
  double (*sw)[MAX_TILE_PTS] = reinterpret_cast<double (*)[MAX_TILE_PTS]>(wk + 5 * TILE);               // [3]
  double (*shinv)[MAX_TILE_PTS] = reinterpret_cast<double (*)[MAX_TILE_PTS]>(wk + 5 * TILE + 3 * MAX_TILE_PTS);  // [6]
  double (*sgp)[MAX_TILE_PTS] = reinterpret_cast<double (*)[MAX_TILE_PTS]>(wk + 5 * TILE + 9 * MAX_TILE_PTS);    // [3] (not MATVEC)
  uint32_t* sptm = reinterpret_cast<uint32_t*>(wk + 5 * TILE + (9 + NGP) * MAX_TILE_PTS);
  double (*scont)[DC] = reinterpret_cast<double (*)[DC]>(wk + WORK);                     // [8] run sums of continuation lanes, per warp
  const uint32_t bar = smem_u32(wk + WORK + 8 * DC);                                     // mbarrier of the Jacobian stage
  double* ywin = wk + WORK + 8 * DC + 2;                                                 // [window][DCP], 16-byte aligned rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t win_begin = __ldg(a.range_win0 + blockIdx.x), win_end = __ldg(a.range_win0 + blockIdx.x + 1);   // consumed at the window loop
  // the chunks of a range are contiguous: everything a chunk needs is requested ONE CHUNK AHEAD -
  //   Jacobian planes  STAGED: one cp.async.bulk (TMA) of the chunk's 16*(dc+3)*256 contiguous bytes into the stage, armed on an
  //                    mbarrier; the threads move their 12 x 16 bytes from the stage into registers, and as soon as every
  //                    thread has done so (the first barrier of the chunk) the stage is refilled with the next chunk, so the
  //                    copy runs under the whole computation of the current chunk. Registers are the second pipeline stage.
  //                    !STAGED: plain 128-bit loads at the top of the chunk's iteration (exposed once per chunk and CTA).
  //   slot metadata, window indices, chunk descriptor: registers, loaded during the previous chunk
  //   landmark inverses / gradients / run tables: cp.async into shared memory once the previous chunk's phase 2 has read its own
  // (range r = chunks [nn r / P, nn (r+1) / P) like the host built it: no dependent load in front of the first requests)
  const uint32_t chunk_first = (uint32_t)((uint64_t)a.nnormal * blockIdx.x / gridDim.x), chunk_last = (uint32_t)((uint64_t)a.nnormal * (blockIdx.x + 1) / gridDim.x);
  if (chunk_first == chunk_last) return;
  auto issue_stage = [&](uint32_t chunk) {
    mbar_expect_tx(bar, JBYTES);
    bulk_g2s(smem_u32(sm), a.J + (size_t)chunk * 2 * NPAIR * TILE, JBYTES, bar);
  };
  auto issue_landmarks = [&](const uint2& dsc) {   // {first landmark, landmarks} of the chunk
    if ((uint32_t)tid < dsc.y) {
      const uint32_t lp = dsc.x + tid;
#pragma unroll
      for (int k = 0; k < 6; ++k) cp_async8(&shinv[k][tid], a.hinv + (size_t)k * a.npl + lp);
      if (MODE != MODE_MATVEC) {
#pragma unroll
        for (int k = 0; k < 3; ++k) cp_async8(&sgp[k][tid], a.gp + (size_t)k * a.npl + lp);
      }
      cp_async4(&sptm[tid], a.cpt_meta + lp);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (STAGED) {
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) issue_stage(chunk_first);
  }
  uint2 meta_n = __ldg(a.cslot_meta + (size_t)chunk_first * TILE + tid);
  uint32_t widx_n = __ldg(a.cslot_widx + (size_t)chunk_first * TILE + tid);
  // (only the fields that are used: a 128-bit load of the descriptor leaves a dead destination register that the compiler
  // reuses at once, and the write-after-write wait on it exposed the prefetch's whole latency)
  uint2 dsc_n = __ldg(reinterpret_cast<const uint2*>(a.chunk_desc + chunk_first));
  uint32_t nobs_n = __ldg(&a.chunk_desc[chunk_first].nobs);
  issue_landmarks(dsc_n);
  // Everything above reads structure and linearisation data that no kernel of the PCG loop writes: as a programmatic dependent
  // (see pcg_tail_kernel) the CTA has done it while the tail in front was still running. x, the done flag and every write follow.
  pdl_wait();
  if (a.check_done && __ldcg(&a.st->pcg_done)) {   // (the copies in flight land in this CTA's shared memory: wait for them)
    if (STAGED) mbar_wait(bar, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }
  pdl_trigger();
  // x of the chunk's cameras, gathered one chunk ahead as well (into the registers the camera half of the previous chunk frees)
  double xv[4 * (DC / 4) + 4];
  // (only with the two-CTA register budget of the staged kernel; at 80 registers the gather stays at its use)
  auto prefetch_x = [&](uint32_t cam) {
    if (STAGED && MODE != MODE_RHS && cam != PAD_CAM) gather_x_padded<DC>(a.xpad + (size_t)cam * XS, xv);
  };
  prefetch_x(meta_n.x);
  uint32_t stage_parity = 0;
  uint32_t fix_flags = 0, fix_widx = 0;   // pending continuation sums of the previous chunk (this warp's lane 0)
  // add the parked continuation sums of the previous chunk: lanes 0..DC-1 of the warp whose lane 0 started the chain
  auto fix_up = [&]() {
    if ((fix_flags & CONT_FIRST_BIT) && lane < DC) {
      const int len = (int)((fix_flags >> CONT_LEN_SHIFT) & 7u);
      double s = scont[warp][lane];
      for (int v = 1; v < len; ++v) s += scont[warp + v][lane];
      ywin[fix_widx * DCP + lane] += s;
    }
    fix_flags = 0;
  };
  for (uint32_t win = win_begin; win < win_end; ++win) {
    const uint4 wd = __ldg(reinterpret_cast<const uint4*>(a.win_desc + win));   // chunk_begin, chunk_end, cam0, ncams
    if (MODE != MODE_BACKSUB) {
      for (uint32_t i = tid; i < wd.w * DCP; i += TILE) ywin[i] = 0.0;   // ordered before the first read-modify-write by the chunk's barriers
    }
    for (uint32_t chunk = wd.x; chunk < wd.y; ++chunk) {
      const uint2 meta = meta_n;
      const uint32_t widx = widx_n;
      const uint2 dsc = dsc_n;
      const uint32_t nobs = nobs_n;
      const bool has_next = chunk + 1 < chunk_last;
      if (has_next) {
        meta_n = __ldg(a.cslot_meta + (size_t)(chunk + 1) * TILE + tid);
        widx_n = __ldg(a.cslot_widx + (size_t)(chunk + 1) * TILE + tid);
        dsc_n = __ldg(reinterpret_cast<const uint2*>(a.chunk_desc + chunk + 1));
        nobs_n = __ldg(&a.chunk_desc[chunk + 1].nobs);
      }
      double jc[2 * DC], jp[6];
      if (STAGED) {
        if (MODE != MODE_BACKSUB) fix_up();
        mbar_wait(bar, stage_parity);
        stage_parity ^= 1u;
        // landmark half first: shared-memory loads of a warp return in order, and phase 1 consumes the camera half before the
        // barrier that lets the next copy overwrite the stage
        const double2* p = reinterpret_cast<const double2*>(sm) + tid;
#pragma unroll
        for (int m = 0; m < 3; ++m) { const double2 v = p[(DC + m) * TILE]; jp[2 * m] = v.x; jp[2 * m + 1] = v.y; }
#pragma unroll
        for (int m = 0; m < DC; ++m) { const double2 v = p[m * TILE]; jc[2 * m] = v.x; jc[2 * m + 1] = v.y; }
      } else {
        const double2* p = reinterpret_cast<const double2*>(a.J) + (size_t)chunk * NPAIR * TILE + tid;
#pragma unroll
        for (int m = 0; m < DC; ++m) { const double2 v = ld_stream2(p + (size_t)m * TILE); jc[2 * m] = v.x; jc[2 * m + 1] = v.y; }
#pragma unroll
        for (int m = 0; m < 3; ++m) { const double2 v = ld_stream2(p + (size_t)(DC + m) * TILE); jp[2 * m] = v.x; jp[2 * m + 1] = v.y; }
        if (MODE != MODE_BACKSUB) fix_up();   // while the loads are in flight
      }
      const uint32_t npt = dsc.y;
      const uint32_t camc = meta.x;                 // camera of camera-sorted slot tid (PAD_CAM behind the chunk's observations)
      const bool pt_valid = (uint32_t)tid < nobs;   // point-major slot tid holds an observation
      // ---- phase 1: a_o = Jc_o x_c on the camera-sorted lanes, sent to the observation's point-major slot ----
      if (MODE != MODE_RHS) {
        double a0 = 0.0, a1 = 0.0;
        if (camc != PAD_CAM) {
          if (!STAGED) gather_x_padded<DC>(a.xpad + (size_t)camc * XS, xv);
          double a0b = 0.0, a1b = 0.0;   // two chains per sum: four independent DFMA chains instead of two of length dc
#pragma unroll
          for (int k = 0; k + 1 < DC; k += 2) {
            a0 = fma(jc[k], xv[k], a0); a1 = fma(jc[DC + k], xv[k], a1);
            a0b = fma(jc[k + 1], xv[k + 1], a0b); a1b = fma(jc[DC + k + 1], xv[k + 1], a1b);
          }
          if (DC & 1) { a0 = fma(jc[DC - 1], xv[DC - 1], a0); a1 = fma(jc[2 * DC - 1], xv[DC - 1], a1); }
          a0 += a0b; a1 += a1b;
          const uint32_t ipos = (meta.y >> 16) & 0xFFu;
          sab[0][ipos] = a0; sab[1][ipos] = a1;
        } else if (STAGED) {
          // padding lanes consume their (zero) camera half as well: every thread has its stage reads behind it at the barrier
#pragma unroll
          for (int k = 0; k < 2 * DC; ++k) a0 += jc[k];
          if (a0 != 0.0) sab[0][0] = a0;   // never taken
        }
        __syncthreads();
        if (STAGED && tid == 0 && has_next) issue_stage(chunk + 1);
        // u_o = Jp_o^T a_o; the sum over a landmark's observations (consecutive lanes) starts as a segmented warp-shuffle
        // reduction; only the first lane of each run writes its partial to shared memory
        double u[3] = {0.0, 0.0, 0.0};
        if (pt_valid) {
          const double b0 = sab[0][tid], b1 = sab[1][tid];
#pragma unroll
          for (int k = 0; k < 3; ++k) u[k] = fma(jp[k], b0, jp[3 + k] * b1);
        }
        const uint32_t key = pt_valid ? (meta.y & 0xFFu) : 0xFFFFu;  // chunk-local landmark
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t ok = __shfl_down_sync(0xffffffffu, key, d);
          const bool hit = lane + d < 32 && ok == key;
          if (!__any_sync(0xffffffffu, hit)) break;  // runs are contiguous: no partner at distance d => none further away
          const double o0 = __shfl_down_sync(0xffffffffu, u[0], d), o1 = __shfl_down_sync(0xffffffffu, u[1], d), o2 = __shfl_down_sync(0xffffffffu, u[2], d);
          if (hit) { u[0] += o0; u[1] += o1; u[2] += o2; }
        }
        const uint32_t pk = __shfl_up_sync(0xffffffffu, key, 1);
        if ((lane == 0 || pk != key) && pt_valid) { su[0][tid] = u[0]; su[1][tid] = u[1]; su[2][tid] = u[2]; }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (MODE == MODE_BACKSUB) {
        // ---- back-substitution: one thread per landmark, dp = Hpp^-1 (-g_p - t_p) ----
        if ((uint32_t)tid < npt) {
          double t0 = 0.0, t1 = 0.0, t2 = 0.0;
          const uint32_t mm = sptm[tid], off = mm & 0xFFFFu, cnt = mm >> 16;
          if (cnt) {
            t0 = su[0][off]; t1 = su[1][off]; t2 = su[2][off];
            for (uint32_t b = (off & ~31u) + 32; b < off + cnt; b += 32) { t0 += su[0][b]; t1 += su[1][b]; t2 += su[2][b]; }  // runs continuing in the next warps
          }
          const double v0 = -sgp[0][tid] - t0, v1 = -sgp[1][tid] - t1, v2 = -sgp[2][tid] - t2;
          const double h00 = shinv[0][tid], h01 = shinv[1][tid], h02 = shinv[2][tid], h11 = shinv[3][tid], h12 = shinv[4][tid], h22 = shinv[5][tid];
          const size_t lp = dsc.x + tid;
          a.step_pt[3 * lp] = h00 * v0 + h01 * v1 + h02 * v2;
          a.step_pt[3 * lp + 1] = h01 * v0 + h11 * v1 + h12 * v2;
          a.step_pt[3 * lp + 2] = h02 * v0 + h12 * v1 + h22 * v2;
        }
        __syncthreads();
        if (has_next) { issue_landmarks(dsc_n); prefetch_x(meta_n.x); }
        continue;
      }
      // ---- phases 2 + 3 on the point-major lanes: every observation computes its landmark's w_p = Hpp_p^-1 t_p itself (the lanes
      // of a run read the same shared-memory words: broadcasts) instead of waiting for one thread per landmark behind another
      // barrier; b_o = Jp_o w_p goes to the observation's camera-sorted slot ----
      if (pt_valid) {
        const uint32_t spt = meta.y & 0xFFu, pos = (meta.y >> 8) & 0xFFu;
        double v0, v1, v2;
        if (MODE == MODE_MATVEC) {
          const uint32_t mm = sptm[spt], off = mm & 0xFFFFu, cnt = mm >> 16;
          v0 = su[0][off]; v1 = su[1][off]; v2 = su[2][off];
          for (uint32_t b = (off & ~31u) + 32; b < off + cnt; b += 32) { v0 += su[0][b]; v1 += su[1][b]; v2 += su[2][b]; }  // runs continuing in the next warps
        } else { v0 = -sgp[0][spt]; v1 = -sgp[1][spt]; v2 = -sgp[2][spt]; }
        const double h00 = shinv[0][spt], h01 = shinv[1][spt], h02 = shinv[2][spt], h11 = shinv[3][spt], h12 = shinv[4][spt], h22 = shinv[5][spt];
        const double w0 = h00 * v0 + h01 * v1 + h02 * v2, w1 = h01 * v0 + h11 * v1 + h12 * v2, w2 = h02 * v0 + h12 * v1 + h22 * v2;
        sab[0][pos] = fma(jp[0], w0, fma(jp[1], w1, jp[2] * w2));
        sab[1][pos] = fma(jp[3], w0, fma(jp[4], w1, jp[5] * w2));
      }
      __syncthreads();
      if (has_next) issue_landmarks(dsc_n);   // every lane has read this chunk's landmark data
      // ---- phase 4: c_o = -Jc_o^T b_o, summed over each camera's run of lanes; the first lane of a run owns the result ----
      double cv[DC];
      if (camc != PAD_CAM) {
        const double b0 = sab[0][tid], b1 = sab[1][tid];
#pragma unroll
        for (int k = 0; k < DC; ++k) cv[k] = -fma(jc[k], b0, jc[DC + k] * b1);
      } else {
#pragma unroll
        for (int k = 0; k < DC; ++k) cv[k] = 0.0;
      }
      if (has_next) prefetch_x(meta_n.x);   // the camera half is dead from here on
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ok = __shfl_down_sync(0xffffffffu, camc, d);
        const bool hit = lane + d < 32 && ok == camc && camc != PAD_CAM;
        if (!__any_sync(0xffffffffu, hit)) break;
#pragma unroll
        for (int k = 0; k < DC; ++k) {
          const double o = __shfl_down_sync(0xffffffffu, cv[k], d);
          if (hit) cv[k] += o;
        }
      }
      const uint32_t pc = __shfl_up_sync(0xffffffffu, camc, 1);
      if (camc != PAD_CAM && (lane == 0 || pc != camc)) {
        if (meta.y & CONT_BIT) {   // (only lane 0 can carry the flag)
#pragma unroll
          for (int k = 0; k < DC; ++k) scont[warp][k] = cv[k];
        } else {
          double2* yr = reinterpret_cast<double2*>(ywin + widx * DCP);
#pragma unroll
          for (int m = 0; m < DCP / 2; ++m) {
            double2 v = yr[m];
            v.x += cv[2 * m];
            if (2 * m + 1 < DC) v.y += cv[2 * m + 1];
            yr[m] = v;
          }
        }
      }
      fix_flags = __shfl_sync(0xffffffffu, meta.y, 0);
      fix_widx = __shfl_sync(0xffffffffu, widx, 0);
      __syncthreads();   // work arrays free for the next chunk; run sums visible to fix_up
    }
    if (MODE == MODE_BACKSUB) continue;
    // ---- flush the window ----
    fix_up();
    __syncthreads();
    if (DET) {   // the camera's row of partial results for this window (camera-major rows: the second pass reads them contiguously)
      for (uint32_t i = tid; i < wd.w * DC; i += TILE) {
        const uint32_t c = i / DC, k = i - c * DC;
        a.partial[(size_t)__ldg(a.win_dst + wd.z + c) * DC + k] = ywin[c * DCP + k];
      }
    } else if (a.debug != 1) {
      for (uint32_t i = tid; i < wd.w * DC; i += TILE) {
        const uint32_t c = i / DC, k = i - c * DC;
        red_add(a.y + (size_t)__ldg(a.win_cams + wd.z + c) * DC + k, ywin[c * DCP + k]);
      }
    }
    __syncthreads();   // before the next window zeroes the rows
  }
}

// deterministic flush, second pass: y[camera] += sum of the camera's partial rows (contiguous, in (range, window) order) folded into a
// fixed tree: one warp per camera, lane l adds rows l, l+32, ... in order, then a butterfly over the lanes.
template <int DC>
__global__ void __launch_bounds__(256) det_reduce_kernel(const double* __restrict__ partial, const uint32_t* __restrict__ cam_row_start,
                                                         double* __restrict__ y, const DevState* st, uint32_t ncam, int check_done) {
  if (check_done && st->pcg_done) return;
  const uint32_t cam = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (cam >= ncam) return;
  double acc[DC];
#pragma unroll
  for (int k = 0; k < DC; ++k) acc[k] = 0.0;
  const uint32_t e1 = __ldg(cam_row_start + cam + 1);
  for (uint32_t e = __ldg(cam_row_start + cam) + lane; e < e1; e += 32) {
    const double* row = partial + (size_t)e * DC;
#pragma unroll
    for (int k = 0; k < DC; ++k) acc[k] += __ldcg(row + k);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
    for (int k = 0; k < DC; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
  }
  double mine = 0.0;
#pragma unroll
  for (int k = 0; k < DC; ++k) if (lane == k) mine = acc[k];
  if (lane < DC) y[(size_t)cam * DC + lane] += mine;
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
                                                        double* z, double* p, DevState* st, uint32_t n, int dc, int K, int max_it, double cg_tol,
                                                        unsigned long long* tail_slots, uint32_t ntail_slots) {
  __shared__ double sh[1024];
  for (uint32_t i = threadIdx.x; i < ntail_slots; i += 1024) tail_slots[i] = 0;   // exchange slots of the fused tail: tags restart at 1
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
    st->pcg_alpha = 0.0; st->pcg_beta = 0.0; st->ticket_a = 0; st->ticket_b = 0;
    st->ar_timeout = 0;
    st->tail_tag = 0;
  }
}

// ---- multi-CTA PCG iteration (solve_pcg_block, implicit_schur.rs:604-676) --------------------------------
// Three short kernels per iteration instead of one 1024-thread CTA: every grid-wide dot product is a per-CTA
// partial + "last CTA done" pass that sums the partials in index order (deterministic), and alpha / beta / the
// break tests stay in DevState.
constexpr int PCG_THREADS = 256;
constexpr int PCG_CAMS = 16;  // cameras per CTA of the update kernel

// p = z + beta p (beta = 0 right after pcg_init) and the padded copy the operator kernel gathers from
__global__ void __launch_bounds__(PCG_THREADS) pcg_dir_kernel(const double* __restrict__ z, double* __restrict__ p, double* __restrict__ xpad,
                                                              const DevState* st, uint32_t ncam, int dc, int xs) {
  if (st->pcg_done) return;
  const uint32_t i = blockIdx.x * PCG_THREADS + threadIdx.x;
  if (i >= ncam * (uint32_t)xs) return;
  const uint32_t cam = i / xs, k = i % xs;
  double v = 0.0;
  if (k < (uint32_t)dc) {
    const size_t row = (size_t)cam * dc + k;
    const double beta = st->pcg_beta;
    v = beta == 0.0 ? z[row] : z[row] + beta * p[row];
    p[row] = v;
  }
  xpad[i] = v;
}

// pAp = p . (S p); last CTA: break test |pAp| < 1e-20, alpha = rz_old / pAp
__global__ void __launch_bounds__(PCG_THREADS) pcg_pap_kernel(const double* __restrict__ p, const double* __restrict__ ap, double* part,
                                                              DevState* st, uint32_t n) {
  __shared__ double sh[PCG_THREADS];
  __shared__ bool is_last;
  if (st->pcg_done) return;
  const uint32_t i = blockIdx.x * PCG_THREADS + threadIdx.x;
  double v = i < n ? p[i] * ap[i] : 0.0;
  v = block_reduce_sum(v, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = v;
    __threadfence();
    is_last = atomicAdd(&st->ticket_a, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s = 0.0;
  for (uint32_t b = threadIdx.x; b < gridDim.x; b += PCG_THREADS) s += __ldcg(part + b);
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) {
    st->ticket_a = 0;
    if (fabs(s) < 1e-20) { st->pcg_iters = st->pcg_iters + 1; st->pcg_done = 1; }
    else st->pcg_alpha = st->rz_old / s;
  }
}

// p = z + beta p, its padded copy, and the start value of the operator result y0 = (H_cc + lambda I) p (add_hcc) or 0,
// one CTA per PCG_CAMS cameras (pcg_dir_kernel + hcc_apply_kernel in one launch)
__global__ void __launch_bounds__(PCG_THREADS) pcg_dir_hcc_kernel(const double* __restrict__ z, double* __restrict__ p, double* __restrict__ xpad,
                                                                  const double* __restrict__ hcc, double* __restrict__ y0, const DevState* st,
                                                                  uint32_t ncam, int dc, int xs, int add_hcc) {
  __shared__ double ps[PCG_CAMS * MAX_DC];
  if (st->pcg_done) return;
  const uint32_t cam0 = blockIdx.x * PCG_CAMS;
  const uint32_t nrow = min((uint32_t)PCG_CAMS, ncam - cam0) * dc;
  const size_t row0 = (size_t)cam0 * dc;
  const double beta = st->pcg_beta;
  if (threadIdx.x < nrow) {
    const size_t row = row0 + threadIdx.x;
    const double v = beta == 0.0 ? z[row] : z[row] + beta * p[row];
    p[row] = v;
    ps[threadIdx.x] = v;
    const uint32_t lc = threadIdx.x / dc, k = threadIdx.x % dc;
    xpad[(size_t)(cam0 + lc) * xs + k] = v;
  }
  __syncthreads();
  if (threadIdx.x < nrow) {
    double s = 0.0;
    if (add_hcc) {
      const uint32_t lc = threadIdx.x / dc, a = threadIdx.x % dc;
      const double* H = hcc + ((size_t)(cam0 + lc) * dc + a) * dc;
      const double* pc = ps + lc * dc;
      s = st->damping * pc[a];
      for (int b = 0; b < dc; ++b) s += H[b] * pc[b];
    }
    y0[row0 + threadIdx.x] = s;
  }
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// sum over the ranks of element `off` of their peer-mapped buffers, in rank order (same bits on every rank). The loads of up to
// eight peers are issued back to back before the first sum: a runtime-bounded "s += load" loop waits for one NVLink round trip
// (~1.5-2 us) per rank, which was the larger part of the 35 us per PCG iteration the 8-rank solve spent outside its kernels' work.
__device__ __forceinline__ double peer_sum(double* const* __restrict__ peer_buf, int nranks, size_t off) {
  double s = 0.0;
  for (int r0 = 0; r0 < nranks; r0 += 8) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = r0 + j < nranks ? ld_relaxed_sys(peer_buf[r0 + j] + off) : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];   // (+ 0.0 past the last rank: exact)
  }
  return s;
}

// Fused all-reduce + dot product over NVLink peer memory (comm.cu): publish "my partial operator result of this
// iteration is complete" to every peer, wait for theirs, y = sum over ranks of their partial vectors (rank order:
// bitwise identical everywhere), pAp = p . y with the last-CTA pass of pcg_pap_kernel. The spin is bounded: a peer that
// never arrives raises st->ar_timeout instead of hanging the GPU.
__global__ void __launch_bounds__(PCG_THREADS) ar_reduce_pap_kernel(double* const* __restrict__ peer_buf, unsigned long long* const* __restrict__ peer_flags,
                                                                    const unsigned long long* flags, int par, int nranks, int rank,
                                                                    const double* __restrict__ p, double* __restrict__ y, double* part, DevState* st,
                                                                    uint32_t n) {
  __shared__ double sh[PCG_THREADS];
  __shared__ bool is_last;
  if (st->pcg_done) return;
  const unsigned long long seq = st->ar_seq + 1;  // every CTA reads it before the last CTA (by ticket) advances it
  if (blockIdx.x == 0 && (int)threadIdx.x < nranks) {
    __threadfence_system();
    st_release_sys(peer_flags[threadIdx.x] + rank, seq);
  }
  if ((int)threadIdx.x < nranks) {
    long long spins = 0;
    while (ld_acquire_sys(flags + threadIdx.x) < seq) {
      if (++spins > (1ll << 26)) { atomicExch(&st->ar_timeout, 1); atomicExch(&st->pcg_done, 1); break; }  // the rest of the batch becomes no-ops
    }
  }
  __syncthreads();
  const uint32_t i = blockIdx.x * PCG_THREADS + threadIdx.x;
  double v = 0.0;
  if (i < n) {
    double s = 0.0;
    s = peer_sum(peer_buf, nranks, (size_t)par * n + i);
    y[i] = s;
    v = p[i] * s;
  }
  v = block_reduce_sum(v, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = v;
    __threadfence();
    is_last = atomicAdd(&st->ticket_a, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s = 0.0;
  for (uint32_t b = threadIdx.x; b < gridDim.x; b += PCG_THREADS) s += __ldcg(part + b);
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) {
    st->ticket_a = 0;
    st->ar_seq = seq;
    if (fabs(s) < 1e-20) { st->pcg_iters = st->pcg_iters + 1; st->pcg_done = 1; }
    else st->pcg_alpha = st->rz_old / s;
  }
}

// x += alpha p ; r -= alpha Ap ; z = M^-1 r ; last CTA: ||r|| < tol, |rz_old| < 1e-30, beta, iteration count
__global__ void __launch_bounds__(PCG_THREADS) pcg_update_kernel(const double* __restrict__ ap, const double* __restrict__ pinv,
                                                                 const double* __restrict__ p, double* x, double* r, double* z, double* part,
                                                                 DevState* st, uint32_t ncam, int dc, int K) {
  __shared__ double sh[PCG_THREADS];
  __shared__ double rs[PCG_CAMS * MAX_DC];
  __shared__ bool is_last;
  if (st->pcg_done) return;
  const double alpha = st->pcg_alpha;
  const uint32_t cam0 = blockIdx.x * PCG_CAMS;
  const uint32_t nrow = min((uint32_t)PCG_CAMS, ncam - cam0) * dc;
  const size_t row0 = (size_t)cam0 * dc;
  double rr = 0.0, rz = 0.0;
  if (threadIdx.x < nrow) {
    const size_t row = row0 + threadIdx.x;
    x[row] += alpha * p[row];
    const double rv = r[row] - alpha * ap[row];
    r[row] = rv;
    rs[threadIdx.x] = rv;
    rr = rv * rv;
  }
  __syncthreads();
  if (threadIdx.x < nrow) {
    const uint32_t lc = threadIdx.x / dc, a = threadIdx.x % dc;
    const double* P = pinv + (size_t)(cam0 + lc) * (36 + K * K);
    const double* rc = rs + lc * dc;
    double s = 0.0;
    if (a < 6) { for (int b = 0; b < 6; ++b) s += P[a * 6 + b] * rc[b]; }
    else { const double* Q = P + 36 + (a - 6) * K; for (int b = 0; b < K; ++b) s += Q[b] * rc[6 + b]; }
    z[row0 + threadIdx.x] = s;
    rz = rs[threadIdx.x] * s;
  }
  rr = block_reduce_sum(rr, sh);
  rz = block_reduce_sum(rz, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = rr;
    part[gridDim.x + blockIdx.x] = rz;
    __threadfence();
    is_last = atomicAdd(&st->ticket_b, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s0 = 0.0, s1 = 0.0;
  for (uint32_t b = threadIdx.x; b < gridDim.x; b += PCG_THREADS) { s0 += __ldcg(part + b); s1 += __ldcg(part + gridDim.x + b); }
  s0 = block_reduce_sum(s0, sh);
  s1 = block_reduce_sum(s1, sh);
  if (threadIdx.x == 0) {
    st->ticket_b = 0;
    const int iters = st->pcg_iters + 1;
    const double r_norm = sqrt(s0), rz_old = st->rz_old;
    st->pcg_iters = iters;
    st->r_norm = r_norm;
    if (r_norm < st->pcg_tol || fabs(rz_old) < 1e-30) st->pcg_done = 1;
    else {
      st->pcg_beta = s1 / rz_old;
      st->rz_old = s1;
      if (iters >= st->pcg_max) st->pcg_done = 1;
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// Fused PCG tail: everything between two operator applications in ONE launch.
//   second pass of the deterministic flush (y += the camera's partial rows)  ->  [all-reduce of the partial operator results
//   over NVLink peer memory]  ->  pAp, alpha  ->  x += alpha p, r -= alpha Ap, z = M^-1 r  ->  ||r||, r.z, beta, break tests
//   ->  p = z + beta p, its padded copy, y0 = (H_cc + lambda I) p
// The camera vectors are small (ncam*dc doubles: 128 KB on the Venice shape) and the step is pure latency: as five kernels
// (det_reduce, pcg_pap / ar_reduce_pap, pcg_update, pcg_dir_hcc) it cost ~25-35 us per PCG iteration next to an operator of
// 260 us on one GPU and 40 us on an eighth of the problem. Here one warp owns a camera (lane q = row q of its blocks), the CTAs
// of one launch are co-resident (512 threads, one CTA per SM, no dynamic shared memory) and exchange their partial sums through
// tagged slots (tail_publish / tail_gather below: barrier and gather in one step) - every grid-wide dot product is a per-CTA
// partial summed by every CTA in the same fixed order (same bits on every CTA and, with the rank-ordered peer sum, on every
// rank). Two launches per PCG iteration (operator + tail), each a programmatic dependent of the other.
// Semantics of solve_pcg_block (implicit_schur.rs:604-676) as in pcg_pap / pcg_update / pcg_dir_hcc.
// ----------------------------------------------------------------------------------------------------
constexpr int TAIL_THREADS = 512;           // 16 warps = 16 cameras per CTA, one CTA per SM: the G x G polling traffic of the exchanges is a quarter of what 256-thread CTAs cost
constexpr int TAIL_WARPS = TAIL_THREADS / 32;
constexpr unsigned TAIL_MAX_CTAS = 592;      // 4 per SM x 148: capacity of the exchange slots (in red_scratch)

struct TailArgs {
  double* const* peer_buf;               // null: single rank
  unsigned long long* const* peer_flags;
  const unsigned long long* flags;
  int par, nranks, rank;
  double* ylocal;                        // this rank's operator result: y0 + what the operator kernels reduced into it
  const double* partial;                 // deterministic flush: camera-major partial rows ...
  const uint32_t* cam_row_start;         // ... of camera c: [cam_row_start[c], cam_row_start[c+1]); null = nothing to add
  double* y;                             // the complete operator result (kept for the update stage)
  double* y0_next;
  double* p; double* x; double* r; double* z; double* xpad;
  const double* pinv; const double* hcc;
  unsigned long long* slots;             // [2][TAIL_MAX_CTAS][4] exchange slots of the grid-wide sums, then [TAIL_MAX_CTAS] "rows written" words (zeroed by pcg_init_kernel)
  DevState* st;
  uint32_t ncam;
  int K, xs, add_hcc;
  long long* trace;                      // development probe (APEX_TAIL_TRACE): %globaltimer stamps of CTA 0 per iteration, [4096][8]
};

// ---- grid-wide sums of the tail: barrier and all-gather in one step ("flag in data") ----
// A counter barrier followed by a read of the per-CTA partials costs three dependent L2 round trips (arrive, see the count,
// fetch the partials) plus G atomics on one address. Here a CTA publishes each partial as two 64-bit words {half of the double,
// 32-bit tag} (a 64-bit store is single-copy atomic, so a word with the right tag carries the right half), and one or two warps of
// every CTA poll the G slots directly - four per lane in flight - until all carry the tag: one round trip after the last CTA's
// store, and every CTA adds the partials in the same fixed order (same bits everywhere). Tags count the exchanges of the solve
// (DevState::tail_tag, zeroed with the slots by pcg_init_kernel); exchange t uses slot set t & 1: a CTA can only be one exchange
// ahead of the slowest one, because it cannot leave exchange t+1 before every CTA has published there, i.e. has finished reading t.
__device__ __forceinline__ void tail_publish(unsigned long long* set, unsigned tag, double v0, double v1, bool two) {
  unsigned long long* q = set + (size_t)blockIdx.x * 4;
  const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0);
  const unsigned long long w0 = (b0 & 0xFFFFFFFF00000000ull) | tag, w1 = (b0 << 32) | tag;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(q), "l"(w0), "l"(w1) : "memory");
  if (two) {
    const unsigned long long b1 = (unsigned long long)__double_as_longlong(v1);
    const unsigned long long w2 = (b1 & 0xFFFFFFFF00000000ull) | tag, w3 = (b1 << 32) | tag;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(q + 2), "l"(w2), "l"(w3) : "memory");
  }
}
__device__ __forceinline__ void ld_relaxed_gpu_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
template <bool TWO>
__device__ __forceinline__ void tail_gather(const unsigned long long* set, unsigned G, unsigned tag, double* shg /* shared [2 * TAIL_WARPS], this exchange's own */, double& t0, double& t1) {
  // Few pollers: the slots are hot lines in L2 and every CTA reads all of them, so the polling traffic itself delays the
  // exchange (measured: all 256 threads polling one slot each was 0.5-1 us slower per exchange than one warp polling them all).
  // Warp w < ceil(G / 128) polls slots [128 w, 128 w + 128): four per lane, all in flight at once (G <= 256: one round).
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned npoll = (G + 127u) / 128u;
  double s = 0.0, q = 0.0;
  for (unsigned base = warp * 128u; base < G; base += (TAIL_THREADS / 32) * 128u) {
    unsigned long long w[4][TWO ? 4 : 2];
    long long spins = 0;
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned b = base + 32 * j + lane;
        if (b < G) {
          ld_relaxed_gpu_v2(set + (size_t)b * 4, w[j][0], w[j][1]);
          if (TWO) ld_relaxed_gpu_v2(set + (size_t)b * 4 + 2, w[j][TWO ? 2 : 0], w[j][TWO ? 3 : 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (base + 32 * j + lane < G) {
#pragma unroll
          for (int m = 0; m < (TWO ? 4 : 2); ++m) ok = ok && (unsigned)w[j][m] == tag;
        }
      }
      if (!ok && ++spins > (1ll << 26)) __trap();   // a CTA that never publishes (not co-resident): an error the host sees, not a hang
    } while (!ok);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (base + 32 * j + lane < G) {
        s += __longlong_as_double((long long)((w[j][0] & 0xFFFFFFFF00000000ull) | (w[j][1] >> 32)));
        if (TWO) q += __longlong_as_double((long long)((w[j][TWO ? 2 : 0] & 0xFFFFFFFF00000000ull) | (w[j][TWO ? 3 : 1] >> 32)));
      }
    }
  }
  if (warp < npoll && warp < TAIL_THREADS / 32) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); if (TWO) q += __shfl_xor_sync(0xffffffffu, q, d); }
    if (lane == 0) { shg[warp] = s; shg[TAIL_WARPS + warp] = q; }
  }
  __syncthreads();
  t0 = shg[0]; t1 = shg[TAIL_WARPS];
  for (unsigned w2 = 1; w2 < npoll && w2 < TAIL_WARPS; ++w2) { t0 += shg[w2]; t1 += shg[TAIL_WARPS + w2]; }
}
// CTA sums of two values in a fixed order: butterfly inside the warps, then the warp sums in warp order (thread 0 holds them)
__device__ __forceinline__ void tail_block_sums(double& a, double& b, double* sh /* [2 * TAIL_WARPS] */) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[warp] = a; sh[TAIL_WARPS + warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = sh[0]; b = sh[TAIL_WARPS];
#pragma unroll
    for (int w = 1; w < TAIL_WARPS; ++w) { a += sh[w]; b += sh[TAIL_WARPS + w]; }
  }
}

// second pass of the deterministic flush for one camera: sum of its partial rows, every lane gets all DC sums
template <int DC, int NF>
__device__ __forceinline__ void tail_row_sums(const TailArgs& a, uint32_t e0, uint32_t e1, int lane, double acc[DC]) {   // rows [e0, e1)
#pragma unroll
  for (int k = 0; k < DC; ++k) acc[k] = 0.0;
  for (uint32_t e = e0 + lane; e < e1; e += 32 * NF) {   // NF rows per lane in flight; added in row order
    double v[NF][DC];
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const bool on = e + 32 * j < e1;
      const double* row = a.partial + (size_t)(on ? e + 32 * j : e) * DC;
#pragma unroll
      for (int k = 0; k < DC; ++k) v[j][k] = on ? __ldcg(row + k) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < NF; ++j) {
#pragma unroll
      for (int k = 0; k < DC; ++k) acc[k] += v[j][k];
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
    for (int k = 0; k < DC; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
  }
}

__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TAIL_STAMP(k) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[(size_t)(iters0 & 4095) * 8 + (k)] = global_ns(); } while (0)

// PROGRAMMATIC DEPENDENT LAUNCH. Inside the PCG loop the operator and the tail are launched as programmatic dependents of each
// other (cudaLaunchAttributeProgrammaticStreamSerialization, also inside the captured graph): a kernel's CTAs are scheduled as
// soon as every CTA of the kernel in front of it has executed griddepcontrol.launch_dependents (or exited) and SM resources
// are free, run their prologue - everything that does not depend on the kernel in front - and block in griddepcontrol.wait
// until that grid has completed and its writes are visible. The launch latency, the CTA ramp and the first memory round trips
// (operator: descriptors, the first chunk's Jacobian planes by TMA, the landmark inverses; tail: its rows of p, r, x and the
// preconditioner / H_cc blocks) overlap with the kernel in front. Both kernels trigger only AFTER their own wait, so a kernel is
// never scheduled before the kernel two in front of it has completed: what the tail reads before its wait (written by the
// previous tail) is final. Everything a kernel in front may still be writing is read after the wait, and never through the
// non-coherent path. Launched without the attribute (every use outside the loop) both instructions do nothing.

// ONE: every warp owns at most one camera (gridDim.x * 8 >= ncam): its rows of p, r, x, the preconditioner and H_cc blocks are
// loaded once at entry and stay in registers across the exchanges, so no stage waits for memory behind one.
template <int DC, bool ONE>
__global__ void __launch_bounds__(TAIL_THREADS, 1) pcg_tail_kernel(TailArgs a) {
  __shared__ double sh[2 * TAIL_WARPS];
  __shared__ double tot_b[2 * TAIL_WARPS], tot_c[2 * TAIL_WARPS];
  DevState* st = a.st;
  // (L2 loads: with an early launch this SM's L1 may still hold the lines the previous tail read before it rewrote them)
  if (__ldcg(&st->pcg_done)) return;  // same value in every CTA: it is only written once every CTA of the launch has left the last exchange
  const int tid = threadIdx.x, lane = tid & 31, K = a.K;
  const uint32_t n = a.ncam * DC;
  const uint32_t gw = blockIdx.x * (TAIL_THREADS / 32) + (tid >> 5), nw = gridDim.x * (TAIL_THREADS / 32);
  const unsigned G = gridDim.x;
  const double rz_old = __ldcg(&st->rz_old), tol = __ldcg(&st->pcg_tol), damping = __ldcg(&st->damping);
  const int iters0 = __ldcg(&st->pcg_iters), max_it = __ldcg(&st->pcg_max);
  const unsigned tag0 = __ldcg(&st->tail_tag);
  const unsigned long long seq = __ldcg(&st->ar_seq) + 1;
  const bool multi = a.peer_buf != nullptr;
  double* ymine = multi ? a.peer_buf[a.rank] + (size_t)a.par * n : a.ylocal;
  // ONE: this lane's row of everything
  const bool act = ONE && gw < a.ncam && lane < DC;
  const size_t myrow = (size_t)gw * DC + lane;
  double pq = 0.0, rq = 0.0, xq = 0.0, yq = 0.0, Pq[DC > 6 ? (DC - 6 > 6 ? DC - 6 : 6) : 6], Hq[DC];
  uint32_t row_e0 = 0, row_e1 = 0;   // this camera's partial rows of the deterministic flush (structure: read before the wait)
  if (ONE) {
    if (act) {
      pq = __ldcg(a.p + myrow); rq = __ldcg(a.r + myrow); xq = __ldcg(a.x + myrow);
      const double* P = a.pinv + (size_t)gw * (36 + K * K);
      if (lane < 6) {
#pragma unroll
        for (int b = 0; b < 6; ++b) Pq[b] = __ldg(P + lane * 6 + b);
      } else {
#pragma unroll
        for (int b = 0; b < DC - 6; ++b) Pq[b] = __ldg(P + 36 + (lane - 6) * K + b);
      }
      if (a.add_hcc) {
#pragma unroll
        for (int b = 0; b < DC; ++b) Hq[b] = __ldg(a.hcc + myrow * DC + b);
      }
    }
    if (a.cam_row_start && gw < a.ncam) { row_e0 = __ldg(a.cam_row_start + gw); row_e1 = __ldg(a.cam_row_start + gw + 1); }
  }
  pdl_wait();      // the operator in front has completed: partial rows / reductions into y0 are visible
  pdl_trigger();   // every CTA of this launch is resident: the next operator's CTAs may take free SM resources and run their prologue
  if (ONE && act) yq = __ldcg(ymine + myrow);
  TAIL_STAMP(0);
  // ---- stage A: this rank's operator result (second pass of the deterministic flush) ----
  if (a.cam_row_start) {
    if (ONE) {
      if (gw < a.ncam) {
        double acc[DC];
        tail_row_sums<DC, 2>(a, row_e0, row_e1, lane, acc);
#pragma unroll
        for (int k = 0; k < DC; ++k) if (lane == k) yq += acc[k];
        if (multi && act) ymine[myrow] = yq;
      }
    } else {
      for (uint32_t cam = gw; cam < a.ncam; cam += nw) {
        double acc[DC];
        tail_row_sums<DC, 4>(a, __ldg(a.cam_row_start + cam), __ldg(a.cam_row_start + cam + 1), lane, acc);
        double mine = 0.0;
#pragma unroll
        for (int k = 0; k < DC; ++k) if (lane == k) mine = acc[k];
        if (lane < DC) ymine[(size_t)cam * DC + lane] += mine;
      }
    }
  }
  if (multi) {
    // ---- the operator result of all ranks: publish "my partial result is complete", wait for the peers, sum in rank order.
    TAIL_STAMP(1);
    // No grid barrier: a CTA reads this rank's rows only for its own cameras (which it wrote itself); only the publisher has to
    // know that every CTA's rows are complete. Every CTA leaves "my rows of exchange `seq` are written" in its word of `adone`
    // (fence + relaxed store), CTA 0 alone polls the G words - one poller per word, so no polling traffic on hot lines - and then
    // publishes the rank's flag to the peers behind a system fence.
    // (Measured alternative, N=2: one flag per CTA at every peer - every rank runs the same camera -> warp map, so CTA b only
    // needs the rows the peers' CTAs b wrote: 223 system fences + NVLink flag writes per rank and iteration instead of one, 12 us
    // in the flag wait and CTAs leaving it up to 8 us apart; 38.3 -> 35.9 LM it/s.)
    unsigned long long* adone = a.slots + 2 * (size_t)TAIL_MAX_CTAS * 4;
    __syncthreads();
    if (tid == 0) { __threadfence(); asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(adone + blockIdx.x), "l"(seq) : "memory"); }
    if (blockIdx.x == 0) {
      for (unsigned b = tid; b < G; b += TAIL_THREADS) {
        long long spins = 0;
        unsigned long long v;
        do {
          asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(adone + b) : "memory");
          if (++spins > (1ll << 26)) __trap();   // a CTA that never arrives (not co-resident): an error the host sees, not a hang
        } while (v < seq);
      }
      __threadfence();
      __syncthreads();
      TAIL_STAMP(2);
      if (tid < a.nranks) {
        __threadfence_system();
        st_release_sys(a.peer_flags[tid] + a.rank, seq);
      }
    }
    if (tid < a.nranks) {
      long long spins = 0;
      while (ld_acquire_sys(a.flags + tid) < seq) {
        if (++spins > (1ll << 26)) { atomicExch(&st->ar_timeout, 1); __threadfence(); break; }  // the sums below are garbage; solve_implicit sees the flag
      }
    }
    __syncthreads();
    TAIL_STAMP(3);
  }
  // ---- stage B: y (complete), pAp ----
  double v = 0.0, vdummy = 0.0;
  if (ONE) {
    if (act) {
      if (multi) yq = peer_sum(a.peer_buf, a.nranks, (size_t)a.par * n + myrow);
      v = pq * yq;
    }
  } else {
    for (uint32_t cam = gw; cam < a.ncam; cam += nw) {
      if (lane < DC) {
        const size_t row = (size_t)cam * DC + lane;
        double s;
        if (multi) s = peer_sum(a.peer_buf, a.nranks, (size_t)a.par * n + row);
        else s = __ldcg(ymine + row);
        a.y[row] = s;
        v += __ldcg(a.p + row) * s;
      }
    }
  }
  tail_block_sums(v, vdummy, sh);
  unsigned long long* set_b = a.slots + (size_t)((tag0 + 1u) & 1u) * TAIL_MAX_CTAS * 4;
  if (tid == 0) tail_publish(set_b, tag0 + 1u, v, 0.0, false);
  TAIL_STAMP(4);
  double pap, unused;
  tail_gather<false>(set_b, G, tag0 + 1u, tot_b, pap, unused);
  TAIL_STAMP(5);
  bool timed_out = false;
  if (multi) { __threadfence(); timed_out = *reinterpret_cast<volatile int32_t*>(&st->ar_timeout) != 0; }   // (set and fenced before the CTA that saw it published)
  if (fabs(pap) < 1e-20 || timed_out) {  // break before the update (implicit_schur.rs:626-629); every CTA takes the same branch
    // (every CTA has published above, i.e. is past its entry reads of the scalars: they may be rewritten now)
    if (blockIdx.x == 0 && tid == 0) { st->pcg_iters = iters0 + 1; st->pcg_done = 1; st->tail_tag = tag0 + 1u; if (multi) st->ar_seq = seq; }
    return;
  }
  const double alpha = rz_old / pap;
  // ---- stage C: x, r, z = M^-1 r, ||r||^2, r.z ----
  // z_q = sum_b P[q][b] r_b on the 6x6 pose block and the KxK intrinsics block (apply_preconditioner, implicit_schur.rs:409-443)
  double rr = 0.0, rz = 0.0, zq = 0.0;
  if (ONE) {
    double rv = 0.0;
    if (act) {
      a.x[myrow] = xq + alpha * pq;
      rv = rq - alpha * yq;
      a.r[myrow] = rv;
      rr = rv * rv;
    }
#pragma unroll
    for (int b = 0; b < DC; ++b) {
      const double rb = __shfl_sync(0xffffffffu, rv, b);
      if (lane < 6) { if (b < 6) zq += Pq[b] * rb; }
      else if (lane < DC) { if (b >= 6) zq += Pq[b - 6 < 0 ? 0 : b - 6] * rb; }
    }
    if (act) { a.z[myrow] = zq; rz = rv * zq; }
  } else {
    for (uint32_t cam = gw; cam < a.ncam; cam += nw) {
      const size_t row = (size_t)cam * DC + lane;
      double rv = 0.0;
      if (lane < DC) {
        a.x[row] = __ldcg(a.x + row) + alpha * __ldcg(a.p + row);
        rv = __ldcg(a.r + row) - alpha * a.y[row];
        a.r[row] = rv;
        rr += rv * rv;
      }
      const double* P = a.pinv + (size_t)cam * (36 + K * K);
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < DC; ++b) {
        const double rb = __shfl_sync(0xffffffffu, rv, b);
        if (lane < 6) { if (b < 6) s += __ldg(P + lane * 6 + b) * rb; }
        else if (lane < DC) { if (b >= 6) s += __ldg(P + 36 + (lane - 6) * K + (b - 6)) * rb; }
      }
      if (lane < DC) { a.z[row] = s; rz += rv * s; }
    }
  }
  tail_block_sums(rr, rz, sh);   // (thread 0 read the stage-B warp sums before it published: in front of the gather's barrier)
  unsigned long long* set_c = a.slots + (size_t)((tag0 + 2u) & 1u) * TAIL_MAX_CTAS * 4;
  if (tid == 0) tail_publish(set_c, tag0 + 2u, rr, rz, true);
  double rr_tot, rz_tot;
  tail_gather<true>(set_c, G, tag0 + 2u, tot_c, rr_tot, rz_tot);
  TAIL_STAMP(6);
  const int iters = iters0 + 1;
  const double r_norm = sqrt(rr_tot);
  bool done = r_norm < tol || fabs(rz_old) < 1e-30;
  double beta = 0.0;
  const bool have_beta = !done;
  if (have_beta) { beta = rz_tot / rz_old; if (iters >= max_it) done = true; }
  // ---- stage D: next direction and the start value of the next operator result ----
  if (!done) {
    if (ONE) {
      double pv = 0.0;
      if (act) {
        pv = beta == 0.0 ? zq : zq + beta * pq;
        a.p[myrow] = pv;
        a.xpad[(size_t)gw * a.xs + lane] = pv;
      }
      double s = damping * pv;
#pragma unroll
      for (int b = 0; b < DC; ++b) {
        const double pb = __shfl_sync(0xffffffffu, pv, b);
        if (a.add_hcc && act) s += Hq[b] * pb;
      }
      if (act) a.y0_next[myrow] = a.add_hcc ? s : 0.0;
    } else {
      for (uint32_t cam = gw; cam < a.ncam; cam += nw) {
        const size_t row = (size_t)cam * DC + lane;
        double pv = 0.0;
        if (lane < DC) {
          const double zv = a.z[row];
          pv = beta == 0.0 ? zv : zv + beta * __ldcg(a.p + row);
          a.p[row] = pv;
          a.xpad[(size_t)cam * a.xs + lane] = pv;
        }
        double s = damping * pv;
        const double* H = a.hcc + ((size_t)cam * DC + (lane < DC ? lane : 0)) * DC;
#pragma unroll
        for (int b = 0; b < DC; ++b) {
          const double pb = __shfl_sync(0xffffffffu, pv, b);
          if (a.add_hcc && lane < DC) s += __ldg(H + b) * pb;
        }
        if (lane < DC) a.y0_next[row] = a.add_hcc ? s : 0.0;
      }
    }
  }
  TAIL_STAMP(7);
  // (every CTA passed its entry reads of these scalars before it published into the last exchange)
  if (blockIdx.x == 0 && tid == 0) {
    st->pcg_iters = iters;
    st->r_norm = r_norm;
    st->pcg_alpha = alpha;
    st->tail_tag = tag0 + 2u;
    if (have_beta) { st->pcg_beta = beta; st->rz_old = rz_tot; }
    if (done) st->pcg_done = 1;
    if (multi) st->ar_seq = seq;
  }
}

// launch with one argument struct, optionally as a programmatic dependent of the kernel in front of it on the stream
template <typename A>
static cudaError_t launch_dep(void (*kern)(A), unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool pdl, const A& arg) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, arg);
}

// CTAs of the fused tail (0 = the separate kernels): one warp per camera when the device can keep that many CTAs resident at
// once (occupancy of the instantiation x SMs, at most APEX_PCG_TAIL = 2 per SM), else the grid-stride variant on as many CTAs
// as are co-resident. The grid barriers need every CTA of the launch on an SM at the same time.
template <int DC>
static int tail_plan_dc(Ctx& c, bool& one) {
  const char* e = getenv("APEX_PCG_TAIL");
  const int want = e ? atoi(e) : 2;
  if (want <= 0) return 0;
  const int per_sm_want = std::min(std::min(want, 4), (int)(TAIL_MAX_CTAS / (unsigned)std::max(c.num_sms, 1)));
  if (per_sm_want < 1) return 0;
  const uint32_t need = (c.ncam + TAIL_THREADS / 32 - 1) / (TAIL_THREADS / 32);
  int occ = 0;
  const char* gs = getenv("APEX_PCG_TAIL_STRIDE");   // test switch: the grid-stride variant also where one warp per camera would fit
  const bool force_stride = gs && atoi(gs) != 0;
  if (!force_stride && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcg_tail_kernel<DC, true>, TAIL_THREADS, 0) == cudaSuccess && occ > 0 &&
      need <= (uint32_t)(std::min(occ, per_sm_want) * c.num_sms)) { one = true; return (int)std::max<uint32_t>(need, 1); }
  cudaGetLastError();
  one = false;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcg_tail_kernel<DC, false>, TAIL_THREADS, 0) != cudaSuccess || occ <= 0) { cudaGetLastError(); return 0; }
  const uint32_t cap = force_stride ? std::max<uint32_t>(1, need / 3) : (uint32_t)(std::min(occ, per_sm_want) * c.num_sms);   // (test switch: about three cameras per warp)
  return (int)std::max<uint32_t>(1, std::min<uint32_t>(need, std::min<uint32_t>(cap, (uint32_t)(std::min(occ, per_sm_want) * c.num_sms))));
}
static int pcg_tail_plan(Ctx& c, bool& one) {
  switch (c.dc) {
    case 6: return tail_plan_dc<6>(c, one);
    case 9: return tail_plan_dc<9>(c, one);
    case 10: return tail_plan_dc<10>(c, one);
    case 11: return tail_plan_dc<11>(c, one);
    case 12: return tail_plan_dc<12>(c, one);
    case 14: return tail_plan_dc<14>(c, one);
    case 15: return tail_plan_dc<15>(c, one);
    default: return 0;
  }
}

template <int DC>
static cudaError_t launch_tail_dc(Ctx& c, const TailArgs& a, int ctas, bool one, bool pdl) {
  return one ? launch_dep(pcg_tail_kernel<DC, true>, (unsigned)ctas, TAIL_THREADS, 0, c.stream, pdl, a)
             : launch_dep(pcg_tail_kernel<DC, false>, (unsigned)ctas, TAIL_THREADS, 0, c.stream, pdl, a);
}
static apex_status launch_pcg_tail(Ctx& c, const TailArgs& a, int ctas, bool one, bool pdl) {
  cudaError_t e;
  switch (c.dc) {
    case 6: e = launch_tail_dc<6>(c, a, ctas, one, pdl); break;
    case 9: e = launch_tail_dc<9>(c, a, ctas, one, pdl); break;
    case 10: e = launch_tail_dc<10>(c, a, ctas, one, pdl); break;
    case 11: e = launch_tail_dc<11>(c, a, ctas, one, pdl); break;
    case 12: e = launch_tail_dc<12>(c, a, ctas, one, pdl); break;
    case 14: e = launch_tail_dc<14>(c, a, ctas, one, pdl); break;
    case 15: e = launch_tail_dc<15>(c, a, ctas, one, pdl); break;
    default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
  }
  APEX_CUDA_TRY(c, e);
  c.launches++;
  return APEX_OK;
}

// ----------------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------------
static SchurArgs make_schur_args(Ctx& c, const double* x, double* y, int check_done) {
  SchurArgs a;
  a.tiles = c.tiles.p; a.slot_cam = c.slot_cam.p; a.slot_lp = c.slot_lp.p; a.pt_slot0 = c.pt_slot0.p; a.pt_cnt = c.pt_cnt.p;
  a.J = c.J.p; a.hinv = c.hinv.p; a.gp = c.gp.p; a.x = x; a.y = y; a.step_pt = c.step_pt.p;
  a.npl = c.npl; a.check_done = check_done; a.st = c.state.p;
  a.ntiles = c.ntiles;
  a.chunk_desc = c.chunk_desc.p; a.cslot_meta = c.cslot_meta.p; a.cpt_meta = c.cpt_meta.p; a.xpad = c.xpad.p;
  a.cslot_widx = c.cslot_widx.p; a.win_desc = c.win_desc.p; a.range_win0 = c.range_win0.p; a.win_cams = c.win_cams.p;
  a.window = c.mv_window; a.nnormal = c.nnormal_chunks; a.partial = c.det_partial.p; a.win_dst = c.win_dst.p;
  const char* dbg = getenv("APEX_DEBUG_MATVEC");
  a.debug = dbg ? atoi(dbg) : 0;
  return a;
}

// Shared memory of the chunk kernel for a window of W cameras, and the CTAs per SM it leaves. problem_upload calls this
// before it builds the layout: the number of ranges is the number of resident CTAs of the operator (MATVEC) instantiation.
// The Jacobian stage (TMA prefetch) is used by the operator when a chunk's planes + the window leave room for two CTAs per SM
// (dc <= 10); the once-per-iteration modes (reduced gradient, back-substitution) always load directly.
template <int DC>
constexpr bool mv_staged_fits() { return chunk_smem_bytes<DC, MODE_MATVEC, true>(TILE) <= 110 * 1024; }
static bool mv_staged_wanted() { const char* e = getenv("APEX_MV_STAGED"); return !e || atoi(e) != 0; }

template <int DC>
static void plan_dc(uint32_t& W, bool& staged) {
  W = mv_window_cameras(DC);
  staged = mv_staged_fits<DC>() && mv_staged_wanted();
  if (staged) {  // stage + window fill what two CTAs per SM can have (113 KB each): dc = 9 -> 616 cameras
    if (!getenv("APEX_MV_WINDOW")) W = 2048;
    while (W > (uint32_t)TILE && chunk_smem_bytes<DC, MODE_MATVEC, true>(W) > 113 * 1024) W -= 8;
  }
}
// host-only: window width / staging the chunk kernel will use for a camera block of dc (also for apex_layout_stats_compute)
void schur_plan(int dc, uint32_t& W, bool& staged) {
  switch (dc) {
    case 6: plan_dc<6>(W, staged); break;
    case 9: plan_dc<9>(W, staged); break;
    case 10: plan_dc<10>(W, staged); break;
    case 11: plan_dc<11>(W, staged); break;
    case 12: plan_dc<12>(W, staged); break;
    case 14: plan_dc<14>(W, staged); break;
    case 15: plan_dc<15>(W, staged); break;
    default: W = mv_window_cameras(dc); staged = false;
  }
}

template <int DC>
static apex_status configure_dc(Ctx& c) {
  uint32_t W;
  plan_dc<DC>(W, c.mv_staged);
  c.mv_window = W;
  auto set = [&](const void* fn, size_t bytes) { return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); };
  APEX_CUDA_TRY(c, set((const void*)schur_chunk_kernel<DC, MODE_MATVEC, true, false>, chunk_smem_bytes<DC, MODE_MATVEC, false>(W)));
  APEX_CUDA_TRY(c, set((const void*)schur_chunk_kernel<DC, MODE_MATVEC, false, false>, chunk_smem_bytes<DC, MODE_MATVEC, false>(W)));
  if constexpr (mv_staged_fits<DC>()) {
    APEX_CUDA_TRY(c, set((const void*)schur_chunk_kernel<DC, MODE_MATVEC, true, true>, chunk_smem_bytes<DC, MODE_MATVEC, true>(W)));
    APEX_CUDA_TRY(c, set((const void*)schur_chunk_kernel<DC, MODE_MATVEC, false, true>, chunk_smem_bytes<DC, MODE_MATVEC, true>(W)));
  }
  APEX_CUDA_TRY(c, set((const void*)schur_chunk_kernel<DC, MODE_RHS, true, false>, chunk_smem_bytes<DC, MODE_RHS, false>(W)));
  APEX_CUDA_TRY(c, set((const void*)schur_chunk_kernel<DC, MODE_RHS, false, false>, chunk_smem_bytes<DC, MODE_RHS, false>(W)));
  APEX_CUDA_TRY(c, set((const void*)schur_chunk_kernel<DC, MODE_BACKSUB, false, false>, chunk_smem_bytes<DC, MODE_BACKSUB, false>(W)));
  int per_sm = 0;
  if constexpr (mv_staged_fits<DC>()) {
    if (c.mv_staged) APEX_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, schur_chunk_kernel<DC, MODE_MATVEC, true, true>, TILE, chunk_smem_bytes<DC, MODE_MATVEC, true>(W)));
  }
  if (!c.mv_staged) APEX_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, schur_chunk_kernel<DC, MODE_MATVEC, true, false>, TILE, chunk_smem_bytes<DC, MODE_MATVEC, false>(W)));
  if (per_sm < 1) { c.err = "chunk kernel does not fit on an SM"; return APEX_ERR_UNSUPPORTED; }
  c.mv_ctas_per_sm = (uint32_t)per_sm;
  if (const char* e = getenv("APEX_MV_CTAS_PER_SM")) c.mv_ctas_per_sm = (uint32_t)std::max(1, std::min(atoi(e), per_sm));
  return APEX_OK;
}
apex_status schur_configure(Ctx& c) {
  switch (c.dc) {
    case 6: return configure_dc<6>(c);
    case 9: return configure_dc<9>(c);
    case 10: return configure_dc<10>(c);
    case 11: return configure_dc<11>(c);
    case 12: return configure_dc<12>(c);
    case 14: return configure_dc<14>(c);
    case 15: return configure_dc<15>(c);
    default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
  }
}

template <int DC>
static apex_status launch_tiles_dc(Ctx& c, int mode, const SchurArgs& a0) {
  SchurArgs a = a0;
  const bool det = c.mv_det && mode != MODE_BACKSUB;
  // normal chunks through the chunk kernel, landmarks with more than 256 observations through the tile kernel
  if (c.nnormal_chunks) {
    const unsigned grid = c.mv_nranges;
    const uint32_t W = c.mv_window;
    switch (mode) {
      case MODE_MATVEC:
        if constexpr (mv_staged_fits<DC>()) {
          if (c.mv_staged) {
            if (det) APEX_CUDA_TRY(c, launch_dep(schur_chunk_kernel<DC, MODE_MATVEC, true, true>, grid, TILE, chunk_smem_bytes<DC, MODE_MATVEC, true>(W), c.stream, c.mv_pdl, a));
            else APEX_CUDA_TRY(c, launch_dep(schur_chunk_kernel<DC, MODE_MATVEC, false, true>, grid, TILE, chunk_smem_bytes<DC, MODE_MATVEC, true>(W), c.stream, c.mv_pdl, a));
            break;
          }
        }
        if (det) APEX_CUDA_TRY(c, launch_dep(schur_chunk_kernel<DC, MODE_MATVEC, true, false>, grid, TILE, chunk_smem_bytes<DC, MODE_MATVEC, false>(W), c.stream, c.mv_pdl, a));
        else APEX_CUDA_TRY(c, launch_dep(schur_chunk_kernel<DC, MODE_MATVEC, false, false>, grid, TILE, chunk_smem_bytes<DC, MODE_MATVEC, false>(W), c.stream, c.mv_pdl, a));
        break;
      case MODE_RHS:
        if (det) schur_chunk_kernel<DC, MODE_RHS, true, false><<<grid, TILE, chunk_smem_bytes<DC, MODE_RHS, false>(W), c.stream>>>(a);
        else schur_chunk_kernel<DC, MODE_RHS, false, false><<<grid, TILE, chunk_smem_bytes<DC, MODE_RHS, false>(W), c.stream>>>(a);
        break;
      default: schur_chunk_kernel<DC, MODE_BACKSUB, false, false><<<grid, TILE, chunk_smem_bytes<DC, MODE_BACKSUB, false>(W), c.stream>>>(a); break;
    }
    c.launches++;
    if (det && !c.mv_defer_reduce) {
      det_reduce_kernel<DC><<<(c.ncam + 7) / 8, 256, 0, c.stream>>>(c.det_partial.p, c.cam_row_start.p, a.y, c.state.p, c.ncam, a.check_done);
      c.launches++;
    }
  }
  if (c.ngiant) {
    a.tiles = c.giant_tiles.p;
    a.ntiles = c.ngiant;
    a.debug = 0;
    switch (mode) {
      case MODE_MATVEC: schur_tile_kernel<DC, MODE_MATVEC><<<c.ngiant, TILE, 0, c.stream>>>(a); break;
      case MODE_RHS: schur_tile_kernel<DC, MODE_RHS><<<c.ngiant, TILE, 0, c.stream>>>(a); break;
      default: schur_tile_kernel<DC, MODE_BACKSUB><<<c.ngiant, TILE, 0, c.stream>>>(a); break;
    }
    c.launches++;
  }
  return APEX_OK;
}

apex_status launch_schur_tiles(Ctx& c, int mode, const double* x, double* y, int check_done, bool xpad_ready) {
  if (c.ntiles == 0) return APEX_OK;
  if (mode != MODE_RHS && !xpad_ready) {  // the chunk kernel gathers x from the padded copy
    const int xs = xpad_stride(c.dc);
    pad_x_kernel<<<(c.ncam * xs + 255) / 256, 256, 0, c.stream>>>(x, c.xpad.p, c.ncam, c.dc, xs, c.state.p, check_done);
    c.launches++;
  }
  SchurArgs a = make_schur_args(c, x, y, check_done);
  apex_status st;
  switch (c.dc) {
    case 6: st = launch_tiles_dc<6>(c, mode, a); break;
    case 9: st = launch_tiles_dc<9>(c, mode, a); break;
    case 10: st = launch_tiles_dc<10>(c, mode, a); break;
    case 12: st = launch_tiles_dc<12>(c, mode, a); break;
    case 11: st = launch_tiles_dc<11>(c, mode, a); break;
    case 14: st = launch_tiles_dc<14>(c, mode, a); break;
    case 15: st = launch_tiles_dc<15>(c, mode, a); break;
    default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
  }
  APEX_TRY(st);
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

// y = (H_cc + lambda I) x on rank 0, 0 elsewhere (the all-reduce that follows the chunk kernel adds it once)
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

// this rank's part of y = S x, before the all-reduce: y = (H_cc + lambda I) x, then the chunk / tile kernels add
// -H_cp Hpp^-1 H_cp^T x to it
apex_status schur_operator_local(Ctx& c, const double* x, double* y, int check_done, bool xpad_ready) {
  APEX_TRY(launch_hcc_apply(c, x, y, check_done));
  cudaEvent_t* evp = (c.prof && c.ntiles) ? prof_pair(c.ev_pool, c.ev_mv_used++) : nullptr;
  if (evp) cudaEventRecord(evp[0], c.stream);
  apex_status st = launch_schur_tiles(c, MODE_MATVEC, x, y, check_done, xpad_ready);
  if (evp) cudaEventRecord(evp[1], c.stream);
  return st;
}

// full operator y = S x (all ranks end with the same y)
apex_status schur_operator(Ctx& c, const double* x, double* y, int check_done, bool xpad_ready) {
  APEX_TRY(schur_operator_local(c, x, y, check_done, xpad_ready));
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
  static_assert(sizeof(unsigned long long) == sizeof(double), "exchange slots live in red_scratch");
  pcg_init_kernel<<<1, 1024, 0, s>>>(c.vb.p, c.pinv.p, c.step_cam.p, c.vr.p, c.vz.p, c.vp.p, c.state.p, n, c.dc, c.K, cg_max_it, cg_tol,
                                     reinterpret_cast<unsigned long long*>(c.red_scratch.p), 9u * TAIL_MAX_CTAS);   // (red_scratch holds >= 8 256 words)
  c.launches++;
  APEX_CUDA_TRY(c, cudaGetLastError());
  // PCG: iterations are enqueued in batches of BATCH; every kernel of an iteration is a no-op once the device-side
  // `pcg_done` flag is set (also set at pcg_max), so the host only polls the flag between batches. A batch is
  // captured once per uploaded problem into a CUDA graph (kernels + the NCCL all-reduce) and replayed with one
  // launch: the inner loop is launch-bound on small shards (8 GPUs: ~35 us of kernels per iteration).
  const int BATCH = 10;
  int it_count = 0;  // iteration index within this solve: selects the half of the peer buffer (BATCH is even)
  if (getenv("APEX_TAIL_TRACE") && !c.tail_trace.p) { APEX_CUDA_TRY(c, c.tail_trace.alloc(4096 * 8)); APEX_CUDA_TRY(c, cudaMemsetAsync(c.tail_trace.p, 0, 4096 * 8 * sizeof(long long), s)); }
  bool tail_one = false;
  const int tail_ctas = (c.nranks == 1 || c.p2p_ok) ? pcg_tail_plan(c, tail_one) : 0;
  // operator and tail as programmatic dependents of each other (note in front of pcg_tail_kernel); APEX_PDL=0: plain stream order
  const char* pdl_env = getenv("APEX_PDL");
  const bool pdl = tail_ctas > 0 && !(pdl_env && atoi(pdl_env) == 0);
  if (tail_ctas && cg_max_it > 0) {  // first direction p = z and y0 = (H_cc + lambda I) p into half 0; later ones come from the tail
    const int xs = xpad_stride(c.dc);
    double* y0 = c.p2p_ok ? c.arbuf.p : c.vy.p;
    pcg_dir_hcc_kernel<<<(c.ncam + PCG_CAMS - 1) / PCG_CAMS, PCG_THREADS, 0, s>>>(c.vz.p, c.vp.p, c.xpad.p, c.hcc.p, y0, c.state.p, c.ncam, c.dc, xs, c.rank == 0 ? 1 : 0);
    c.launches++;
    APEX_CUDA_TRY(c, cudaGetLastError());
  }
  auto enqueue_iteration = [&]() -> apex_status {
    const int xs = xpad_stride(c.dc);
    const unsigned gp = (n + PCG_THREADS - 1) / PCG_THREADS, gu = (c.ncam + PCG_CAMS - 1) / PCG_CAMS;
    const int par = it_count++ & 1;
    if (tail_ctas) {
      // fused path: [operator, tail] per iteration; the first direction / y0 were produced before the loop; the second pass
      // of the deterministic flush runs inside the tail
      double* y0 = c.p2p_ok ? c.arbuf.p + (size_t)par * n : c.vy.p;
      cudaEvent_t* evp = (c.prof && c.ntiles) ? prof_pair(c.ev_pool, c.ev_mv_used++) : nullptr;
      if (evp) cudaEventRecord(evp[0], s);
      c.mv_defer_reduce = true;
      c.mv_pdl = pdl;
      apex_status ost = launch_schur_tiles(c, MODE_MATVEC, c.vp.p, y0, 1, true);
      c.mv_defer_reduce = false;
      c.mv_pdl = false;
      APEX_TRY(ost);
      if (evp) cudaEventRecord(evp[1], s);
      TailArgs ta{};
      if (c.p2p_ok) { ta.peer_buf = c.d_peer_buf.p; ta.peer_flags = c.d_peer_flags.p; ta.flags = c.arflags.p; }
      ta.par = par; ta.nranks = c.nranks; ta.rank = c.rank;
      ta.ylocal = y0;
      if (c.mv_det && c.nnormal_chunks) { ta.partial = c.det_partial.p; ta.cam_row_start = c.cam_row_start.p; }
      ta.y = c.vy.p;
      ta.y0_next = c.p2p_ok ? c.arbuf.p + (size_t)(par ^ 1) * n : c.vy.p;
      ta.p = c.vp.p; ta.x = c.step_cam.p; ta.r = c.vr.p; ta.z = c.vz.p; ta.xpad = c.xpad.p;
      ta.pinv = c.pinv.p; ta.hcc = c.hcc.p; ta.slots = reinterpret_cast<unsigned long long*>(c.red_scratch.p); ta.st = c.state.p;
      ta.ncam = c.ncam; ta.K = c.K; ta.xs = xs; ta.add_hcc = c.rank == 0 ? 1 : 0;
      ta.trace = c.tail_trace.p;
      APEX_TRY(launch_pcg_tail(c, ta, tail_ctas, tail_one, pdl));
      return APEX_OK;
    }
    {
    // chunk-kernel path: p / y0 in one kernel, the operator reduces into y0, then (ranks > 1) the peer-memory
    // all-reduce fused with p.Ap, or NCCL when peer mapping is unavailable
    double* y0 = c.p2p_ok ? c.arbuf.p + (size_t)par * n : c.vy.p;
    pcg_dir_hcc_kernel<<<gu, PCG_THREADS, 0, s>>>(c.vz.p, c.vp.p, c.xpad.p, c.hcc.p, y0, c.state.p, c.ncam, c.dc, xs, c.rank == 0 ? 1 : 0);
    c.launches++;
    cudaEvent_t* evp = (c.prof && c.ntiles) ? prof_pair(c.ev_pool, c.ev_mv_used++) : nullptr;
    if (evp) cudaEventRecord(evp[0], s);
    APEX_TRY(launch_schur_tiles(c, MODE_MATVEC, c.vp.p, y0, 1, true));
    if (evp) cudaEventRecord(evp[1], s);
    if (c.p2p_ok) {
      ar_reduce_pap_kernel<<<gp, PCG_THREADS, 0, s>>>(c.d_peer_buf.p, c.d_peer_flags.p, c.arflags.p, par, c.nranks, c.rank, c.vp.p, c.vy.p,
                                                      c.red_scratch.p, c.state.p, n);
    } else {
      APEX_TRY(allreduce_sum(c, c.vy.p, n));
      pcg_pap_kernel<<<gp, PCG_THREADS, 0, s>>>(c.vp.p, c.vy.p, c.red_scratch.p, c.state.p, n);
    }
    }
    pcg_update_kernel<<<gu, PCG_THREADS, 0, s>>>(c.vy.p, c.pinv.p, c.vp.p, c.step_cam.p, c.vr.p, c.vz.p, c.red_scratch.p, c.state.p, c.ncam, c.dc, c.K);
    c.launches += 2;
    return APEX_OK;
  };
  const bool use_graph = !c.prof && !getenv("APEX_NO_GRAPH") && cg_max_it > 0;
  if (use_graph && !c.pcg_graph_exec) {
    const int64_t l0 = c.launches;
    cudaGraph_t graph = nullptr;
    APEX_CUDA_TRY(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    apex_status st = APEX_OK;
    for (int i = 0; i < BATCH && st == APEX_OK; ++i) st = enqueue_iteration();
    it_count = 0;
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    c.pcg_graph_launches = c.launches - l0;
    c.launches = l0;
    if (st != APEX_OK) { if (graph) cudaGraphDestroy(graph); return st; }
    APEX_CUDA_TRY(c, ce);
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    APEX_CUDA_TRY(c, ce);
    c.pcg_graph_exec = exec;
  }
  if (use_graph) {
    // Two batches in flight: batch k+1 is enqueued BEFORE the host looks at the done flag of batch k, so the device never waits for
    // the host between batches (a stream synchronisation + graph launch is 15-20 us of idle GPU, sixteen times per LM iteration:
    // 3 % of an iteration on an eighth of the Venice shape). The price: when PCG ends inside batch k, batch k+1 runs as no-ops
    // (every kernel returns on pcg_done; ~7 us per iteration); at the iteration cap nothing is wasted (the cap is reached on a
    // batch boundary of the enqueue count).
    if (!c.h_batch_done) APEX_CUDA_TRY(c, cudaHostAlloc((void**)&c.h_batch_done, 2 * sizeof(int32_t), cudaHostAllocDefault));
    for (cudaEvent_t& e : c.ev_batch) if (!e) APEX_CUDA_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int enq = 0;
    auto launch_batch = [&](int slot) -> apex_status {
      APEX_CUDA_TRY(c, cudaGraphLaunch((cudaGraphExec_t)c.pcg_graph_exec, s));
      c.launches += c.pcg_graph_launches;
      enq += BATCH;
      APEX_CUDA_TRY(c, cudaMemcpyAsync(&c.h_batch_done[slot], &c.state.p->pcg_done, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      APEX_CUDA_TRY(c, cudaEventRecord(c.ev_batch[slot], s));
      return APEX_OK;
    };
    if (enq < cg_max_it) {
      APEX_TRY(launch_batch(0));
      for (int cur = 0;; cur ^= 1) {
        const bool more = enq < cg_max_it;
        if (more) APEX_TRY(launch_batch(cur ^ 1));
        APEX_CUDA_TRY(c, cudaEventSynchronize(c.ev_batch[cur]));
        if (c.h_batch_done[cur] || !more) break;
      }
    }
    APEX_TRY(sync_state(c));   // (also waits for a batch enqueued ahead)
  } else {
    int enq = 0;
    while (enq < cg_max_it) {
      const int nb = std::min(BATCH, cg_max_it - enq);
      for (int i = 0; i < nb; ++i) APEX_TRY(enqueue_iteration());
      APEX_CUDA_TRY(c, cudaGetLastError());
      enq += nb;
      APEX_TRY(sync_state(c));
      if (c.h_state->pcg_done) break;
    }
  }
  if (cg_max_it <= 0) APEX_TRY(sync_state(c));
  c.last_pcg_iters = c.h_state->pcg_iters;
  if (c.tail_trace.p && c.last_pcg_iters > 20) {   // development probe: where a PCG iteration's time goes, from CTA 0's stamps
    std::vector<long long> h(4096 * 8);
    cudaMemcpy(h.data(), c.tail_trace.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const int n1 = (int)std::min<int64_t>(c.last_pcg_iters - 1, 4000);
    double d[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int i = 10; i < n1; ++i) {
      const long long* a0 = &h[(size_t)i * 8];
      const long long* a1 = &h[(size_t)(i + 1) * 8];
      if (!a0[0] || !a0[7] || !a1[0]) continue;
      long long prev = a0[0];
      for (int k = 1; k < 8; ++k) { if (a0[k]) { d[k] += (double)(a0[k] - prev); prev = a0[k]; } }
      d[8] += (double)(a1[0] - a0[7]);   // end of this tail -> start of the next = operator + launch gaps
      ++cnt;
    }
    if (cnt) fprintf(stderr, "[tail trace rank %d] per iteration (us, %d samples): A %.2f | (barrier A %.2f) | peers' flags %.2f | B + publish %.2f | gather B %.2f | C + gather C %.2f | D %.2f || end of tail -> next tail past its wait (operator + gaps) %.2f\n", c.rank, cnt,
                     d[1] / cnt / 1e3, d[2] / cnt / 1e3, d[3] / cnt / 1e3, d[4] / cnt / 1e3, d[5] / cnt / 1e3, d[6] / cnt / 1e3, d[7] / cnt / 1e3, d[8] / cnt / 1e3);
  }
  if (c.nranks > 1) {  // a timed-out exchange on one rank ends the solve on all of them
    APEX_TRY(agree_error_flags(c));
    APEX_TRY(sync_state(c));
  }
  if (c.h_state->ar_timeout) { c.err = "peer all-reduce: a rank never published its partial result"; return APEX_ERR_NCCL; }
  if (c.h_state->singular_landmark) { c.err = "Landmark block singular"; return APEX_ERR_SINGULAR_MATRIX; }
  APEX_TRY(launch_schur_tiles(c, MODE_BACKSUB, c.step_cam.p, nullptr, 0));
  return APEX_OK;
}

}  // namespace apex
