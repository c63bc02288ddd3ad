// apex_ctx.h — host-side context of the B200 bundle-adjustment path: device buffers, the static
// observation structure built at upload, and the launchers of every kernel group.
//
// Data layout in HBM (all FP64, indices u32; DESIGN.md section 2 has the full description):
//   * Observations are sharded with their landmark (block-cyclic: blocks of 128 consecutive landmarks go round the
//     ranks, so every rank sees every camera neighbourhood) and stored POINT-MAJOR
//     in "slots". Slots are grouped in chunks of TILE=256; a tile is one chunk holding a run of whole
//     landmarks (<=256 observations, <=256 landmarks), or, for a landmark with more than 256
//     observations, several consecutive chunks holding only that landmark. One CTA processes one tile;
//     per-landmark sums (H_pp, g_p, H_cp^T x) are segmented sums inside the CTA and never cross CTAs.
//   * Per slot: camera index, landmark index local to the tile, measured pixel, and after linearization
//     the loss-corrected residual (2 planes) and Jacobian (2*(dc+3) planes), chunk-blocked:
//     J[(chunk*NP + plane)*TILE + lane] so a chunk's Jacobians are one contiguous 8*NP*256-byte run
//     and every warp access is a full 256-byte line pair.
//   * SPLIT SLOT ORDER inside a normal chunk: the landmark half of the Jacobian (planes 2dc..2dc+5), the residual and the
//     pixel sit at the observation's POINT-MAJOR lane; the camera half (planes 0..2dc-1) sits at the observation's
//     CAMERA-SORTED lane (observations of the chunk sorted by camera). cslot_meta holds both permutations. Chunks of a
//     landmark with more than 256 observations keep the camera half at the point-major lane.
//   * Per landmark (SoA planes over local landmarks): H_pp (6, symmetric), g_p (3), (H_pp+lambda I)^-1 (6).
//   * Camera side (replicated on every rank): pose[ncam][7], intr[ncam][K], H_cc[ncam][dc][dc], g_c[ncam][dc],
//     preconditioner blocks, PCG vectors of ncam*dc doubles.
//   * A CAMERA-MAJOR copy of (pixel, landmark index) feeds the camera-side accumulation kernels, which
//     re-evaluate the projection instead of gathering Jacobians (cheaper than a strided re-read).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../include/apex_gpu.h"

namespace apex {

constexpr int TILE = 256;            // slots per chunk = threads per CTA of the tile kernels
constexpr uint32_t PAD_CAM = 0xFFFFFFFFu;
constexpr int CAM_THREADS = 128;     // CTA size of the camera-major accumulation kernels
constexpr int CAM_CHUNK = 2048;      // observations per camera work item
constexpr int MAX_DC = 15;
// per-camera stride (doubles) of the padded copy of the operator input: 32-byte aligned blocks for 256-bit gathers
__host__ __device__ constexpr int xpad_stride(int dc) { return (dc + 3) & ~3; }
constexpr int MAX_K = 9;

struct TileDesc {
  uint32_t pt0;      // first local landmark
  uint32_t npt;      // landmarks in the tile
  uint32_t chunk0;   // first chunk
  uint32_t nchunks;  // 1 for a normal tile; >1 only when npt == 1
};

// Per-chunk descriptor of the chunk kernel.
struct ChunkDesc {
  uint32_t pt0, npt;   // landmarks of the chunk
  uint32_t nwseg;      // warp-segments: runs of one camera among the camera-sorted lanes, cut at warp boundaries
  uint32_t nobs;       // observations of the chunk (point-major lanes / camera-sorted lanes 0..nobs-1 are in use)
};
// cslot_meta[chunk*TILE + t] = {x, y}:
//   x       camera of CAMERA-SORTED lane t (PAD_CAM for t >= nobs)
//   y[ 7:0] chunk-local landmark of POINT-MAJOR lane t
//   y[15:8] camera-sorted lane of the observation at point-major lane t
//   y[23:16] point-major lane of the observation at camera-sorted lane t
//   y[24]   camera-sorted lane t is the first lane of a warp and CONTINUES the run of the previous warp's last lane
//   y[25]   ... and is the first continuation of that run (the previous warp did not start with the same run)
//   y[28:26] number of consecutive warps, from this one on, that start with a continuation of the same run (1..7)
// cslot_widx[chunk*TILE + t]: index of the camera of camera-sorted lane t in the camera list of the chunk's WINDOW.
//
// Operator work distribution (schur.cu). The normal chunks are cut into `nranges` contiguous ranges of equal length, one per
// resident CTA of the chunk kernel; a range is cut into WINDOWS: maximal runs of chunks whose observations touch at most
// W distinct cameras (W >= 256, so a chunk always fits). The CTA keeps the window's rows of the operator result in shared
// memory and flushes them once per window.
struct WinDesc {
  uint32_t chunk_begin, chunk_end;  // chunks of the window
  uint32_t cam0, ncams;             // its sorted camera list = win_cams[cam0 .. cam0+ncams); also its rows of the partial results
};
constexpr uint32_t CONT_BIT = 1u << 24, CONT_FIRST_BIT = 1u << 25;
constexpr int CONT_LEN_SHIFT = 26;
// cameras per window: 32 KB of shared memory for the window's rows of y, never below one chunk's worth of cameras
inline uint32_t mv_window_cameras(int dc) {
  if (const char* e = getenv("APEX_MV_WINDOW")) { const int w = atoi(e); if (w >= TILE) return (uint32_t)std::min(w, 4096); }
  return (uint32_t)std::max(TILE, 32768 / (8 * dc));
}
// Landmark ownership: blocks of SHARD_BLOCK consecutive landmarks are dealt round-robin to the ranks. A contiguous
// split would give each rank one camera neighbourhood of a locality-ordered reconstruction, i.e. 1/nranks of the
// camera rows to reduce into (measured: 0.084 ms vs 0.059 ms per operator application on an eighth of Venice-1778).
constexpr uint32_t SHARD_BLOCK = 128;
struct ShardMap {
  uint32_t npts = 0, nranks = 1, rank = 0;
  __host__ __device__ bool owns(uint32_t p) const { return (p / SHARD_BLOCK) % nranks == rank; }
  __host__ __device__ uint32_t to_local(uint32_t p) const { return ((p / SHARD_BLOCK - rank) / nranks) * SHARD_BLOCK + p % SHARD_BLOCK; }
  __host__ __device__ uint32_t to_global(uint32_t lp) const { return ((lp / SHARD_BLOCK) * nranks + rank) * SHARD_BLOCK + lp % SHARD_BLOCK; }
  uint32_t count() const {  // landmarks owned by `rank`
    const uint32_t nblk = (npts + SHARD_BLOCK - 1) / SHARD_BLOCK;
    if (nblk == 0 || rank >= nblk) return 0;
    const uint32_t mine = (nblk - rank + nranks - 1) / nranks;
    const uint32_t last_blk = (mine - 1) * nranks + rank;  // my last block
    const uint32_t last_size = last_blk == nblk - 1 ? npts - last_blk * SHARD_BLOCK : SHARD_BLOCK;
    return (mine - 1) * SHARD_BLOCK + last_size;
  }
};
constexpr int MAX_TILE_PTS = 128;     // landmarks per normal tile

struct CamItem { uint32_t cam, begin, end, pad; };
struct PairBlock { uint32_t cam_i, cam_j, begin, end; };   // block (cam_i >= cam_j) of the reduced camera system: its pairs [begin, end)
struct LossSpecPod { int id; double p0, p1; };   // layout of ba_device.cuh's LossSpec (one LossFunction instance on the device)  // [begin,end) in the camera-major arrays

// Scalars that live on the device (LM bookkeeping, PCG control, norms, error flags).
struct DevState {
  // LM state (update_damping / compute_step_quality / check_convergence)
  double damping, nu;
  double current_cost, new_cost, previous_cost;
  double predicted, rho;
  double grad_norm, step_norm, step_dot_grad, param_norm;
  int32_t accepted, status, iteration, pad0;
  // partial norms
  double g2_cam, s2_cam, sg_cam;          // camera side (replicated)
  double g2_pt, s2_pt, sg_pt;             // landmark side (rank-local, all-reduced)
  double pn2_cam, pn2_pt;
  double cost2_local;                     // sum r~^2 of the local observations
  // PCG
  double rz_old, pcg_tol, b_norm, r_norm;
  double pcg_alpha, pcg_beta;
  uint32_t ticket_a, ticket_b;            // last-block-done counters of the multi-CTA PCG kernels
  uint32_t pad4[4];
  int32_t pcg_iters, pcg_done, pcg_max, pad1;
  // errors
  unsigned long long ar_seq;              // sequence number of the peer-memory all-reduce (comm.cu)
  int32_t ar_timeout, pad3;               // a peer never published its flag (bounded spin expired)
  int32_t singular_landmark;              // a landmark block could not be inverted
  int32_t chol_fail;                      // first failing column + 1 of the dense Cholesky
  uint32_t tail_tag;                      // tag of the fused PCG tail's last sum exchange in this solve (schur.cu, tail_publish / tail_gather)
  int32_t pad2;
  // multi-rank agreement (agree_error_flags / the LM loop's timeout): every rank must take the same exit
  double agree[2];                        // {singular_landmark, ar_timeout} summed over the ranks
  double elapsed;                         // rank 0's wall clock, all-reduced, for the TIMEOUT test of check_convergence
};

// Host vector whose resize() leaves trivially constructible elements uninitialised: the layout build fills the big
// per-slot arrays from OpenMP workers (parallel first touch) instead of zero-filling 230 MB on one thread first.
template <typename T>
struct NoInitAlloc : std::allocator<T> {
  template <typename U> struct rebind { using other = NoInitAlloc<U>; };
  template <typename U, typename... A>
  void construct(U* p, A&&... a) {
    if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(p)) U;
    else ::new (static_cast<void*>(p)) U(static_cast<A&&>(a)...);
  }
};
template <typename T> using HostVec = std::vector<T, NoInitAlloc<T>>;

// Host staging array of apex_problem_upload: grows, never shrinks, lives as long as its context. With `pinned` set the
// memory is page-locked (cudaHostAlloc), so the upload's H2D copies run at link speed instead of being staged through the
// driver's bounce buffer (11 GB/s measured), and a context that is uploaded to again touches no fresh pages. The host-only
// entry points (apex_layout_stats_compute) use it unpinned. Same small API as the vectors it replaces.
template <typename T>
struct StageBuf {
  using value_type = T;
  T* p = nullptr;
  size_t n = 0, cap = 0;
  bool pinned = false, is_pinned_alloc = false;
  StageBuf() = default;
  StageBuf(const StageBuf&) = delete;
  StageBuf& operator=(const StageBuf&) = delete;
  ~StageBuf() { release(); }
  void release() {
    if (p) { if (is_pinned_alloc) cudaFreeHost(p); else free(p); }
    p = nullptr; n = cap = 0;
  }
  void resize(size_t count) {  // contents are undefined after a resize (every user fills what it reads)
    if (count > cap || (pinned && !is_pinned_alloc && p && count > 0)) {   // (a buffer that became "pinned" is re-allocated page-locked)
      const bool want_pinned = pinned;
      release();
      const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
      if (want_pinned && cudaHostAlloc(reinterpret_cast<void**>(&p), bytes, cudaHostAllocDefault) == cudaSuccess) is_pinned_alloc = true;
      else { cudaGetLastError(); p = static_cast<T*>(malloc(bytes)); is_pinned_alloc = false; }
      cap = p ? std::max(count, cap) : 0;
    }
    n = count <= cap ? count : 0;
  }
  T* begin() { return p; }
  T* end() { return p + n; }
  const T* begin() const { return p; }
  const T* end() const { return p + n; }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count; else p = nullptr;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct NcclApi;  // nccl_dyn.h

struct Ctx {
  std::string err;
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  // page-locked bounce ring of the upload (problem.cu): two buffers + the events of their last DMA
  static constexpr size_t BOUNCE_BYTES = (size_t)16 << 20;
  void* bounce[2] = {nullptr, nullptr};
  cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
  int bounce_next = 0;
  bool chol_attr_set = false;              // dynamic shared-memory limits of the Cholesky kernels set on this context's device
  cudaStream_t stream2 = nullptr;          // low-priority stream of the dense Cholesky's look-ahead (explicit.cu)
  std::vector<cudaEvent_t> chol_events;    // pairs per outer panel: panel factored / rest update done
  int64_t launches = 0;
  int num_sms = 148;
  void* nccl_comm = nullptr;

  // ---- problem (host copies of what the host needs) ----
  bool have_problem = false;
  int model = 0, K = 0, dc = 6, np = 18;  // np = 2*(dc+3) Jacobian planes
  uint32_t opt = 0;
  bool opt_intr = false, intr_vars = false;
  bool shared_intr = false;          // one intrinsics variable for all cameras (APEX_OPT_SHARED_INTRINSICS)
  uint32_t ncam = 0, npts = 0;
  uint64_t nobs = 0;
  uint64_t cam_dof_ref = 0;  // reference-layout camera dof (includes unreferenced intr columns)
  int loss_id = 0;
  double loss_p[4] = {0, 0, 0, 0};
  bool jacobi_on = false;            // linearisation applies the Jacobi column scaling (inside lm_solve with use_jacobi_scaling)
  bool per_obs_loss = false;         // per-block loss functions (apex_problem_desc::obs_loss)
  uint32_t npl = 0;                  // landmarks owned by this rank (block-cyclic, see ShardMap)
  uint64_t nobs_local = 0;
  ShardMap shard;
  uint32_t nchunks = 0, ntiles = 0, nitems = 0;
  uint32_t ngiant = 0, nnormal_chunks = 0;
  uint32_t mv_nranges = 0, mv_window = 0, mv_ctas_per_sm = 0;   // chunk kernel: ranges (= CTAs), cameras per window
  uint32_t mv_nwindows = 0;
  uint64_t mv_nrows = 0;           // sum of the windows' camera counts (rows of det_partial)
  bool mv_staged = false;          // operator kernel prefetches the Jacobian planes through a TMA-fed shared-memory stage
  bool mv_defer_reduce = false;    // the caller adds the partial rows itself (fused PCG tail)
  bool mv_pdl = false;             // launch the operator as a programmatic dependent of the kernel in front of it (PCG loop, schur.cu)
  bool mv_det = false;             // flush the windows as per-window partial rows + fixed-order second pass (bitwise reproducible)
  size_t nslots = 0;
  HostVec<uint64_t> slot_obs;      // slot -> caller's observation index (UINT64_MAX for padding)
  HostVec<uint8_t> slot_pos;       // slot -> lane of the camera half of its Jacobian inside the chunk
  std::shared_ptr<void> staging;   // the HostLayout of problem.cu, kept between uploads (pinned staging arrays)
  std::vector<uint32_t> h_pt_cnt;

  // ---- device: static structure ----
  DevBuf<TileDesc> tiles;
  DevBuf<uint32_t> slot_cam;
  DevBuf<uint16_t> slot_lp;
  DevBuf<TileDesc> giant_tiles;      // the tiles with nchunks > 1 (a landmark with more than 256 observations)
  DevBuf<ChunkDesc> chunk_desc;      // [nnormal_chunks]
  DevBuf<uint2> cslot_meta;          // [nnormal_chunks][TILE], see ChunkDesc
  DevBuf<uint32_t> cpt_meta;         // per landmark: chunk-local first slot | count << 16
  DevBuf<double> xpad;               // operator input at the padded per-camera stride
  DevBuf<uint16_t> cslot_widx;       // [nnormal_chunks][TILE] window-local camera index of the camera-sorted lanes
  DevBuf<WinDesc> win_desc;          // [mv_nwindows]
  DevBuf<uint32_t> range_win0;       // [mv_nranges + 1] first window of each range
  DevBuf<uint32_t> win_cams;         // [mv_nrows] sorted camera lists of the windows
  // deterministic flush: partial rows + per-camera row lists
  DevBuf<uint32_t> cam_row_start;    // [ncam+1] camera-major rows of det_partial: camera c owns rows [start[c], start[c+1]) ...
  DevBuf<uint32_t> win_dst;          // ... and entry i of win_cams flushes to row win_dst[i] (a camera's rows in (range, window) order)
  DevBuf<double> det_partial;        // [mv_nrows][dc]
  DevBuf<double> slot_uv;            // [chunk][2][TILE]
  DevBuf<uint8_t> slot_loss, cm_loss; // per-block loss index per slot / per camera-major observation (only with obs_loss)
  DevBuf<LossSpecPod> loss_tab;      // {id, p0, p1} per entry of the caller's loss_table
  DevBuf<uint32_t> pt_slot0, pt_cnt; // per local landmark
  DevBuf<CamItem> items;
  DevBuf<uint32_t> cam_item_start;   // [ncam+1]
  DevBuf<double> cm_uv;              // [2][nobs_local] camera-major
  DevBuf<uint32_t> cm_lp;            // [nobs_local] local landmark index
  DevBuf<uint8_t> pose_fixed, pt_fixed;
  DevBuf<uint16_t> intr_fixed;

  // ---- device: variables ----
  DevBuf<double> pose, intr, pt;     // pose[ncam][7], intr[ncam][K], pt[npl][3] (local landmarks only)
  DevBuf<double> pt_full;            // scratch for params_download with nranks > 1

  // ---- device: linearization ----
  DevBuf<double> J;                  // [chunk][np][TILE]
  DevBuf<double> R;                  // [chunk][2][TILE]
  DevBuf<double> hpp, gp, hinv;      // [6][npl], [3][npl], [6][npl]
  DevBuf<double> hcc;                // [ncam][dc][dc] followed by g_c [ncam][dc] (one all-reduce)
  double* gc = nullptr;              // = hcc.p + ncam*dc*dc
  DevBuf<double> partial;            // [nitems][nacc] camera work-item partial sums
  DevBuf<double> scale_cam, scale_pt; // Jacobi column scaling 1 / (1 + ||column||): [ncam][dc], [npl][3] (allocated on first use)
  DevBuf<double> sj;                 // [ncam][36 + K*K] Schur-Jacobi subtrahends
  DevBuf<double> pinv;               // [ncam][36 + K*K] preconditioner block inverses
  bool linearized = false;
  double lin_lambda = 0.0;
  bool have_step = false;            // step_cam / step_pt hold the step of a solve (apex_get_step)

  // ---- device: solve ----
  DevBuf<double> vb, vx, vr, vz, vp, vy;  // PCG vectors, ncam*dc each
  DevBuf<double> step_cam, step_pt;       // [ncam][dc], [npl][3]
  DevBuf<double> red_scratch;             // block partials of the deterministic reductions
  // formation of the explicit S by camera-pair blocks (built on the first explicit solve after an upload)
  bool pairs_ready = false;
  uint32_t npair_blocks = 0;
  DevBuf<PairBlock> pair_blocks;          // sorted by (cam_i, cam_j)
  DevBuf<uint2> pair_slots;               // {slot of observation i, slot of observation j} of the same landmark, cam_j <= cam_i
  DevBuf<uint32_t> slot_lpg;              // slot -> local landmark
  DevBuf<uint8_t> slot_cs8;               // slot -> lane of the camera half of its Jacobian inside its chunk
  DevBuf<double> S;                       // dense reduced camera system (explicit variants), n*n row-major
  DevBuf<double> E;                       // [chunk][3*dc][TILE] H_cp blocks for S formation
  DevBuf<double> dvec;                    // dense-solver work vectors
  DevBuf<long long> tail_trace;           // development probe APEX_TAIL_TRACE (schur.cu)
  DevBuf<double> sh_vec;                  // shared intrinsics: [gradient | step] in the reduced layout [poses 6 each | intrinsics K]
  DevBuf<double> l2flush;                 // > L2 buffer for apex_schur_matvec_bench
  DevBuf<DevState> state;
  DevState* h_state = nullptr;            // pinned mirror
  DevBuf<apex_iter_trace> trace;
  int64_t last_pcg_iters = 0;
  void* pcg_graph_exec = nullptr;       // cudaGraphExec_t of one batch of PCG iterations (rebuilt after every upload)
  int64_t pcg_graph_launches = 0;       // kernel launches one replay stands for

  // ---- observers (apex_add_observer): fed by lm_solve once per iteration ----
  std::vector<apex_observer> observers;
  apex_ctx* self = nullptr;                      // the handle the callbacks receive

  // ---- NVLink peer-memory all-reduce of the operator result (comm.cu) ----
  bool p2p_ok = false;
  size_t ar_n = 0;
  DevBuf<double> arbuf;                          // [2][ar_n], mapped by every peer
  DevBuf<unsigned long long> arflags;            // [nranks], slot r written by peer r
  std::vector<double*> peer_buf;                 // peer_buf[r] = rank r's arbuf in this process
  std::vector<unsigned long long*> peer_flags;
  DevBuf<double*> d_peer_buf;
  DevBuf<unsigned long long*> d_peer_flags;

  // ---- profiling (apex_profile_*) ----
  bool prof = false;
  std::vector<cudaEvent_t> ev_pool;   // pairs: [2i] start, [2i+1] stop
  size_t ev_mv_used = 0;              // event pairs used by operator launches since the last read
  std::vector<cudaEvent_t> ev_lin;    // pairs around launch_linearize
  size_t ev_lin_used = 0;
  std::vector<cudaEvent_t> ev_form, ev_chol;  // explicit variants: pairs around the formation of S / the dense factorisation
  size_t ev_form_used = 0, ev_chol_used = 0;
  uint64_t chol_n = 0;
  uint64_t upload_h2d_bytes = 0;      // H2D bytes of the last problem upload
  cudaEvent_t ev_lm0 = nullptr, ev_lm1 = nullptr;
  cudaEvent_t ev_batch[2] = {nullptr, nullptr};   // end of the PCG batches in flight (solve_implicit)
  int32_t* h_batch_done = nullptr;                // pinned [2]: pcg_done as of the end of those batches
  bool lm_timed = false;
};

// event pair `idx` of a pool, created on demand
inline cudaEvent_t* prof_pair(std::vector<cudaEvent_t>& pool, size_t idx) {
  while (pool.size() < 2 * (idx + 1)) { cudaEvent_t e = nullptr; cudaEventCreate(&e); pool.push_back(e); }
  return &pool[2 * idx];
}

// ---- error helpers --------------------------------------------------------------------------------
#define APEX_CUDA_TRY(ctx, expr)                                                                    \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      (ctx).err = std::string(#expr) + ": " + cudaGetErrorString(_e);                               \
      return APEX_ERR_CUDA;                                                                         \
    }                                                                                               \
  } while (0)

#define APEX_TRY(expr)                    \
  do {                                    \
    apex_status _s = (expr);              \
    if (_s != APEX_OK) return _s;         \
  } while (0)

// ---- launchers (one per kernel group; defined in the .cu files) --------------------------------------
// problem.cu
apex_status problem_upload(Ctx& c, const apex_problem_desc* d);
std::vector<double> gather_local_points(const Ctx& c, const double* pt_full);
apex_status layout_stats(const apex_problem_desc* d, int nranks, int rank, apex_layout_stats* out, std::string& err);
apex_status build_schur_pairs(Ctx& c);   // camera-pair blocks of the explicit S from the host copy of the layout (lazy, once per upload)

// linearize.cu
apex_status launch_normalize_poses(Ctx& c);
apex_status launch_linearize(Ctx& c);                       // K1 + K2 + K3 (lambda from state->damping)
apex_status launch_cost(Ctx& c, double* d_cost2_out);       // K1': writes sum r~^2 (local) to state->cost2_local
apex_status launch_schur_jacobi_blocks(Ctx& c, int kind);   // K5 build: fills pinv for preconditioner `kind`
// schur.cu
enum TileMode { MODE_MATVEC = 0, MODE_RHS = 1, MODE_BACKSUB = 2 };
apex_status launch_schur_tiles(Ctx& c, int mode, const double* x, double* y, int check_done, bool xpad_ready = false);
apex_status launch_hcc_apply(Ctx& c, const double* x, double* y, int check_done);  // y = (H_cc + lambda I) x on rank 0, else 0
apex_status schur_operator(Ctx& c, const double* x, double* y, int check_done, bool xpad_ready = false);    // y = S x, all-reduced
apex_status schur_operator_local(Ctx& c, const double* x, double* y, int check_done, bool xpad_ready = false);  // this rank's part
apex_status launch_reduced_gradient(Ctx& c, double* b);
apex_status solve_implicit(Ctx& c, int precond, int cg_max_it, double cg_tol);
void schur_plan(int dc, uint32_t& W, bool& staged);  // host-only: window width / Jacobian staging for a camera block of dc
apex_status schur_configure(Ctx& c);  // at upload: window width and CTAs per SM of the chunk kernel for the context's dc
// explicit.cu
apex_status solve_explicit(Ctx& c, bool use_pcg, int cg_max_it, double cg_tol);
apex_status dense_cholesky_bench(Ctx& c, uint32_t n, int reps, double* ms_out);
// lm.cu
apex_status launch_step_norms(Ctx& c);
apex_status launch_apply_step(Ctx& c, double sign, bool only_if_rejected);
apex_status launch_param_norm(Ctx& c);
apex_status lm_solve(Ctx& c, const apex_lm_config* cfg, apex_lm_result* res, apex_iter_trace* trace, int trace_cap);
// comm.cu
constexpr size_t AR_CAPACITY = (size_t)1 << 19;   // camera dofs the peer buffers are mapped for at context creation (2 x 4 MB)
apex_status setup_peer_allreduce(Ctx& c, size_t n);
void release_peer_allreduce(Ctx& c);
// comm (apex_gpu.cu)
apex_status allreduce_sum(Ctx& c, double* dev, size_t count);
apex_status sync_state(Ctx& c);  // copy DevState to the pinned mirror and wait
apex_status agree_error_flags(Ctx& c);  // nranks > 1: OR of the per-rank error flags (singular landmark block, peer all-reduce timeout) on every rank

}  // namespace apex
