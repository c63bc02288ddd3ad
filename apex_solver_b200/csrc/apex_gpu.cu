// apex_gpu.cu — the C ABI of include/apex_gpu.h over the sm_100a kernels. No CPU fallback: without a CUDA
// device every compute entry point returns APEX_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <new>

#include "apex_ctx.h"
#include "ba_device.cuh"
#include "kernels_common.cuh"
#include "nccl_dyn.h"

struct apex_ctx { apex::Ctx c; };

namespace apex {

apex_status allreduce_sum(Ctx& c, double* dev, size_t count) {
  if (c.nranks <= 1 || count == 0) return APEX_OK;
  NcclApi& api = nccl_api();
  int r = api.AllReduce(dev, dev, count, NCCL_DOUBLE, NCCL_SUM, (NcclComm)c.nccl_comm, c.stream);
  if (r != 0) { c.err = std::string("ncclAllReduce: ") + (api.GetErrorString ? api.GetErrorString(r) : "error"); return APEX_ERR_NCCL; }
  return APEX_OK;
}

apex_status sync_state(Ctx& c) {
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.h_state, c.state.p, sizeof(DevState), cudaMemcpyDeviceToHost, c.stream));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(c.stream));
  return APEX_OK;
}

// Error flags are raised per rank (a singular landmark block lives in one shard, a peer all-reduce times out on the rank
// that waited), but every rank has to leave the LM loop through the same exit - a rank that returned an error while its
// peers enter the next collective would leave them blocked in it. Sum the flags over the ranks and write them back.
__global__ void pack_error_flags_kernel(DevState* st) {
  st->agree[0] = st->singular_landmark ? 1.0 : 0.0;
  st->agree[1] = st->ar_timeout ? 1.0 : 0.0;
}
__global__ void unpack_error_flags_kernel(DevState* st) {
  if (st->agree[0] > 0.0) st->singular_landmark = 1;
  if (st->agree[1] > 0.0) st->ar_timeout = 1;
}
apex_status agree_error_flags(Ctx& c) {
  if (c.nranks <= 1) return APEX_OK;
  pack_error_flags_kernel<<<1, 1, 0, c.stream>>>(c.state.p);
  APEX_TRY(allreduce_sum(c, c.state.p->agree, 2));
  unpack_error_flags_kernel<<<1, 1, 0, c.stream>>>(c.state.p);
  c.launches += 2;
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

static apex_status set_damping(Ctx& c, double lambda) {
  APEX_CUDA_TRY(c, cudaMemcpyAsync(&c.state.p->damping, &lambda, sizeof(double), cudaMemcpyHostToDevice, c.stream));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(c.stream));  // `lambda` is a stack variable
  return APEX_OK;
}

static apex_status do_linearize(Ctx& c, double lambda) {
  APEX_TRY(set_damping(c, lambda));
  APEX_CUDA_TRY(c, cudaMemsetAsync(&c.state.p->singular_landmark, 0, sizeof(int32_t), c.stream));
  APEX_TRY(launch_linearize(c));
  c.linearized = true;
  c.lin_lambda = lambda;
  return APEX_OK;
}

static void lm_config_default(apex_lm_config* c) {  // levenberg_marquardt.rs:319-359
  std::memset(c, 0, sizeof(*c));
  c->schur_variant = APEX_SCHUR_EXPLICIT;  // SchurVariant::default() = Sparse
  c->schur_preconditioner = APEX_PRECOND_SCHUR_JACOBI;
  c->max_iterations = 50;
  c->cg_max_iterations = 200;  // explicit_schur.rs:211
  c->cost_tolerance = 1e-6; c->parameter_tolerance = 1e-8; c->gradient_tolerance = 1e-10;
  c->timeout_seconds = 0.0;
  c->damping = 1e-3; c->damping_min = 1e-12; c->damping_max = 1e12;
  c->damping_increase_factor = 10.0; c->damping_decrease_factor = 0.3; c->damping_nu = 2.0;
  c->trust_region_radius = 1e4; c->min_step_quality = 0.0; c->good_step_quality = 0.75;
  c->min_diagonal = 1e-6; c->max_diagonal = 1e32;
  c->min_cost_threshold = std::numeric_limits<double>::quiet_NaN();
  c->min_trust_region_radius = 1e-32;
  c->max_condition_number = std::numeric_limits<double>::quiet_NaN();
  c->min_relative_decrease = 1e-3;
  c->cg_tolerance = 1e-6;  // explicit_schur.rs:212
  c->use_jacobi_scaling = 0; c->compute_covariances = 0;
}

}  // namespace apex

using namespace apex;

#define CTX_OR_FAIL(ctx)                 \
  if (!(ctx)) return APEX_ERR_INVALID_STATE; \
  Ctx& c = (ctx)->c;                     \
  c.err.clear();                         \
  if (cudaSetDevice(c.device) != cudaSuccess) { c.err = "cudaSetDevice failed"; return APEX_ERR_CUDA; }

#define NEED_PROBLEM() \
  if (!c.have_problem) { c.err = "no problem uploaded"; return APEX_ERR_INVALID_STATE; }

extern "C" {

int32_t apex_abi_version(void) { return 100; }

int32_t apex_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void apex_lm_config_default(apex_lm_config* cfg) { lm_config_default(cfg); }

void apex_lm_config_for_bundle_adjustment(apex_lm_config* cfg) {  // levenberg_marquardt.rs:519-530
  lm_config_default(cfg);
  cfg->schur_variant = APEX_SCHUR_EXPLICIT_PCG;  // SchurVariant::Iterative as dispatched today
  cfg->schur_preconditioner = APEX_PRECOND_SCHUR_JACOBI;
  cfg->damping = 1e-3; cfg->max_iterations = 20;
  cfg->cost_tolerance = 1e-6; cfg->parameter_tolerance = 1e-8; cfg->gradient_tolerance = 1e-10;
}

apex_status apex_nccl_unique_id(void* out128) {
  NcclApi& api = nccl_api();
  if (!api.ok()) return APEX_ERR_NCCL;
  NcclUniqueId id;
  if (api.GetUniqueId(&id) != 0) return APEX_ERR_NCCL;
  std::memcpy(out128, id.internal, 128);
  return APEX_OK;
}

apex_status apex_ctx_create(const apex_ctx_desc* desc, apex_ctx** out) {
  if (!desc || !out) return APEX_ERR_INVALID_INPUT;
  *out = nullptr;
  if (apex_device_count() <= 0) return APEX_ERR_NO_DEVICE;
  if (desc->nranks < 1 || desc->rank < 0 || desc->rank >= desc->nranks) return APEX_ERR_INVALID_INPUT;
  if (cudaSetDevice(desc->device) != cudaSuccess) { cudaGetLastError(); return APEX_ERR_CUDA; }
  apex_ctx* h = new (std::nothrow) apex_ctx();
  if (!h) return APEX_ERR_CUDA;
  Ctx& c = h->c;
  c.device = desc->device; c.rank = desc->rank; c.nranks = desc->nranks;
  bool ok = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && c.state.alloc(1) == cudaSuccess;
  ok = ok && cudaMemset(c.state.p, 0, sizeof(DevState)) == cudaSuccess;
  ok = ok && cudaHostAlloc((void**)&c.h_state, sizeof(DevState), cudaHostAllocDefault) == cudaSuccess;
  cudaDeviceProp prop;
  if (ok && cudaGetDeviceProperties(&prop, c.device) == cudaSuccess) c.num_sms = prop.multiProcessorCount;
  if (!ok) { cudaGetLastError(); delete h; return APEX_ERR_CUDA; }
  std::memset(c.h_state, 0, sizeof(DevState));
  for (int k = 0; k < 2; ++k) {   // upload bounce ring; without it the upload falls back to plain pageable copies
    if (cudaHostAlloc(&c.bounce[k], Ctx::BOUNCE_BYTES, cudaHostAllocDefault) != cudaSuccess || cudaEventCreateWithFlags(&c.bounce_ev[k], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      for (int j = 0; j <= k; ++j) { if (c.bounce[j]) cudaFreeHost(c.bounce[j]); c.bounce[j] = nullptr; }
      break;
    }
  }
  if (c.nranks > 1) {
    NcclApi& api = nccl_api();
    if (!api.ok() || !desc->nccl_unique_id) { delete h; return APEX_ERR_NCCL; }
    NcclUniqueId id;
    std::memcpy(id.internal, desc->nccl_unique_id, 128);
    NcclComm comm = nullptr;
    if (api.CommInitRank(&comm, c.nranks, id, c.rank) != 0) { delete h; return APEX_ERR_NCCL; }
    c.nccl_comm = comm;
    // Per-process communication setup belongs here, not into the first problem upload: the first collective on a new
    // communicator establishes its channels and the first cudaIpcOpenMemHandle enables peer access (together ~300 ms measured at
    // N=2), so the peer-memory buffers of the fused all-reduce are mapped now at a fixed capacity (AR_CAPACITY camera dofs; a
    // problem with more re-maps them at upload).
    if (setup_peer_allreduce(c, AR_CAPACITY) != APEX_OK) { cudaGetLastError(); c.p2p_ok = false; }
  }
  c.self = h;
  *out = h;
  return APEX_OK;
}

void apex_ctx_destroy(apex_ctx* ctx) {
  if (!ctx) return;
  Ctx& c = ctx->c;
  cudaSetDevice(c.device);
  if (c.stream) cudaStreamSynchronize(c.stream);
  if (c.pcg_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)c.pcg_graph_exec);
  release_peer_allreduce(c);
  c.arbuf.release(); c.arflags.release(); c.d_peer_buf.release(); c.d_peer_flags.release();
  if (c.nccl_comm) nccl_api().CommDestroy((NcclComm)c.nccl_comm);
  DevBuf<double>* dbl[] = {&c.slot_uv, &c.cm_uv, &c.pose, &c.intr, &c.pt, &c.pt_full, &c.J, &c.R, &c.hpp, &c.gp, &c.hinv, &c.hcc, &c.partial,
                           &c.sj, &c.pinv, &c.vb, &c.vx, &c.vr, &c.vz, &c.vp, &c.vy, &c.step_cam, &c.step_pt, &c.red_scratch, &c.S, &c.E,
                           &c.dvec, &c.l2flush};
  for (auto* b : dbl) b->release();
  c.giant_tiles.release(); c.xpad.release(); c.chunk_desc.release(); c.cslot_meta.release(); c.cpt_meta.release();
  c.cslot_widx.release(); c.win_desc.release(); c.range_win0.release(); c.win_cams.release(); c.cam_row_start.release(); c.win_dst.release(); c.pair_blocks.release(); c.pair_slots.release(); c.slot_lpg.release(); c.slot_cs8.release(); c.sh_vec.release(); c.tail_trace.release(); c.scale_cam.release(); c.scale_pt.release(); c.slot_loss.release(); c.cm_loss.release(); c.loss_tab.release(); c.det_partial.release();
  c.tiles.release(); c.slot_cam.release(); c.slot_lp.release(); c.pt_slot0.release(); c.pt_cnt.release(); c.items.release();
  c.cam_item_start.release(); c.cm_lp.release(); c.pose_fixed.release(); c.pt_fixed.release(); c.intr_fixed.release();
  c.state.release(); c.trace.release();
  for (cudaEvent_t e : c.ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : c.ev_lin) cudaEventDestroy(e);
  for (cudaEvent_t e : c.ev_form) cudaEventDestroy(e);
  for (cudaEvent_t e : c.ev_chol) cudaEventDestroy(e);
  if (c.ev_lm0) { cudaEventDestroy(c.ev_lm0); cudaEventDestroy(c.ev_lm1); }
  if (c.h_state) cudaFreeHost(c.h_state);
  if (c.h_batch_done) cudaFreeHost(c.h_batch_done);
  for (cudaEvent_t e : c.ev_batch) if (e) cudaEventDestroy(e);
  for (int k = 0; k < 2; ++k) { if (c.bounce[k]) cudaFreeHost(c.bounce[k]); if (c.bounce_ev[k]) cudaEventDestroy(c.bounce_ev[k]); }
  for (cudaEvent_t e : c.chol_events) cudaEventDestroy(e);
  if (c.stream2) cudaStreamDestroy(c.stream2);
  if (c.stream) cudaStreamDestroy(c.stream);
  delete ctx;
}

const char* apex_last_error(const apex_ctx* ctx) { return ctx ? ctx->c.err.c_str() : "null context"; }

apex_status apex_problem_upload(apex_ctx* ctx, const apex_problem_desc* desc) {
  CTX_OR_FAIL(ctx);
  if (!desc) { c.err = "null problem"; return APEX_ERR_INVALID_INPUT; }
  return problem_upload(c, desc);
}

apex_status apex_get_dims(const apex_ctx* ctx, apex_dims* out) {
  if (!ctx || !out) return APEX_ERR_INVALID_INPUT;
  const Ctx& c = ctx->c;
  out->ncam = c.ncam; out->npts = c.npts; out->nobs = c.nobs; out->intr_dim = c.K; out->dc = c.dc;
  out->cam_dof = c.cam_dof_ref; out->lm_dof = 3 * (uint64_t)c.npts; out->npts_local = c.npl; out->flags = c.p2p_ok ? 1u : 0u; out->nobs_local = c.nobs_local;
  return APEX_OK;
}

apex_status apex_params_upload(apex_ctx* ctx, const double* pose, const double* intr, const double* pt) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  cudaStream_t s = c.stream;
  if (pose) APEX_CUDA_TRY(c, cudaMemcpyAsync(c.pose.p, pose, 7 * (size_t)c.ncam * sizeof(double), cudaMemcpyHostToDevice, s));
  std::vector<double> intr_rep;
  if (intr && c.shared_intr) {  // one variable: row 0 is the value of every camera's copy
    intr_rep.resize((size_t)c.K * c.ncam);
    for (uint32_t cam = 0; cam < c.ncam; ++cam) std::copy(intr, intr + c.K, intr_rep.begin() + (size_t)cam * c.K);
    intr = intr_rep.data();
  }
  if (intr) APEX_CUDA_TRY(c, cudaMemcpyAsync(c.intr.p, intr, (size_t)c.K * c.ncam * sizeof(double), cudaMemcpyHostToDevice, s));
  std::vector<double> pt_local;
  if (pt && c.npl) {
    pt_local = gather_local_points(c, pt);
    APEX_CUDA_TRY(c, cudaMemcpyAsync(c.pt.p, pt_local.data(), 3 * (size_t)c.npl * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  c.linearized = false;
  return APEX_OK;
}

// local landmark rows -> their global positions (block-cyclic ownership)
__global__ void scatter_rows_kernel(const double* __restrict__ local, double* __restrict__ full, uint32_t npl, int w, ShardMap m) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)w * npl) return;
  const uint32_t lp = (uint32_t)(i / w);
  full[(size_t)m.to_global(lp) * w + i % w] = local[i];
}

// landmark-indexed device array of local rows [npl][w] -> host array of all rows [npts][w]
static apex_status gather_point_rows(Ctx& c, const double* dev_local, int w, double* host_full) {
  cudaStream_t s = c.stream;
  if (c.nranks == 1) {
    APEX_CUDA_TRY(c, cudaMemcpyAsync(host_full, dev_local, (size_t)w * c.npts * sizeof(double), cudaMemcpyDeviceToHost, s));
  } else {
    const size_t total = (size_t)w * c.npts;
    APEX_CUDA_TRY(c, c.pt_full.alloc(total));
    APEX_CUDA_TRY(c, cudaMemsetAsync(c.pt_full.p, 0, total * sizeof(double), s));
    if (c.npl) {
      scatter_rows_kernel<<<(unsigned)(((size_t)w * c.npl + 255) / 256), 256, 0, s>>>(dev_local, c.pt_full.p, c.npl, w, c.shard);
      c.launches++;
    }
    APEX_TRY(allreduce_sum(c, c.pt_full.p, total));
    APEX_CUDA_TRY(c, cudaMemcpyAsync(host_full, c.pt_full.p, total * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  return APEX_OK;
}

apex_status apex_params_download(apex_ctx* ctx, double* pose, double* intr, double* pt) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  cudaStream_t s = c.stream;
  if (pose) APEX_CUDA_TRY(c, cudaMemcpyAsync(pose, c.pose.p, 7 * (size_t)c.ncam * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (intr) APEX_CUDA_TRY(c, cudaMemcpyAsync(intr, c.intr.p, (size_t)c.K * c.ncam * sizeof(double), cudaMemcpyDeviceToHost, s));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  if (pt) APEX_TRY(gather_point_rows(c, c.pt.p, 3, pt));
  return APEX_OK;
}

apex_status apex_linearize(apex_ctx* ctx, double lambda) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  APEX_TRY(do_linearize(c, lambda));
  APEX_TRY(sync_state(c));
  return APEX_OK;
}

apex_status apex_cost(apex_ctx* ctx, double* cost) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  APEX_TRY(launch_cost(c, nullptr));
  APEX_TRY(sync_state(c));
  const double cn = std::sqrt(c.h_state->cost2_local);
  if (cost) *cost = 0.5 * cn * cn;
  return APEX_OK;
}

apex_status apex_get_linearization(apex_ctx* ctx, double* r, double* jc, double* jp) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  if (!c.linearized) { c.err = "not linearized"; return APEX_ERR_INVALID_STATE; }
  if (c.nranks != 1) { c.err = "get_linearization is single-rank only"; return APEX_ERR_UNSUPPORTED; }
  const int dc = c.dc, np = c.np;
  std::vector<double> hJ((size_t)c.nchunks * np * TILE), hR((size_t)c.nchunks * 2 * TILE);
  APEX_CUDA_TRY(c, cudaMemcpyAsync(hJ.data(), c.J.p, hJ.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(hR.data(), c.R.p, hR.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(c.stream));
  for (size_t slot = 0; slot < c.nslots; ++slot) {
    const uint64_t o = c.slot_obs[slot];
    if (o == UINT64_MAX) continue;
    const size_t ch = slot / TILE, lane = slot % TILE;
    if (r) { r[2 * o] = hR[(ch * 2 + 0) * TILE + lane]; r[2 * o + 1] = hR[(ch * 2 + 1) * TILE + lane]; }
    if (jc) for (int k = 0; k < 2 * dc; ++k) jc[o * 2 * dc + k] = hJ[jplane_index(ch, np, k, c.slot_pos[slot])];  // camera half: camera-sorted lane
    if (jp) for (int k = 0; k < 6; ++k) jp[o * 6 + k] = hJ[jplane_index(ch, np, 2 * dc + k, lane)];
  }
  return APEX_OK;
}

apex_status apex_get_blocks(apex_ctx* ctx, double* hcc, double* gc, double* hpp, double* gp, double* hpp_inv) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  if (!c.linearized) { c.err = "not linearized"; return APEX_ERR_INVALID_STATE; }
  if (c.nranks != 1) { c.err = "get_blocks is single-rank only"; return APEX_ERR_UNSUPPORTED; }
  cudaStream_t s = c.stream;
  const size_t ncd = (size_t)c.ncam * c.dc, n = c.npl;
  if (hcc) APEX_CUDA_TRY(c, cudaMemcpyAsync(hcc, c.hcc.p, ncd * c.dc * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (gc) APEX_CUDA_TRY(c, cudaMemcpyAsync(gc, c.gc, ncd * sizeof(double), cudaMemcpyDeviceToHost, s));
  std::vector<double> h6(6 * n), g3(3 * n), i6(6 * n);
  APEX_CUDA_TRY(c, cudaMemcpyAsync(h6.data(), c.hpp.p, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(g3.data(), c.gp.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(i6.data(), c.hinv.p, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  static const int sym[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
  for (size_t p = 0; p < n; ++p) {
    if (hpp) for (int k = 0; k < 9; ++k) hpp[9 * p + k] = h6[sym[k] * n + p];
    if (hpp_inv) for (int k = 0; k < 9; ++k) hpp_inv[9 * p + k] = i6[sym[k] * n + p];
    if (gp) for (int k = 0; k < 3; ++k) gp[3 * p + k] = g3[k * n + p];
  }
  return APEX_OK;
}

apex_status apex_schur_matvec(apex_ctx* ctx, const double* x, double* y) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  if (!c.linearized) { c.err = "not linearized"; return APEX_ERR_INVALID_STATE; }
  if (!x || !y) { c.err = "null vector"; return APEX_ERR_INVALID_INPUT; }
  const size_t n = (size_t)c.ncam * c.dc;
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.vp.p, x, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  APEX_TRY(schur_operator(c, c.vp.p, c.vy.p, 0));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(y, c.vy.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(c.stream));
  return APEX_OK;
}

apex_status apex_schur_matvec_bench(apex_ctx* ctx, int32_t reps, int32_t flush_l2, double* ms_per_call) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  if (!c.linearized) { c.err = "not linearized"; return APEX_ERR_INVALID_STATE; }
  if (reps < 1) reps = 1;
  const size_t n = (size_t)c.ncam * c.dc;
  const size_t flush_n = (size_t)256 << 20 >> 3;  // 256 MiB > 126 MB L2
  if (flush_l2) APEX_CUDA_TRY(c, c.l2flush.alloc(flush_n));
  std::vector<double> ones(n, 1.0);
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.vp.p, ones.data(), n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  cudaEvent_t e0, e1;
  APEX_CUDA_TRY(c, cudaEventCreate(&e0));
  APEX_CUDA_TRY(c, cudaEventCreate(&e1));
  double total_ms = 0.0;
  for (int i = -3; i < reps; ++i) {  // 3 warm-up applications
    if (flush_l2) APEX_CUDA_TRY(c, cudaMemsetAsync(c.l2flush.p, i & 1, flush_n * sizeof(double), c.stream));
    APEX_CUDA_TRY(c, cudaEventRecord(e0, c.stream));
    APEX_TRY(schur_operator_local(c, c.vp.p, c.vy.p, 0));
    APEX_CUDA_TRY(c, cudaEventRecord(e1, c.stream));
    APEX_CUDA_TRY(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    APEX_CUDA_TRY(c, cudaEventElapsedTime(&ms, e0, e1));
    if (i >= 0) total_ms += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_per_call) *ms_per_call = total_ms / reps;
  return APEX_OK;
}

apex_status apex_dense_cholesky_bench(apex_ctx* ctx, uint32_t n, int32_t reps, double* ms_per_factorization) {
  CTX_OR_FAIL(ctx);
  if (n == 0 || n > 60000) { c.err = "n out of range"; return APEX_ERR_INVALID_INPUT; }
  return dense_cholesky_bench(c, n, reps, ms_per_factorization);
}

apex_status apex_solve_augmented(apex_ctx* ctx, int32_t schur_variant, int32_t preconditioner, int32_t cg_max_iterations, double cg_tolerance,
                                 double lambda, double* step_cam, double* step_pt, double* grad_norm, int32_t* pcg_iters) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  if (schur_variant < APEX_SCHUR_EXPLICIT || schur_variant > APEX_SCHUR_EXPLICIT_PCG) { c.err = "bad schur_variant"; return APEX_ERR_INVALID_PARAMETERS; }
  if (!c.linearized || c.lin_lambda != lambda) APEX_TRY(do_linearize(c, lambda));
  apex_status st;
  if (schur_variant == APEX_SCHUR_IMPLICIT) st = solve_implicit(c, preconditioner, cg_max_iterations, cg_tolerance);
  else st = solve_explicit(c, schur_variant == APEX_SCHUR_EXPLICIT_PCG, cg_max_iterations, cg_tolerance);
  if (st != APEX_OK) return st;
  c.have_step = true;
  APEX_TRY(launch_step_norms(c));
  APEX_TRY(sync_state(c));
  if (c.h_state->singular_landmark) { c.err = "Landmark block singular"; return APEX_ERR_SINGULAR_MATRIX; }
  if (grad_norm) *grad_norm = std::sqrt(c.h_state->g2_cam + c.h_state->g2_pt);
  if (pcg_iters) *pcg_iters = (int32_t)c.last_pcg_iters;
  if (step_cam) {
    APEX_CUDA_TRY(c, cudaMemcpyAsync(step_cam, c.step_cam.p, (size_t)c.ncam * c.dc * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    APEX_CUDA_TRY(c, cudaStreamSynchronize(c.stream));
  }
  if (step_pt) APEX_TRY(gather_point_rows(c, c.step_pt.p, 3, step_pt));
  return APEX_OK;
}

apex_status apex_get_step(apex_ctx* ctx, double* step_cam, double* step_pt) {
  CTX_OR_FAIL(ctx);
  NEED_PROBLEM();
  if (!c.have_step) { c.err = "no step computed yet"; return APEX_ERR_INVALID_STATE; }
  if (step_cam) {
    APEX_CUDA_TRY(c, cudaMemcpyAsync(step_cam, c.step_cam.p, (size_t)c.ncam * c.dc * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    APEX_CUDA_TRY(c, cudaStreamSynchronize(c.stream));
  }
  if (step_pt) APEX_TRY(gather_point_rows(c, c.step_pt.p, 3, step_pt));
  return APEX_OK;
}

apex_status apex_lm_solve(apex_ctx* ctx, const apex_lm_config* cfg, apex_lm_result* result, apex_iter_trace* trace, int32_t trace_cap) {
  CTX_OR_FAIL(ctx);
  if (!cfg || !result) { c.err = "null config/result"; return APEX_ERR_INVALID_INPUT; }
  return lm_solve(c, cfg, result, trace, trace_cap);
}

apex_status apex_add_observer(apex_ctx* ctx, const apex_observer* observer) {
  CTX_OR_FAIL(ctx);
  if (!observer) { c.err = "null observer"; return APEX_ERR_INVALID_INPUT; }
  c.observers.push_back(*observer);
  return APEX_OK;
}
apex_status apex_clear_observers(apex_ctx* ctx) {
  CTX_OR_FAIL(ctx);
  c.observers.clear();
  return APEX_OK;
}

int64_t apex_kernel_launches(const apex_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

apex_status apex_profile_enable(apex_ctx* ctx, int32_t on) {
  CTX_OR_FAIL(ctx);
  c.prof = on != 0;
  c.ev_mv_used = 0;
  c.ev_lin_used = 0;
  c.ev_form_used = c.ev_chol_used = 0;
  return APEX_OK;
}

apex_status apex_profile_read(apex_ctx* ctx, apex_profile* out) {
  CTX_OR_FAIL(ctx);
  if (!out) { c.err = "null profile"; return APEX_ERR_INVALID_INPUT; }
  APEX_CUDA_TRY(c, cudaStreamSynchronize(c.stream));
  std::memset(out, 0, sizeof(*out));
  auto total = [&](std::vector<cudaEvent_t>& pool, size_t used, double& ms, int64_t& n) {
    for (size_t i = 0; i < used; ++i) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, pool[2 * i], pool[2 * i + 1]) == cudaSuccess) { ms += t; ++n; } else cudaGetLastError();
    }
  };
  total(c.ev_pool, c.ev_mv_used, out->matvec_ms, out->matvec_launches);
  total(c.ev_lin, c.ev_lin_used, out->linearize_ms, out->linearize_launches);
  total(c.ev_form, c.ev_form_used, out->schur_form_ms, out->schur_forms);
  total(c.ev_chol, c.ev_chol_used, out->cholesky_ms, out->cholesky_factorizations);
  out->cholesky_n = c.chol_n;
  out->upload_h2d_bytes = c.upload_h2d_bytes;
  c.ev_mv_used = 0;
  c.ev_lin_used = 0;
  c.ev_form_used = c.ev_chol_used = 0;
  if (c.lm_timed) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, c.ev_lm0, c.ev_lm1) == cudaSuccess) out->lm_device_ms = t; else cudaGetLastError();
  }
  return APEX_OK;
}

apex_status apex_layout_stats_compute(const apex_problem_desc* desc, int32_t nranks, int32_t rank, apex_layout_stats* out) {
  if (!desc || !out || nranks < 1 || rank < 0 || rank >= nranks) return APEX_ERR_INVALID_INPUT;
  std::memset(out, 0, sizeof(*out));
  std::string err;
  return layout_stats(desc, nranks, rank, out, err);
}

apex_status apex_shard_info(uint32_t npts, uint64_t nobs, const uint32_t* obs_pt, int32_t nranks, int32_t rank, uint32_t* block,
                            uint32_t* npts_local, uint64_t* nobs_local) {
  if (!obs_pt && nobs) return APEX_ERR_INVALID_INPUT;
  if (nranks < 1 || rank < 0 || rank >= nranks) return APEX_ERR_INVALID_INPUT;
  const ShardMap m{npts, (uint32_t)nranks, (uint32_t)rank};
  uint64_t mine = 0;
  for (uint64_t o = 0; o < nobs; ++o) {
    if (obs_pt[o] >= npts) return APEX_ERR_INVALID_INPUT;
    mine += m.owns(obs_pt[o]) ? 1 : 0;
  }
  if (block) *block = SHARD_BLOCK;
  if (npts_local) *npts_local = m.count();
  if (nobs_local) *nobs_local = mine;
  return APEX_OK;
}

}  // extern "C"
