// problem.cu — apex_problem_upload: replaces Problem construction + initialize_optimization_state
// (src/core/problem.rs:518-808, src/optimizer/mod.rs:522-563) for the SoA factor graph of
// bin/bundle_adjustment.rs:212-441. Builds the static observation structure described in apex_ctx.h:
// landmark sharding across ranks, point-major tiles, camera-major work items.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "apex_ctx.h"
#include "ba_device.cuh"

namespace apex {

template <typename T>
static cudaError_t upload_vec(DevBuf<T>& buf, const std::vector<T>& v, cudaStream_t s) {
  cudaError_t e = buf.alloc(v.size());
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  return cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

// Contiguous landmark ranges balanced by observation count; identical on every rank. pt_start receives the
// exclusive prefix sum of the per-landmark observation counts.
void shard_range(uint32_t npts, uint64_t nobs, const uint32_t* obs_pt, int nranks, int rank, std::vector<uint64_t>& pt_start, uint32_t& p0,
                 uint32_t& p1) {
  pt_start.assign((size_t)npts + 1, 0);
  for (uint64_t o = 0; o < nobs; ++o) pt_start[obs_pt[o] + 1]++;
  for (uint32_t p = 0; p < npts; ++p) pt_start[p + 1] += pt_start[p];
  auto boundary = [&](int r) -> uint32_t {
    if (r <= 0) return 0;
    if (r >= nranks) return npts;
    if (nobs == 0) return (uint32_t)((uint64_t)npts * r / nranks);
    uint64_t target = nobs * (uint64_t)r / (uint64_t)nranks;
    return (uint32_t)(std::lower_bound(pt_start.begin(), pt_start.begin() + npts, target) - pt_start.begin());
  };
  p0 = boundary(rank);
  p1 = boundary(rank + 1);
}

apex_status problem_upload(Ctx& c, const apex_problem_desc* d) {
  // ---- validation (same failures as the reference / oracle) ----
  int K = model_intr_dim(d->camera_model);
  if (K < 0) { c.err = "camera model not supported on the GPU path"; return APEX_ERR_UNSUPPORTED; }
  if (d->intr_dim != K) { c.err = "intr_dim does not match camera model"; return APEX_ERR_INVALID_INPUT; }
  if ((d->opt_flags & (APEX_OPT_POSE | APEX_OPT_LANDMARK)) != (APEX_OPT_POSE | APEX_OPT_LANDMARK)) {
    c.err = "only BundleAdjustment / SelfCalibration OptimizeParams are live (bin/bundle_adjustment.rs)";
    return APEX_ERR_UNSUPPORTED;
  }
  if (d->ncam == 0) { c.err = "No camera variables found"; return APEX_ERR_INVALID_INPUT; }    // explicit_schur.rs:278-282
  if (d->npts == 0) { c.err = "No landmark variables found"; return APEX_ERR_INVALID_INPUT; }  // explicit_schur.rs:283-287
  if (d->nobs > 0xFFFFFFF0ull) { c.err = "too many observations for u32 slots"; return APEX_ERR_UNSUPPORTED; }
  if (d->loss_id < APEX_LOSS_NONE || d->loss_id > APEX_LOSS_T_DISTRIBUTION) { c.err = "unknown loss id"; return APEX_ERR_INVALID_INPUT; }
  const uint64_t nobs = d->nobs;
  for (uint64_t o = 0; o < nobs; ++o)
    if (d->obs_cam[o] >= d->ncam || d->obs_pt[o] >= d->npts) { c.err = "observation index out of range"; return APEX_ERR_INVALID_INPUT; }

  c.have_problem = false;
  c.linearized = false;
  c.model = d->camera_model; c.K = K; c.opt = d->opt_flags;
  c.opt_intr = (d->opt_flags & APEX_OPT_INTRINSIC) != 0;
  c.intr_vars = d->intr_vars_present != 0;
  c.dc = 6 + (c.opt_intr ? K : 0);
  c.np = 2 * (c.dc + 3);
  c.ncam = d->ncam; c.npts = d->npts; c.nobs = nobs;
  c.cam_dof_ref = (uint64_t)c.ncam * (6 + ((c.opt_intr || c.intr_vars) ? K : 0));
  c.loss_id = d->loss_id;
  for (int i = 0; i < 4; ++i) c.loss_p[i] = d->loss_params[i];

  // ---- landmark sharding: contiguous ranges balanced by observation count ----
  std::vector<uint64_t> pt_start;
  shard_range(c.npts, nobs, d->obs_pt, c.nranks, c.rank, pt_start, c.p0, c.p1);
  c.npl = c.p1 - c.p0;
  c.nobs_local = pt_start[c.p1] - pt_start[c.p0];

  // ---- point-major order of the local observations (stable in the caller's insertion order) ----
  std::vector<uint64_t> pm(c.nobs_local);
  {
    std::vector<uint64_t> cur(pt_start.begin() + c.p0, pt_start.begin() + c.p1);
    const uint64_t base = pt_start[c.p0];
    for (uint64_t o = 0; o < nobs; ++o) {
      uint32_t p = d->obs_pt[o];
      if (p >= c.p0 && p < c.p1) pm[cur[p - c.p0]++ - base] = o;
    }
  }

  // ---- tiles ----
  // Normal tiles (<=256 observations, <=128 landmarks) get chunk ids 0..nn-1 in landmark order so that the
  // persistent operator kernel can pair chunks (2s, 2s+1); landmarks with more than 256 observations get
  // their chunks after those.
  std::vector<TileDesc> tiles;
  std::vector<uint32_t> pt_slot0(c.npl), pt_cnt(c.npl);
  uint32_t chunk = 0;
  {
    uint32_t cur_pt0 = 0, cur_npt = 0, cur_obs = 0;
    auto flush = [&]() {
      if (cur_npt == 0) return;
      tiles.push_back({cur_pt0, cur_npt, chunk, 1});
      chunk += 1;
      cur_npt = 0; cur_obs = 0;
    };
    for (uint32_t lp = 0; lp < c.npl; ++lp) {
      uint64_t k64 = pt_start[c.p0 + lp + 1] - pt_start[c.p0 + lp];
      uint32_t k = (uint32_t)k64;
      pt_cnt[lp] = k;
      if (k > (uint32_t)TILE) {
        flush();
        tiles.push_back({lp, 1, 0xFFFFFFFFu, (k + TILE - 1) / TILE});  // chunk0 assigned below
        continue;
      }
      if (cur_npt > 0 && (cur_obs + k > (uint32_t)TILE || cur_npt >= (uint32_t)MAX_TILE_PTS)) flush();
      if (cur_npt == 0) cur_pt0 = lp;
      pt_slot0[lp] = chunk * TILE + cur_obs;
      cur_npt++;
      cur_obs += k;
    }
    flush();
  }
  c.nnormal_chunks = chunk;
  std::vector<TileDesc> giant_tiles;
  for (TileDesc& t : tiles)
    if (t.nchunks > 1) {
      t.chunk0 = chunk;
      pt_slot0[t.pt0] = chunk * TILE;
      chunk += t.nchunks;
      giant_tiles.push_back(t);
    }
  c.ngiant = (uint32_t)giant_tiles.size();
  c.ntiles = (uint32_t)tiles.size();
  c.nchunks = chunk;
  c.nslots = (size_t)chunk * TILE;
  c.h_pt_cnt = pt_cnt;

  // ---- slot arrays ----
  std::vector<uint32_t> slot_cam(c.nslots, PAD_CAM);
  std::vector<uint16_t> slot_lp(c.nslots, 0);
  std::vector<double> slot_uv(c.nslots * 2, 0.0);
  c.slot_obs.assign(c.nslots, UINT64_MAX);
  {
    uint64_t q = 0;
    for (const TileDesc& t : tiles) {
      for (uint32_t i = 0; i < t.npt; ++i) {
        uint32_t lp = t.pt0 + i;
        for (uint32_t k = 0; k < pt_cnt[lp]; ++k, ++q) {
          size_t slot = (size_t)pt_slot0[lp] + k;
          uint64_t o = pm[q];
          slot_cam[slot] = d->obs_cam[o];
          slot_lp[slot] = (uint16_t)i;
          size_t ch = slot / TILE, lane = slot % TILE;
          slot_uv[(ch * 2 + 0) * TILE + lane] = d->obs_uv[2 * o];
          slot_uv[(ch * 2 + 1) * TILE + lane] = d->obs_uv[2 * o + 1];
          c.slot_obs[slot] = o;
        }
      }
    }
  }

  // ---- supertiles: camera-sorted segment structure over chunk pairs (2s, 2s+1) ----
  c.nsuper = (c.nnormal_chunks + 1) / 2;
  std::vector<SuperDesc> supers(c.nsuper);
  std::vector<uint2> slot_meta(c.nslots, make_uint2(PAD_CAM, 0));
  std::vector<uint32_t> pt_meta(c.npl, 0);
  std::vector<uint32_t> seg_cam((size_t)c.nsuper * STILE, 0);
  std::vector<uint16_t> seg_begin((size_t)c.nsuper * (STILE + 2), 0);
  {
    std::vector<const TileDesc*> by_chunk(c.nnormal_chunks, nullptr);
    for (const TileDesc& t : tiles) if (t.nchunks == 1) by_chunk[t.chunk0] = &t;
    std::vector<std::pair<uint32_t, uint32_t>> order;  // (camera, supertile-local slot)
    for (uint32_t st = 0; st < c.nsuper; ++st) {
      SuperDesc& d = supers[st];
      const TileDesc* ta = by_chunk[2 * st];
      const TileDesc* tb = 2 * st + 1 < c.nnormal_chunks ? by_chunk[2 * st + 1] : nullptr;
      d = SuperDesc{ta->pt0, ta->npt, tb ? tb->pt0 : 0u, tb ? tb->npt : 0u, 0u, tb ? 1u : 0u, {0u, 0u}};
      order.clear();
      for (int h = 0; h < 2; ++h) {
        const TileDesc* t = h == 0 ? ta : tb;
        if (!t) continue;
        const uint32_t spt0 = h == 0 ? 0 : ta->npt;
        for (uint32_t i = 0; i < t->npt; ++i) {
          const uint32_t lp = t->pt0 + i;
          const uint32_t off = h * TILE + (pt_slot0[lp] - t->chunk0 * TILE);
          pt_meta[lp] = off | (pt_cnt[lp] << 16);
          for (uint32_t k = 0; k < pt_cnt[lp]; ++k) {
            const size_t slot = (size_t)pt_slot0[lp] + k;
            slot_meta[slot].x = slot_cam[slot];
            slot_meta[slot].y = spt0 + i;  // position filled below
            order.push_back({slot_cam[slot], off + k});
          }
        }
      }
      std::sort(order.begin(), order.end());
      uint32_t nseg = 0;
      for (size_t pos = 0; pos < order.size(); ++pos) {
        if (pos == 0 || order[pos].first != order[pos - 1].first) {
          seg_cam[(size_t)st * STILE + nseg] = order[pos].first;
          seg_begin[(size_t)st * (STILE + 2) + nseg] = (uint16_t)pos;
          ++nseg;
        }
        const uint32_t sl = order[pos].second;
        const size_t slot = (size_t)(2 * st + sl / TILE) * TILE + sl % TILE;
        slot_meta[slot].y |= (uint32_t)pos << 16;
      }
      seg_begin[(size_t)st * (STILE + 2) + nseg] = (uint16_t)order.size();
      d.nseg = nseg;
    }
  }

  // ---- per-chunk camera-sorted segment structure (ping-pong operator kernel) ----
  const uint32_t nchunk_even = (c.nnormal_chunks + 1) & ~1u;
  std::vector<ChunkDesc> chunk_desc(nchunk_even, ChunkDesc{0, 0, 0, 0});
  std::vector<uint2> cslot_meta((size_t)nchunk_even * TILE, make_uint2(PAD_CAM, 0));
  std::vector<uint32_t> cpt_meta(c.npl, 0);
  std::vector<uint32_t> cseg_cam((size_t)nchunk_even * TILE, 0);
  std::vector<uint16_t> cseg_begin((size_t)nchunk_even * CSEG_LD, 0);
  {
    std::vector<std::pair<uint32_t, uint32_t>> order;
    for (const TileDesc& t : tiles) {
      if (t.nchunks != 1) continue;
      const uint32_t ch = t.chunk0;
      order.clear();
      for (uint32_t i = 0; i < t.npt; ++i) {
        const uint32_t lp = t.pt0 + i;
        const uint32_t off = pt_slot0[lp] - ch * TILE;
        cpt_meta[lp] = off | (pt_cnt[lp] << 16);
        for (uint32_t k = 0; k < pt_cnt[lp]; ++k) {
          const size_t slot = (size_t)pt_slot0[lp] + k;
          cslot_meta[slot].x = slot_cam[slot];
          cslot_meta[slot].y = i;
          order.push_back({slot_cam[slot], off + k});
        }
      }
      std::sort(order.begin(), order.end());
      uint32_t nseg = 0;
      for (size_t pos = 0; pos < order.size(); ++pos) {
        if (pos == 0 || order[pos].first != order[pos - 1].first) {
          cseg_cam[(size_t)ch * TILE + nseg] = order[pos].first;
          cseg_begin[(size_t)ch * CSEG_LD + nseg] = (uint16_t)pos;
          ++nseg;
        }
        cslot_meta[(size_t)ch * TILE + order[pos].second].y |= ((uint32_t)pos << 8) | ((nseg - 1) << 16);
      }
      cseg_begin[(size_t)ch * CSEG_LD + nseg] = (uint16_t)order.size();
      chunk_desc[ch] = ChunkDesc{t.pt0, t.npt, nseg, 0};
    }
  }

  // ---- camera-major copy of the local observations + work items ----
  std::vector<uint32_t> cam_start((size_t)c.ncam + 1, 0);
  for (uint64_t q = 0; q < c.nobs_local; ++q) cam_start[d->obs_cam[pm[q]] + 1]++;
  for (uint32_t k = 0; k < c.ncam; ++k) cam_start[k + 1] += cam_start[k];
  std::vector<double> cm_uv(2 * (size_t)c.nobs_local);
  std::vector<uint32_t> cm_lp(c.nobs_local);
  {
    std::vector<uint32_t> cur(cam_start.begin(), cam_start.end() - 1);
    uint64_t q = 0;
    for (uint32_t lp = 0; lp < c.npl; ++lp)
      for (uint32_t k = 0; k < pt_cnt[lp]; ++k, ++q) {
        uint64_t o = pm[q];
        uint32_t pos = cur[d->obs_cam[o]]++;
        cm_uv[pos] = d->obs_uv[2 * o];
        cm_uv[(size_t)c.nobs_local + pos] = d->obs_uv[2 * o + 1];
        cm_lp[pos] = lp;
      }
  }
  std::vector<CamItem> items;
  std::vector<uint32_t> cam_item_start((size_t)c.ncam + 1, 0);
  for (uint32_t k = 0; k < c.ncam; ++k) {
    cam_item_start[k] = (uint32_t)items.size();
    for (uint32_t b = cam_start[k]; b < cam_start[k + 1]; b += CAM_CHUNK)
      items.push_back({k, b, std::min<uint32_t>(b + CAM_CHUNK, cam_start[k + 1]), 0});
  }
  cam_item_start[c.ncam] = (uint32_t)items.size();
  c.nitems = (uint32_t)items.size();

  // ---- fixed masks ----
  std::vector<uint8_t> pose_fixed(c.ncam, 0), pt_fixed(c.npl, 0);
  std::vector<uint16_t> intr_fixed(c.ncam, 0);
  if (d->pose_fixed) std::copy(d->pose_fixed, d->pose_fixed + c.ncam, pose_fixed.begin());
  if (d->intr_fixed) std::copy(d->intr_fixed, d->intr_fixed + c.ncam, intr_fixed.begin());
  if (d->pt_fixed) std::copy(d->pt_fixed + c.p0, d->pt_fixed + c.p1, pt_fixed.begin());

  // ---- to the device ----
  cudaStream_t s = c.stream;
  APEX_CUDA_TRY(c, upload_vec(c.tiles, tiles, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_cam, slot_cam, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_lp, slot_lp, s));
  APEX_CUDA_TRY(c, upload_vec(c.giant_tiles, giant_tiles, s));
  APEX_CUDA_TRY(c, upload_vec(c.supers, supers, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_meta, slot_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_meta, pt_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.seg_cam, seg_cam, s));
  APEX_CUDA_TRY(c, upload_vec(c.seg_begin, seg_begin, s));
  APEX_CUDA_TRY(c, upload_vec(c.chunk_desc, chunk_desc, s));
  APEX_CUDA_TRY(c, upload_vec(c.cslot_meta, cslot_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.cpt_meta, cpt_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.cseg_cam, cseg_cam, s));
  APEX_CUDA_TRY(c, upload_vec(c.cseg_begin, cseg_begin, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_uv, slot_uv, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_slot0, pt_slot0, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_cnt, pt_cnt, s));
  APEX_CUDA_TRY(c, upload_vec(c.items, items, s));
  APEX_CUDA_TRY(c, upload_vec(c.cam_item_start, cam_item_start, s));
  APEX_CUDA_TRY(c, upload_vec(c.cm_uv, cm_uv, s));
  APEX_CUDA_TRY(c, upload_vec(c.cm_lp, cm_lp, s));
  APEX_CUDA_TRY(c, upload_vec(c.pose_fixed, pose_fixed, s));
  APEX_CUDA_TRY(c, upload_vec(c.intr_fixed, intr_fixed, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_fixed, pt_fixed, s));

  const size_t ncd = (size_t)c.ncam * c.dc;
  const int pb = 36 + K * K;
  APEX_CUDA_TRY(c, c.pose.alloc((size_t)c.ncam * 7));
  APEX_CUDA_TRY(c, c.intr.alloc((size_t)c.ncam * K));
  APEX_CUDA_TRY(c, c.pt.alloc((size_t)c.npl * 3));
  APEX_CUDA_TRY(c, c.J.alloc((size_t)c.nchunks * c.np * TILE));
  APEX_CUDA_TRY(c, c.R.alloc((size_t)c.nchunks * 2 * TILE));
  APEX_CUDA_TRY(c, c.hpp.alloc((size_t)c.npl * 6));
  APEX_CUDA_TRY(c, c.gp.alloc((size_t)c.npl * 3));
  APEX_CUDA_TRY(c, c.hinv.alloc((size_t)c.npl * 6));
  APEX_CUDA_TRY(c, c.hcc.alloc(ncd * c.dc + ncd));
  c.gc = c.hcc.p + ncd * c.dc;
  const size_t nacc_max = (size_t)c.dc * (c.dc + 1) / 2 + c.dc + 64;
  APEX_CUDA_TRY(c, c.partial.alloc((size_t)std::max<uint32_t>(c.nitems, 1) * nacc_max));
  APEX_CUDA_TRY(c, c.sj.alloc((size_t)c.ncam * pb));
  APEX_CUDA_TRY(c, c.pinv.alloc((size_t)c.ncam * pb));
  APEX_CUDA_TRY(c, c.vb.alloc(ncd));
  APEX_CUDA_TRY(c, c.vx.alloc(ncd));
  APEX_CUDA_TRY(c, c.vr.alloc(ncd));
  APEX_CUDA_TRY(c, c.vz.alloc(ncd));
  APEX_CUDA_TRY(c, c.vp.alloc(ncd));
  APEX_CUDA_TRY(c, c.vy.alloc(ncd));
  APEX_CUDA_TRY(c, c.ypart.alloc((size_t)c.num_sms * ncd));
  APEX_CUDA_TRY(c, c.xpad.alloc((size_t)c.ncam * (c.dc + 2)));
  APEX_CUDA_TRY(c, c.step_cam.alloc(ncd));
  APEX_CUDA_TRY(c, c.step_pt.alloc((size_t)c.npl * 3));
  APEX_CUDA_TRY(c, c.red_scratch.alloc(8 * (size_t)std::max<uint32_t>(std::max<uint32_t>(c.nchunks, c.ncam), 1024u) + 64));
  APEX_CUDA_TRY(c, cudaMemsetAsync(c.step_cam.p, 0, ncd * sizeof(double), s));
  APEX_CUDA_TRY(c, cudaMemsetAsync(c.step_pt.p, 0, std::max<size_t>((size_t)c.npl * 3, 1) * sizeof(double), s));

  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.pose.p, d->pose, (size_t)c.ncam * 7 * sizeof(double), cudaMemcpyHostToDevice, s));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.intr.p, d->intr, (size_t)c.ncam * K * sizeof(double), cudaMemcpyHostToDevice, s));
  if (c.npl)
    APEX_CUDA_TRY(c, cudaMemcpyAsync(c.pt.p, d->pt + 3 * (size_t)c.p0, (size_t)c.npl * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));  // the host vectors above die with this scope
  c.have_problem = true;
  return APEX_OK;
}

}  // namespace apex
