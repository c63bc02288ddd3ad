// problem.cu — apex_problem_upload: replaces Problem construction + initialize_optimization_state
// (src/core/problem.rs:518-808, src/optimizer/mod.rs:522-563) for the SoA factor graph of
// bin/bundle_adjustment.rs:212-441. Builds the static observation structure described in apex_ctx.h:
// landmark sharding across ranks, point-major tiles with per-chunk camera segments, camera-major work items.
// build_layout() is pure host code (OpenMP over tiles) so it can be exercised without a device.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <omp.h>

#include "apex_ctx.h"
#include "ba_device.cuh"

namespace apex {

static thread_local uint64_t g_h2d_bytes = 0;  // bytes enqueued by upload_vec since problem_upload reset it

template <typename T>
static cudaError_t upload_vec(DevBuf<T>& buf, const StageBuf<T>& v, cudaStream_t s) {
  cudaError_t e = buf.alloc(v.size());
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  g_h2d_bytes += v.size() * sizeof(T);
  return cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

template <typename T, typename A>
static cudaError_t upload_vec(DevBuf<T>& buf, const std::vector<T, A>& v, cudaStream_t s) {
  cudaError_t e = buf.alloc(v.size());
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  g_h2d_bytes += v.size() * sizeof(T);
  return cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

apex_status validate_problem(const apex_problem_desc* d, std::string& err) {
  int K = model_intr_dim(d->camera_model);
  if (K < 0) { err = "camera model not supported on the GPU path"; return APEX_ERR_UNSUPPORTED; }
  if (d->intr_dim != K) { err = "intr_dim does not match camera model"; return APEX_ERR_INVALID_INPUT; }
  if ((d->opt_flags & (APEX_OPT_POSE | APEX_OPT_LANDMARK)) != (APEX_OPT_POSE | APEX_OPT_LANDMARK)) {
    err = "only BundleAdjustment / SelfCalibration OptimizeParams are live (bin/bundle_adjustment.rs)";
    return APEX_ERR_UNSUPPORTED;
  }
  if (d->ncam == 0) { err = "No camera variables found"; return APEX_ERR_INVALID_INPUT; }    // explicit_schur.rs:278-282
  if (d->npts == 0) { err = "No landmark variables found"; return APEX_ERR_INVALID_INPUT; }  // explicit_schur.rs:283-287
  if (d->nobs > 0xFFFFFFF0ull) { err = "too many observations for u32 slots"; return APEX_ERR_UNSUPPORTED; }
  if (d->loss_id < APEX_LOSS_NONE || d->loss_id > APEX_LOSS_T_DISTRIBUTION) { err = "unknown loss id"; return APEX_ERR_INVALID_INPUT; }
  const uint64_t nobs = d->nobs;
  int bad = 0;
#pragma omp parallel for reduction(| : bad) schedule(static)
  for (int64_t o = 0; o < (int64_t)nobs; ++o) bad |= (d->obs_cam[o] >= d->ncam || d->obs_pt[o] >= d->npts) ? 1 : 0;
  if (bad) { err = "observation index out of range"; return APEX_ERR_INVALID_INPUT; }
  return APEX_OK;
}

// The static structure of one rank's shard, on the host.
struct HostLayout {
  ShardMap shard;
  uint32_t npl = 0;
  uint64_t nobs_local = 0;
  uint32_t nnormal_chunks = 0, nchunks = 0, npairs = 0;
  std::vector<TileDesc> tiles, giant_tiles;
  std::vector<uint32_t> pt_slot0, pt_cnt;
  StageBuf<uint32_t> slot_cam;
  StageBuf<uint16_t> slot_lp;
  StageBuf<double> slot_uv;
  HostVec<uint64_t> slot_obs;
  std::vector<ChunkDesc> chunk_desc;
  StageBuf<uint2> cslot_meta;
  std::vector<uint32_t> cpt_meta;
  StageBuf<uint32_t> cseg_cam;
  StageBuf<uint16_t> cseg_begin;
  StageBuf<double> cm_uv;
  StageBuf<uint32_t> cm_lp;
  void set_pinned(bool on) {
    slot_cam.pinned = slot_lp.pinned = slot_uv.pinned = cslot_meta.pinned = cseg_cam.pinned = cseg_begin.pinned = cm_uv.pinned = cm_lp.pinned = on;
  }
  void reset() { tiles.clear(); giant_tiles.clear(); items.clear(); grp_win0.clear(); }  // what build_layout appends to
  std::vector<CamItem> items;
  std::vector<uint32_t> cam_item_start;
  uint32_t mv_G = 0, mv_W = 0;
  std::vector<uint32_t> grp_win0;
};

// Camera window of every group of G consecutive normal chunks (window kernel of the Schur operator): the run of W
// consecutive cameras, modulo ncam, that holds the most observations of the group.
static void build_windows(HostLayout& L, uint32_t ncam, uint32_t G, uint32_t W) {
  L.mv_G = G; L.mv_W = W;
  L.grp_win0.clear();
  if (!G || !W || !L.nnormal_chunks) { L.mv_G = L.mv_W = 0; return; }
  const uint32_t ngroups = (L.nnormal_chunks + G - 1) / G;
  L.grp_win0.assign(ngroups, 0);
  if (W >= ncam) return;  // every camera fits: window = [0, ncam)
#pragma omp parallel
  {
    std::vector<int64_t> off;
#pragma omp for schedule(dynamic, 16)
    for (int64_t g = 0; g < (int64_t)ngroups; ++g) {
      const size_t s0 = (size_t)g * G * TILE, s1 = std::min<size_t>((size_t)(g + 1) * G, L.nnormal_chunks) * TILE;
      off.clear();
      int64_t ref = -1;
      const int64_t half = ncam / 2;
      for (size_t s = s0; s < s1; ++s) {
        const uint32_t cam = L.cslot_meta[s].x;
        if (cam == PAD_CAM) continue;
        if (ref < 0) ref = cam;
        off.push_back((((int64_t)cam - ref + half) % ncam + ncam) % ncam - half);  // circular offset in [-ncam/2, ncam/2)
      }
      if (off.empty()) continue;
      std::sort(off.begin(), off.end());
      size_t best = 0, best_i = 0, j = 0;
      for (size_t i = 0; i < off.size(); ++i) {
        if (i && off[i] == off[i - 1]) continue;
        if (j < i) j = i;
        while (j < off.size() && off[j] < off[i] + (int64_t)W) ++j;
        if (j - i > best) { best = j - i; best_i = i; }
      }
      L.grp_win0[g] = (uint32_t)(((ref + off[best_i]) % ncam + ncam) % ncam);
    }
  }
}

static void build_layout(const apex_problem_desc* d, int nranks, int rank, HostLayout& L, int dc) {
  const bool timing = getenv("APEX_LAYOUT_TIMING") != nullptr;
  auto tprev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[layout] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tprev).count());
    tprev = now;
  };
  const uint64_t nobs = d->nobs;
  const uint32_t ncam = d->ncam;
  // ---- landmark sharding (block-cyclic) + point-major order of the local observations (stable in the caller's
  // insertion order) ----
  L.shard = ShardMap{d->npts, (uint32_t)nranks, (uint32_t)rank};
  L.npl = L.shard.count();
  const uint32_t npl = L.npl;
  // Stable counting sort by local landmark, in parallel: the observation list is cut into T contiguous pieces, piece t
  // counts its observations per landmark, a prefix over (landmark, piece) gives every piece its write cursor.
  std::vector<uint64_t> pt_start((size_t)npl + 1, 0);  // exclusive prefix of the LOCAL landmarks' observation counts
  HostVec<uint64_t> pm;
  {
    int T = std::max(1, std::min(omp_get_max_threads(), 16));
    while (T > 1 && (size_t)T * npl * sizeof(uint32_t) > ((size_t)1 << 30)) T /= 2;  // counters: at most 1 GiB
    if (nobs < 100000) T = 1;
    std::vector<HostVec<uint32_t>> cnt(T);
    auto piece = [&](int t) { return std::make_pair(nobs * (uint64_t)t / T, nobs * (uint64_t)(t + 1) / T); };
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      cnt[t].resize(npl);
      std::fill(cnt[t].begin(), cnt[t].end(), 0u);
      const auto [o0, o1] = piece(t);
      for (uint64_t o = o0; o < o1; ++o) {
        const uint32_t p = d->obs_pt[o];
        if (L.shard.owns(p)) cnt[t][L.shard.to_local(p)]++;
      }
    }
#pragma omp parallel for schedule(static)
    for (int64_t lp = 0; lp < (int64_t)npl; ++lp) {
      uint64_t k = 0;
      for (int t = 0; t < T; ++t) k += cnt[t][lp];
      pt_start[lp + 1] = k;
    }
    for (uint32_t lp = 0; lp < npl; ++lp) pt_start[lp + 1] += pt_start[lp];
    L.nobs_local = pt_start[npl];
    pm.resize(L.nobs_local);
#pragma omp parallel for schedule(static)
    for (int64_t lp = 0; lp < (int64_t)npl; ++lp) {  // counts -> cursors relative to pt_start (fits u32: nobs < 2^32)
      uint32_t run = 0;
      for (int t = 0; t < T; ++t) { const uint32_t k = cnt[t][lp]; cnt[t][lp] = run; run += k; }
    }
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      const auto [o0, o1] = piece(t);
      for (uint64_t o = o0; o < o1; ++o) {
        const uint32_t p = d->obs_pt[o];
        if (!L.shard.owns(p)) continue;
        const uint32_t lp = L.shard.to_local(p);
        pm[pt_start[lp] + cnt[t][lp]++] = o;
      }
    }
  }

  lap("point-major order");
  // ---- tiles ----
  // Normal tiles (<=256 observations, <=128 landmarks) get chunk ids 0..nn-1 in landmark order so that the
  // operator kernel's two thread groups can take chunks (2s, 2s+1); landmarks with more than 256 observations get
  // their chunks after those.
  L.pt_slot0.assign(npl, 0);
  L.pt_cnt.assign(npl, 0);
  std::vector<uint64_t> tile_q0;  // first point-major observation of each tile
  uint32_t chunk = 0;
  {
    uint32_t cur_pt0 = 0, cur_npt = 0, cur_obs = 0;
    uint64_t q = 0, cur_q0 = 0;
    auto flush = [&]() {
      if (cur_npt == 0) return;
      L.tiles.push_back({cur_pt0, cur_npt, chunk, 1});
      tile_q0.push_back(cur_q0);
      chunk += 1;
      cur_npt = 0; cur_obs = 0;
    };
    for (uint32_t lp = 0; lp < npl; ++lp) {
      const uint32_t k = (uint32_t)(pt_start[lp + 1] - pt_start[lp]);
      L.pt_cnt[lp] = k;
      if (k > (uint32_t)TILE) {
        flush();
        L.tiles.push_back({lp, 1, 0xFFFFFFFFu, (k + TILE - 1) / TILE});  // chunk0 assigned below
        tile_q0.push_back(q);
        q += k;
        continue;
      }
      if (cur_npt > 0 && (cur_obs + k > (uint32_t)TILE || cur_npt >= (uint32_t)MAX_TILE_PTS)) flush();
      if (cur_npt == 0) { cur_pt0 = lp; cur_q0 = q; }
      L.pt_slot0[lp] = chunk * TILE + cur_obs;
      cur_npt++;
      cur_obs += k;
      q += k;
    }
    flush();
  }
  L.nnormal_chunks = chunk;
  for (TileDesc& t : L.tiles)
    if (t.nchunks > 1) {
      t.chunk0 = chunk;
      L.pt_slot0[t.pt0] = chunk * TILE;
      chunk += t.nchunks;
      L.giant_tiles.push_back(t);
    }
  L.nchunks = chunk;
  const size_t nslots = (size_t)chunk * TILE;
  const uint32_t nchunk_even = (L.nnormal_chunks + 1) & ~1u;
  L.npairs = nchunk_even / 2;

  lap("tiles");
  // ---- slot arrays + per-chunk camera-sorted segment structure (parallel over tiles) ----
  // (uninitialised here; every chunk is filled with its padding defaults by the worker that builds it)
  L.slot_cam.resize(nslots);
  L.slot_lp.resize(nslots);
  L.slot_uv.resize(nslots * 2);
  L.slot_obs.resize(nslots);
  L.chunk_desc.assign(nchunk_even, ChunkDesc{0, 0, 0, 0});
  L.cslot_meta.resize((size_t)nchunk_even * TILE);
  L.cpt_meta.assign(npl, 0);
  L.cseg_cam.resize((size_t)nchunk_even * TILE);
  L.cseg_begin.resize((size_t)nchunk_even * CSEG_LD);
  auto clear_chunk = [&](size_t ch) {
    std::fill_n(L.slot_cam.begin() + ch * TILE, TILE, PAD_CAM);
    std::fill_n(L.slot_lp.begin() + ch * TILE, TILE, (uint16_t)0);
    std::fill_n(L.slot_uv.begin() + ch * 2 * TILE, 2 * TILE, 0.0);
    std::fill_n(L.slot_obs.begin() + ch * TILE, TILE, UINT64_MAX);
  };
  auto clear_tables = [&](size_t ch) {
    std::fill_n(L.cslot_meta.begin() + ch * TILE, TILE, make_uint2(PAD_CAM, 0));
    std::fill_n(L.cseg_cam.begin() + ch * TILE, TILE, 0u);
    std::fill_n(L.cseg_begin.begin() + ch * CSEG_LD, CSEG_LD, (uint16_t)0);
  };
  for (size_t ch = L.nnormal_chunks; ch < nchunk_even; ++ch) clear_tables(ch);  // the odd chunk of the last pair
  const int64_t ntiles = (int64_t)L.tiles.size();
  lap("slot array allocation");
#pragma omp parallel
  {
    std::vector<std::pair<uint32_t, uint32_t>> order;  // (camera, chunk-local slot)
#pragma omp for schedule(dynamic, 64)
    for (int64_t ti = 0; ti < ntiles; ++ti) {
      const TileDesc& t = L.tiles[ti];
      uint64_t q = tile_q0[ti];
      order.clear();
      for (uint32_t ch = t.chunk0; ch < t.chunk0 + t.nchunks; ++ch) clear_chunk(ch);
      if (t.nchunks == 1) clear_tables(t.chunk0);
      for (uint32_t i = 0; i < t.npt; ++i) {
        const uint32_t lp = t.pt0 + i;
        const uint32_t off = L.pt_slot0[lp] - t.chunk0 * TILE;
        if (t.nchunks == 1) L.cpt_meta[lp] = off | (L.pt_cnt[lp] << 16);
        for (uint32_t k = 0; k < L.pt_cnt[lp]; ++k, ++q) {
          const size_t slot = (size_t)L.pt_slot0[lp] + k;
          const uint64_t o = pm[q];
          const uint32_t cam = d->obs_cam[o];
          L.slot_cam[slot] = cam;
          L.slot_lp[slot] = (uint16_t)i;
          const size_t ch = slot / TILE, lane = slot % TILE;
          L.slot_uv[(ch * 2 + 0) * TILE + lane] = d->obs_uv[2 * o];
          L.slot_uv[(ch * 2 + 1) * TILE + lane] = d->obs_uv[2 * o + 1];
          L.slot_obs[slot] = o;
          if (t.nchunks == 1) {
            L.cslot_meta[slot].x = cam;
            L.cslot_meta[slot].y = i;
            order.push_back({cam, off + k});
          }
        }
      }
      if (t.nchunks != 1) continue;
      const uint32_t ch = t.chunk0;
      std::sort(order.begin(), order.end());
      uint32_t nseg = 0;
      for (size_t pos = 0; pos < order.size(); ++pos) {
        if (pos == 0 || order[pos].first != order[pos - 1].first) {
          L.cseg_cam[(size_t)ch * TILE + nseg] = order[pos].first;
          L.cseg_begin[(size_t)ch * CSEG_LD + nseg] = (uint16_t)pos;
          ++nseg;
        }
        L.cslot_meta[(size_t)ch * TILE + order[pos].second].y |= ((uint32_t)pos << 8) | ((nseg - 1) << 16);
      }
      L.cseg_begin[(size_t)ch * CSEG_LD + nseg] = (uint16_t)order.size();
      L.chunk_desc[ch] = ChunkDesc{t.pt0, t.npt, nseg, 0};
    }
  }

  lap("slots + segments");
  // ---- camera windows of the operator's chunk groups ----
  {
    const char* eg = getenv("APEX_MV_GROUP");
    const char* ew = getenv("APEX_MV_WINDOW");
    const uint32_t G = eg ? (uint32_t)std::max(0, atoi(eg)) : 8u;
    // opt-in (APEX_MV_WINDOW = cameras per window): measured slower than the chunk kernel
    const uint32_t W = ew ? std::min<uint32_t>(mv_window_cameras(dc, ncam), (uint32_t)std::max(0, atoi(ew))) : 0u;
    build_windows(L, ncam, G, W);
  }

  // ---- camera-major copy of the local observations + work items ----
  // Stable counting sort by camera of the point-major list, in parallel over contiguous pieces of that list.
  std::vector<uint32_t> cam_start((size_t)ncam + 1, 0);
  L.cm_uv.resize(2 * (size_t)L.nobs_local);
  L.cm_lp.resize(L.nobs_local);
  {
    const int T = L.nobs_local < 100000 ? 1 : std::max(1, std::min(omp_get_max_threads(), 64));
    std::vector<std::vector<uint32_t>> cnt(T, std::vector<uint32_t>(ncam, 0));
    std::vector<uint32_t> lp_cut(T + 1, npl);  // piece t = landmarks [lp_cut[t], lp_cut[t+1]), balanced by observations
    lp_cut[0] = 0;
    for (int t = 1; t < T; ++t)
      lp_cut[t] = (uint32_t)(std::lower_bound(pt_start.begin(), pt_start.end(), L.nobs_local * (uint64_t)t / T) - pt_start.begin());
    for (int t = 1; t <= T; ++t) lp_cut[t] = std::max(lp_cut[t], lp_cut[t - 1]);
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t)
      for (uint64_t q = pt_start[lp_cut[t]]; q < pt_start[lp_cut[t + 1]]; ++q) cnt[t][d->obs_cam[pm[q]]]++;
    for (uint32_t k = 0; k < ncam; ++k) {
      uint32_t run = cam_start[k];
      for (int t = 0; t < T; ++t) { const uint32_t n = cnt[t][k]; cnt[t][k] = run; run += n; }
      cam_start[k + 1] = run;
    }
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      uint64_t q = pt_start[lp_cut[t]];
      for (uint32_t lp = lp_cut[t]; lp < lp_cut[t + 1]; ++lp)
        for (uint32_t k = 0; k < L.pt_cnt[lp]; ++k, ++q) {
          const uint64_t o = pm[q];
          const uint32_t pos = cnt[t][d->obs_cam[o]]++;
          L.cm_uv[pos] = d->obs_uv[2 * o];
          L.cm_uv[(size_t)L.nobs_local + pos] = d->obs_uv[2 * o + 1];
          L.cm_lp[pos] = lp;
        }
    }
  }
  L.cam_item_start.assign((size_t)ncam + 1, 0);
  for (uint32_t k = 0; k < ncam; ++k) {
    L.cam_item_start[k] = (uint32_t)L.items.size();
    for (uint32_t b = cam_start[k]; b < cam_start[k + 1]; b += CAM_CHUNK)
      L.items.push_back({k, b, std::min<uint32_t>(b + CAM_CHUNK, cam_start[k + 1]), 0});
  }
  L.cam_item_start[ncam] = (uint32_t)L.items.size();
  lap("camera-major copy");
}

apex_status layout_stats(const apex_problem_desc* d, int nranks, int rank, apex_layout_stats* out, std::string& err) {
  APEX_TRY(validate_problem(d, err));
  const auto t0 = std::chrono::steady_clock::now();
  HostLayout L;
  const int K = model_intr_dim(d->camera_model);
  build_layout(d, nranks, rank, L, 6 + ((d->opt_flags & APEX_OPT_INTRINSIC) ? K : 0));
  out->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  out->shard_block = SHARD_BLOCK; out->npts_local = L.npl; out->nobs_local = L.nobs_local;
  out->ntiles = (uint32_t)L.tiles.size(); out->nlong_tiles = (uint32_t)L.giant_tiles.size();
  out->nchunks = L.nchunks; out->nnormal_chunks = L.nnormal_chunks; out->ncam_items = (uint32_t)L.items.size();
  uint64_t nseg = 0, maxseg = 0, covered = 0;
  for (const ChunkDesc& cd : L.chunk_desc) { nseg += cd.nseg; maxseg = std::max<uint64_t>(maxseg, cd.nseg); }
  for (uint64_t o : L.slot_obs) covered += o != UINT64_MAX;
  out->nsegments = nseg; out->max_segments_per_chunk = (uint32_t)maxseg; out->slots_used = covered;
  // structural self-check: every local observation sits in exactly one slot, camera-sorted positions are a
  // permutation of the chunk's observations and every segment is a run of one camera
  out->consistent = covered == L.nobs_local ? 1 : 0;
  for (const TileDesc& t : L.tiles) {
    if (t.nchunks != 1) continue;
    const size_t base = (size_t)t.chunk0 * TILE;
    int seen[TILE] = {0};
    uint32_t nobs_tile = 0;
    for (uint32_t i = 0; i < t.npt; ++i) nobs_tile += L.pt_cnt[t.pt0 + i];
    const ChunkDesc& cd = L.chunk_desc[t.chunk0];
    for (uint32_t s = 0; s < nobs_tile; ++s) {
      const uint2 m = L.cslot_meta[base + s];
      const uint32_t pos = (m.y >> 8) & 0xFFu, seg = (m.y >> 16) & 0xFFu;
      if (pos >= nobs_tile || seen[pos]++ || seg >= cd.nseg || L.cseg_cam[base + seg] != m.x) { out->consistent = 0; break; }
      const uint32_t b = L.cseg_begin[(size_t)t.chunk0 * CSEG_LD + seg], e = L.cseg_begin[(size_t)t.chunk0 * CSEG_LD + seg + 1];
      if (pos < b || pos >= e) { out->consistent = 0; break; }
    }
  }
  out->mv_group = L.mv_G; out->mv_window = L.mv_W; out->mv_ngroups = (uint32_t)L.grp_win0.size(); out->reserved2 = 0;
  uint64_t inwin = 0;
  if (L.mv_W)
    for (size_t s = 0; s < (size_t)L.nnormal_chunks * TILE; ++s) {
      const uint32_t cam = L.cslot_meta[s].x;
      if (cam == PAD_CAM) continue;
      const uint32_t w0 = L.grp_win0[s / ((size_t)L.mv_G * TILE)];
      if (w0 >= d->ncam) { out->consistent = 0; break; }
      const uint32_t l = cam >= w0 ? cam - w0 : cam + d->ncam - w0;
      inwin += l < L.mv_W;
    }
  out->nobs_in_window = inwin;
  return APEX_OK;
}

// rows of the owned landmarks out of a full [npts][3] host array
std::vector<double> gather_local_points(const Ctx& c, const double* pt_full) {
  std::vector<double> out((size_t)c.npl * 3);
  for (uint32_t lp = 0; lp < c.npl; ++lp) {
    const size_t g = c.shard.to_global(lp);
    out[3 * (size_t)lp] = pt_full[3 * g]; out[3 * (size_t)lp + 1] = pt_full[3 * g + 1]; out[3 * (size_t)lp + 2] = pt_full[3 * g + 2];
  }
  return out;
}

apex_status problem_upload(Ctx& c, const apex_problem_desc* d) {
  const bool timing = getenv("APEX_LAYOUT_TIMING") != nullptr;
  auto tprev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[upload] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tprev).count());
    tprev = now;
  };
  // One process per GPU on one node: every rank builds its layout at the same time, so each takes its share of the host
  // cores - not all of them, and not the single thread torchrun's default OMP_NUM_THREADS=1 would leave it (that made the
  // 2-GPU upload 220 ms against 49 ms on one GPU). APEX_HOST_THREADS overrides; single-rank runs keep OpenMP's own setting.
  const int omp_before = omp_get_max_threads();
  if (const char* ht = getenv("APEX_HOST_THREADS")) omp_set_num_threads(std::max(1, atoi(ht)));
  else if (c.nranks > 1) omp_set_num_threads(std::max(1, omp_get_num_procs() / c.nranks));
  struct OmpRestore { int n; ~OmpRestore() { omp_set_num_threads(n); } } omp_restore{omp_before};
  APEX_TRY(validate_problem(d, c.err));
  lap("validate");
  const int K = model_intr_dim(d->camera_model);
  c.have_problem = false;
  c.linearized = false;
  c.have_step = false;
  if (c.pcg_graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)c.pcg_graph_exec); c.pcg_graph_exec = nullptr; }
  c.model = d->camera_model; c.K = K; c.opt = d->opt_flags;
  c.opt_intr = (d->opt_flags & APEX_OPT_INTRINSIC) != 0;
  c.intr_vars = d->intr_vars_present != 0;
  c.dc = 6 + (c.opt_intr ? K : 0);
  c.np = 2 * (c.dc + 3);
  c.ncam = d->ncam; c.npts = d->npts; c.nobs = d->nobs;
  c.cam_dof_ref = (uint64_t)c.ncam * (6 + ((c.opt_intr || c.intr_vars) ? K : 0));
  c.loss_id = d->loss_id;
  for (int i = 0; i < 4; ++i) c.loss_p[i] = d->loss_params[i];

  if (!c.staging) {
    auto hl = std::make_shared<HostLayout>();
    hl->set_pinned(getenv("APEX_NO_PINNED_STAGING") == nullptr);
    c.staging = hl;
  }
  HostLayout& L = *static_cast<HostLayout*>(c.staging.get());
  L.reset();
  build_layout(d, c.nranks, c.rank, L, c.dc);
  c.mv_G = L.mv_G; c.mv_W = L.mv_W; c.mv_ngroups = (uint32_t)L.grp_win0.size();
  c.shard = L.shard; c.npl = L.npl; c.nobs_local = L.nobs_local;
  c.nnormal_chunks = L.nnormal_chunks; c.nchunks = L.nchunks; c.npairs = L.npairs;
  c.ntiles = (uint32_t)L.tiles.size(); c.ngiant = (uint32_t)L.giant_tiles.size(); c.nitems = (uint32_t)L.items.size();
  c.nslots = (size_t)L.nchunks * TILE;
  c.slot_obs.swap(L.slot_obs);
  c.h_pt_cnt = L.pt_cnt;
  lap("build_layout");

  // ---- fixed masks ----
  std::vector<uint8_t> pose_fixed(c.ncam, 0), pt_fixed(c.npl, 0);
  std::vector<uint16_t> intr_fixed(c.ncam, 0);
  if (d->pose_fixed) std::copy(d->pose_fixed, d->pose_fixed + c.ncam, pose_fixed.begin());
  if (d->intr_fixed) std::copy(d->intr_fixed, d->intr_fixed + c.ncam, intr_fixed.begin());
  if (d->pt_fixed) for (uint32_t lp = 0; lp < c.npl; ++lp) pt_fixed[lp] = d->pt_fixed[c.shard.to_global(lp)];

  // ---- to the device ----
  cudaStream_t s = c.stream;
  g_h2d_bytes = 0;
  APEX_CUDA_TRY(c, upload_vec(c.tiles, L.tiles, s));
  APEX_CUDA_TRY(c, upload_vec(c.giant_tiles, L.giant_tiles, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_cam, L.slot_cam, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_lp, L.slot_lp, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_uv, L.slot_uv, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_slot0, L.pt_slot0, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_cnt, L.pt_cnt, s));
  APEX_CUDA_TRY(c, upload_vec(c.chunk_desc, L.chunk_desc, s));
  APEX_CUDA_TRY(c, upload_vec(c.cslot_meta, L.cslot_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.cpt_meta, L.cpt_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.cseg_cam, L.cseg_cam, s));
  APEX_CUDA_TRY(c, upload_vec(c.cseg_begin, L.cseg_begin, s));
  APEX_CUDA_TRY(c, upload_vec(c.grp_win0, L.grp_win0, s));
  APEX_CUDA_TRY(c, upload_vec(c.items, L.items, s));
  APEX_CUDA_TRY(c, upload_vec(c.cam_item_start, L.cam_item_start, s));
  APEX_CUDA_TRY(c, upload_vec(c.cm_uv, L.cm_uv, s));
  APEX_CUDA_TRY(c, upload_vec(c.cm_lp, L.cm_lp, s));
  APEX_CUDA_TRY(c, upload_vec(c.pose_fixed, pose_fixed, s));
  APEX_CUDA_TRY(c, upload_vec(c.intr_fixed, intr_fixed, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_fixed, pt_fixed, s));

  lap("H2D of the structure (enqueue)");
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
  APEX_CUDA_TRY(c, c.xpad.alloc((size_t)c.ncam * xpad_stride(c.dc)));
  APEX_CUDA_TRY(c, c.step_cam.alloc(ncd));
  APEX_CUDA_TRY(c, c.step_pt.alloc((size_t)c.npl * 3));
  APEX_CUDA_TRY(c, c.red_scratch.alloc(8 * (size_t)std::max<uint32_t>(std::max<uint32_t>(c.nchunks, c.ncam), 1024u) + 64));
  APEX_CUDA_TRY(c, cudaMemsetAsync(c.step_cam.p, 0, ncd * sizeof(double), s));
  APEX_CUDA_TRY(c, cudaMemsetAsync(c.step_pt.p, 0, std::max<size_t>((size_t)c.npl * 3, 1) * sizeof(double), s));

  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.pose.p, d->pose, (size_t)c.ncam * 7 * sizeof(double), cudaMemcpyHostToDevice, s));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.intr.p, d->intr, (size_t)c.ncam * K * sizeof(double), cudaMemcpyHostToDevice, s));
  std::vector<double> pt_local;  // one rank owns every landmark in the caller's order: no gather needed
  if (c.nranks > 1) pt_local = gather_local_points(c, d->pt);
  if (c.npl) APEX_CUDA_TRY(c, cudaMemcpyAsync(c.pt.p, c.nranks > 1 ? pt_local.data() : d->pt, (size_t)c.npl * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  c.upload_h2d_bytes = g_h2d_bytes + ((size_t)c.ncam * (7 + K) + (size_t)c.npl * 3) * sizeof(double);
  lap("allocations + parameters");
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));  // the host vectors above die with this scope
  lap("stream sync");
  APEX_TRY(setup_peer_allreduce(c, ncd));
  lap("peer setup");
  c.have_problem = true;
  return APEX_OK;
}

}  // namespace apex
