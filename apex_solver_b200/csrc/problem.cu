// problem.cu — apex_problem_upload: replaces Problem construction + initialize_optimization_state
// (src/core/problem.rs:518-808, src/optimizer/mod.rs:522-563) for the SoA factor graph of
// bin/bundle_adjustment.rs:212-441. Builds the static observation structure described in apex_ctx.h:
// landmark sharding across ranks, point-major tiles with per-chunk camera segments, camera-major work items.
// build_layout() is pure host code (OpenMP over tiles) so it can be exercised without a device.
#include <algorithm>
#include <functional>
#include <thread>
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

// Pageable host memory goes to the device through the context's page-locked bounce ring (two buffers, allocated at context
// creation): the workers copy piece k+1 into one buffer while the DMA engine drains the other. cudaMemcpyAsync straight from
// pageable memory goes through the driver's own staging at 3-10 GB/s (measured: 316 MB of structure in 33-97 ms).
static thread_local Ctx* g_bounce_ctx = nullptr;
static cudaError_t bounce_copy(void* dev, const void* host, size_t bytes, cudaStream_t s) {
  Ctx* c = g_bounce_ctx;
  if (!c || !c->bounce[0] || bytes < ((size_t)1 << 20)) return cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s);
  const size_t piece = Ctx::BOUNCE_BYTES;
  for (size_t off = 0; off < bytes; off += piece) {
    const size_t nb = std::min(piece, bytes - off);
    const int k = c->bounce_next;
    c->bounce_next ^= 1;
    cudaError_t e = cudaEventSynchronize(c->bounce_ev[k]);   // the DMA that last read this buffer is done
    if (e != cudaSuccess) return e;
    const char* src = static_cast<const char*>(host) + off;
    char* dst = static_cast<char*>(c->bounce[k]);
    const int T = std::max(1, std::min(omp_get_max_threads(), 16));
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      const size_t b0 = nb * (size_t)t / T, b1 = nb * (size_t)(t + 1) / T;
      std::memcpy(dst + b0, src + b0, b1 - b0);
    }
    e = cudaMemcpyAsync(static_cast<char*>(dev) + off, dst, nb, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(c->bounce_ev[k], s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

template <typename T>
static cudaError_t upload_vec(DevBuf<T>& buf, const StageBuf<T>& v, cudaStream_t s) {
  cudaError_t e = buf.alloc(v.size());
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  g_h2d_bytes += v.size() * sizeof(T);
  if (!v.is_pinned_alloc) return bounce_copy(buf.p, v.data(), v.size() * sizeof(T), s);
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
  if ((d->opt_flags & APEX_OPT_SHARED_INTRINSICS) && (d->opt_flags & APEX_OPT_INTRINSIC)) {
    if (d->loss_id != APEX_LOSS_NONE && d->loss_id != APEX_LOSS_L2) { err = "shared intrinsics: the loss must be NONE or L2"; return APEX_ERR_UNSUPPORTED; }
    if (d->obs_loss) { err = "shared intrinsics: per-block losses are not supported"; return APEX_ERR_UNSUPPORTED; }
  }
  if (d->ncam == 0) { err = "No camera variables found"; return APEX_ERR_INVALID_INPUT; }    // explicit_schur.rs:278-282
  if (d->npts == 0) { err = "No landmark variables found"; return APEX_ERR_INVALID_INPUT; }  // explicit_schur.rs:283-287
  if (d->nobs > 0xFFFFFFF0ull) { err = "too many observations for u32 slots"; return APEX_ERR_UNSUPPORTED; }
  if (d->loss_id < APEX_LOSS_NONE || d->loss_id > APEX_LOSS_T_DISTRIBUTION) { err = "unknown loss id"; return APEX_ERR_INVALID_INPUT; }
  const uint64_t nobs = d->nobs;
  if (d->obs_loss) {
    if (!d->loss_table || d->n_losses < 1 || d->n_losses > 256) { err = "obs_loss needs a loss_table of 1..256 entries"; return APEX_ERR_INVALID_INPUT; }
    for (int i = 0; i < d->n_losses; ++i)
      if (d->loss_table[i].loss_id < APEX_LOSS_NONE || d->loss_table[i].loss_id > APEX_LOSS_T_DISTRIBUTION) { err = "unknown loss id in loss_table"; return APEX_ERR_INVALID_INPUT; }
  }
  int bad = 0, bad_loss = 0;
#pragma omp parallel for reduction(| : bad, bad_loss) schedule(static)
  for (int64_t o = 0; o < (int64_t)nobs; ++o) {
    bad |= (d->obs_cam[o] >= d->ncam || d->obs_pt[o] >= d->npts) ? 1 : 0;
    if (d->obs_loss) bad_loss |= d->obs_loss[o] >= d->n_losses ? 1 : 0;
  }
  if (bad) { err = "observation index out of range"; return APEX_ERR_INVALID_INPUT; }
  if (bad_loss) { err = "obs_loss index out of range"; return APEX_ERR_INVALID_INPUT; }
  return APEX_OK;
}

// The static structure of one rank's shard, on the host.
struct HostLayout {
  ShardMap shard;
  uint32_t npl = 0;
  uint64_t nobs_local = 0;
  uint32_t nnormal_chunks = 0, nchunks = 0;
  uint64_t nwseg_total = 0;   // warp-segments of all normal chunks (statistics)
  std::vector<TileDesc> tiles, giant_tiles;
  std::vector<uint32_t> pt_slot0, pt_cnt;
  StageBuf<uint32_t> slot_cam;
  StageBuf<uint16_t> slot_lp;
  StageBuf<double> slot_uv;
  StageBuf<uint8_t> slot_loss, cm_loss;   // per-block loss indices (only when the problem has obs_loss)
  HostVec<uint64_t> slot_obs;
  HostVec<uint8_t> slot_pos;
  std::vector<ChunkDesc> chunk_desc;
  StageBuf<uint2> cslot_meta;
  std::vector<uint32_t> cpt_meta;
  StageBuf<uint16_t> cslot_widx;
  std::vector<WinDesc> win_desc;
  std::vector<uint32_t> range_win0, win_cams, cam_row_start;
  StageBuf<uint32_t> win_dst;
  StageBuf<double> cm_uv;
  StageBuf<uint32_t> cm_lp;
  void set_pinned(bool on) { slot_cam.pinned = slot_lp.pinned = slot_uv.pinned = cslot_meta.pinned = cslot_widx.pinned = win_dst.pinned = cm_uv.pinned = cm_lp.pinned = on; }
  void reset() { tiles.clear(); giant_tiles.clear(); items.clear(); }  // what build_layout appends to
  std::vector<CamItem> items;
  std::vector<uint32_t> cam_item_start;
};

// `on_tiles(nchunks, nnormal_chunks, nobs_local, npl)`, if set, is called as soon as the chunk count is known (after the point-major
// sort and the serial tile cut, about a third into the build): the caller starts its device allocations there.
static void build_layout(const apex_problem_desc* d, int nranks, int rank, HostLayout& L, uint32_t nctas, uint32_t W, bool want_det_lists,
                         const std::function<void(uint32_t, uint32_t, uint64_t, uint32_t)>& on_tiles = nullptr) {
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
  HostVec<uint32_t> pm;   // point-major order: observation index (nobs < 2^32, validate_problem)
  {
    int T = std::max(1, std::min(omp_get_max_threads(), 16));
    while (T > 1 && (size_t)T * npl * sizeof(uint32_t) > ((size_t)1 << 30)) T /= 2;  // counters: at most 1 GiB
    if (nobs < 100000) T = 1;
    std::vector<HostVec<uint32_t>> cnt(T);
    auto piece = [&](int t) { return std::make_pair(nobs * (uint64_t)t / T, nobs * (uint64_t)(t + 1) / T); };
    // ShardMap::owns / to_local per observation, twice: with a runtime rank count that is two integer divisions per call - most of
    // this phase on a rank of several, which scans ALL observations. Powers of two (the usual rank counts) get shifts and masks.
    static_assert((SHARD_BLOCK & (SHARD_BLOCK - 1)) == 0, "block size is a power of two");
    const uint32_t nr = (uint32_t)nranks, rk = (uint32_t)rank;
    const bool pow2 = (nr & (nr - 1)) == 0;
    uint32_t rsh = 0, bsh = 0;
    while ((1u << rsh) < nr) ++rsh;
    while ((1u << bsh) < SHARD_BLOCK) ++bsh;
    auto owns = [&](uint32_t p) { const uint32_t b = p >> bsh; return pow2 ? (b & (nr - 1)) == rk : b % nr == rk; };
    auto to_local = [&](uint32_t p) { const uint32_t b = p >> bsh; return ((pow2 ? (b - rk) >> rsh : (b - rk) / nr) << bsh) | (p & (SHARD_BLOCK - 1)); };
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      cnt[t].resize(npl);
      std::fill(cnt[t].begin(), cnt[t].end(), 0u);
      const auto [o0, o1] = piece(t);
      for (uint64_t o = o0; o < o1; ++o) {
        const uint32_t p = d->obs_pt[o];
        if (owns(p)) cnt[t][to_local(p)]++;
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
        if (!owns(p)) continue;
        const uint32_t lp = to_local(p);
        pm[pt_start[lp] + cnt[t][lp]++] = (uint32_t)o;
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

  lap("tiles");
  if (on_tiles) on_tiles(L.nchunks, L.nnormal_chunks, L.nobs_local, npl);
  // ---- slot arrays + the camera-sorted lane order of every normal chunk (parallel over tiles) ----
  // (uninitialised here; every chunk is filled with its padding defaults by the worker that builds it)
  L.slot_cam.resize(nslots);
  L.slot_lp.resize(nslots);
  L.slot_uv.resize(nslots * 2);
  L.slot_obs.resize(nslots);
  L.slot_pos.resize(nslots);
  L.slot_loss.resize(d->obs_loss ? nslots : 0);
  L.chunk_desc.assign(L.nnormal_chunks, ChunkDesc{0, 0, 0, 0});
  L.cslot_meta.resize((size_t)L.nnormal_chunks * TILE);
  L.cpt_meta.assign(npl, 0);
  // padding defaults of lanes [l0, TILE) of chunk ch (the observation lanes in front of them are written exactly once below:
  // filling whole chunks first wrote every slot array twice)
  auto pad_chunk = [&](size_t ch, uint32_t l0, bool normal) {
    const size_t n = (size_t)TILE - l0, b = ch * TILE + l0;
    if (n == 0) return;
    std::fill_n(L.slot_cam.begin() + b, n, PAD_CAM);
    std::fill_n(L.slot_lp.begin() + b, n, (uint16_t)0);
    std::fill_n(L.slot_uv.begin() + ch * 2 * TILE + l0, n, 0.0);
    std::fill_n(L.slot_uv.begin() + (ch * 2 + 1) * TILE + l0, n, 0.0);
    std::fill_n(L.slot_obs.begin() + b, n, UINT64_MAX);
    if (d->obs_loss) std::fill_n(L.slot_loss.begin() + b, n, (uint8_t)0);
    for (uint32_t t = l0; t < (uint32_t)TILE; ++t) L.slot_pos[ch * TILE + t] = (uint8_t)t;  // default: camera half at the slot's own lane
    if (normal) std::fill_n(L.cslot_meta.begin() + b, n, make_uint2(PAD_CAM, 0));
  };
  const int64_t ntiles = (int64_t)L.tiles.size();
  int cam_bits = 8;   // radix passes of the per-chunk camera sort cover the bits of ncam - 1
  while (cam_bits < 32 && ((uint64_t)1 << cam_bits) < (uint64_t)ncam) cam_bits += 8;
  lap("slot array allocation");
#pragma omp parallel
  {
    struct CamLane { uint32_t first, second; };         // (camera, chunk-local point-major lane)
    std::vector<uint64_t> keys, tmp;                    // camera << 32 | lane
    std::vector<CamLane> order;
#pragma omp for schedule(dynamic, 64)
    for (int64_t ti = 0; ti < ntiles; ++ti) {
      const TileDesc& t = L.tiles[ti];
      uint64_t q = tile_q0[ti];
      keys.clear();
      {
        uint64_t tobs = 0;
        for (uint32_t i = 0; i < t.npt; ++i) tobs += L.pt_cnt[t.pt0 + i];
        if (t.nchunks == 1) pad_chunk(t.chunk0, (uint32_t)tobs, true);
        else {   // one landmark over several chunks: camera halves stay at the observations' own lanes; the last chunk's tail is padding
          for (uint32_t ch = t.chunk0; ch < t.chunk0 + t.nchunks; ++ch)
            for (int l = 0; l < TILE; ++l) L.slot_pos[(size_t)ch * TILE + l] = (uint8_t)l;
          pad_chunk(t.chunk0 + t.nchunks - 1, (uint32_t)(tobs - (uint64_t)(t.nchunks - 1) * TILE), false);
        }
      }
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
          if (d->obs_loss) L.slot_loss[slot] = d->obs_loss[o];
          if (t.nchunks == 1) {
            L.cslot_meta[slot].y = i;  // [7:0] chunk-local landmark of point-major lane off + k
            keys.push_back((uint64_t)cam << 32 | (off + k));
          }
        }
      }
      if (t.nchunks != 1) continue;
      // camera-sorted lanes (the (camera, lane) pairs are unique, so the order is fully determined); a run of one camera
      // that crosses a warp boundary is marked on the first lane of the following warp(s) (continuation flags, apex_ctx.h)
      const uint32_t ch = t.chunk0;
      const size_t base = (size_t)ch * TILE;
      // stable LSD radix sort by camera, 8 bits per pass (the keys were pushed in lane order, so equal cameras keep it): a chunk has
      // at most 256 keys and ncam needs two passes up to 65 536 cameras - a third of the time std::sort took on the Venice shape
      {
        const size_t n = keys.size();
        tmp.resize(n);
        uint64_t* a = keys.data();
        uint64_t* b = tmp.data();
        for (int shift = 32; shift < 32 + cam_bits; shift += 8) {
          uint16_t cnt[256] = {0};
          for (size_t i = 0; i < n; ++i) cnt[(a[i] >> shift) & 0xFF]++;
          uint16_t run = 0;
          for (int v = 0; v < 256; ++v) { const uint16_t c = cnt[v]; cnt[v] = run; run = (uint16_t)(run + c); }
          for (size_t i = 0; i < n; ++i) b[cnt[(a[i] >> shift) & 0xFF]++] = a[i];
          std::swap(a, b);
        }
        if (a != keys.data()) std::copy(a, a + n, keys.data());
      }
      order.resize(keys.size());
      for (size_t i = 0; i < keys.size(); ++i) order[i] = CamLane{(uint32_t)(keys[i] >> 32), (uint32_t)keys[i]};
      uint32_t nwseg = 0;
      for (size_t pos = 0; pos < order.size(); ++pos) {
        if (pos % 32 == 0 || order[pos].first != order[pos - 1].first) ++nwseg;
        const uint32_t ptl = order[pos].second;
        L.cslot_meta[base + pos].x = order[pos].first;
        L.cslot_meta[base + pos].y |= ptl << 16;
        L.cslot_meta[base + ptl].y |= (uint32_t)pos << 8;
        L.slot_pos[base + ptl] = (uint8_t)pos;
      }
      auto is_cont = [&](size_t pos) { return pos >= 32 && pos < order.size() && order[pos].first == order[pos - 1].first; };
      for (size_t pos = 32; pos < order.size(); pos += 32) {
        if (!is_cont(pos)) continue;
        uint32_t fl = CONT_BIT;
        if (!(is_cont(pos - 32) && order[pos - 32].first == order[pos].first)) {
          uint32_t len = 1;
          while (is_cont(pos + 32 * len) && order[pos + 32 * len].first == order[pos].first) ++len;
          fl |= CONT_FIRST_BIT | (len << CONT_LEN_SHIFT);
        }
        L.cslot_meta[base + pos].y |= fl;
      }
      L.chunk_desc[ch] = ChunkDesc{t.pt0, t.npt, nwseg, (uint32_t)order.size()};
    }
  }
  L.nwseg_total = 0;
  for (uint32_t ch = 0; ch < L.nnormal_chunks; ++ch) L.nwseg_total += L.chunk_desc[ch].nwseg;
  lap("slots + camera-sorted lanes");

  // ---- ranges and windows of the chunk kernel (apex_ctx.h): range r = chunks [nn*r/P, nn*(r+1)/P); inside a range, greedy
  // windows of consecutive chunks whose observations touch at most W distinct cameras ----
  {
    const uint32_t nn = L.nnormal_chunks;
    const uint32_t P = std::max<uint32_t>(1, std::min<uint32_t>(nn, nctas));
    std::vector<std::vector<WinDesc>> rw(P);
    std::vector<std::vector<uint32_t>> rc(P);
    L.cslot_widx.resize((size_t)nn * TILE);
#pragma omp parallel
    {
      // per thread: stamp[cam] == epoch <=> cam is in the current window; idx[cam] = its window-local index once the window closes
      std::vector<uint32_t> cur, fresh, stamp(ncam, 0);
      std::vector<uint16_t> idx(ncam, 0);
      uint32_t epoch = 0;
#pragma omp for schedule(dynamic, 4)
      for (int64_t r = 0; r < (int64_t)P; ++r) {
        const uint32_t c0 = (uint32_t)((uint64_t)nn * r / P), c1 = (uint32_t)((uint64_t)nn * (r + 1) / P);
        auto close = [&](uint32_t cb, uint32_t ce) {
          // window [cb, ce) with the camera list `cur`: sorted, it is the window's row order; every camera-sorted lane gets its row
          std::sort(cur.begin(), cur.end());
          rw[r].push_back(WinDesc{cb, ce, (uint32_t)rc[r].size(), (uint32_t)cur.size()});
          rc[r].insert(rc[r].end(), cur.begin(), cur.end());
          for (size_t j = 0; j < cur.size(); ++j) idx[cur[j]] = (uint16_t)j;
          for (uint32_t ch = cb; ch < ce; ++ch) {
            const size_t base = (size_t)ch * TILE;
            const uint32_t n = L.chunk_desc[ch].nobs;
            for (uint32_t pos = 0; pos < n; ++pos) L.cslot_widx[base + pos] = idx[L.cslot_meta[base + pos].x];
            for (uint32_t pos = n; pos < (uint32_t)TILE; ++pos) L.cslot_widx[base + pos] = 0;
          }
        };
        cur.clear();
        ++epoch;
        uint32_t wb = c0;
        for (uint32_t ch = c0; ch < c1; ++ch) {
          const size_t base = (size_t)ch * TILE;
          const uint32_t n = L.chunk_desc[ch].nobs;
          // cameras of this chunk the window does not hold yet (a camera is one run of the camera-sorted lanes)
          auto collect = [&]() {
            fresh.clear();
            for (uint32_t pos = 0; pos < n; ++pos) {
              const uint32_t cam = L.cslot_meta[base + pos].x;
              if ((pos == 0 || cam != L.cslot_meta[base + pos - 1].x) && stamp[cam] != epoch) fresh.push_back(cam);
            }
          };
          collect();
          if (cur.size() + fresh.size() > W && ch > wb) {   // the chunk would take the window beyond W cameras: it starts the next one
            close(wb, ch);
            wb = ch;
            cur.clear();
            ++epoch;
            collect();
          }
          for (uint32_t cam : fresh) { stamp[cam] = epoch; cur.push_back(cam); }
        }
        if (c1 > wb) close(wb, c1);
      }
    }
    L.range_win0.assign((size_t)P + 1, 0);
    uint64_t nrows = 0;
    for (uint32_t r = 0; r < P; ++r) L.range_win0[r + 1] = L.range_win0[r] + (uint32_t)rw[r].size();
    L.win_desc.clear();
    L.win_desc.reserve(L.range_win0[P]);
    for (uint32_t r = 0; r < P; ++r) nrows += rc[r].size();
    L.win_cams.clear();
    L.win_cams.reserve(nrows);
    for (uint32_t r = 0; r < P; ++r) {
      const uint32_t off = (uint32_t)L.win_cams.size();
      for (WinDesc w : rw[r]) { w.cam0 += off; L.win_desc.push_back(w); }
      L.win_cams.insert(L.win_cams.end(), rc[r].begin(), rc[r].end());
    }
    if (nn == 0) { L.range_win0.assign(1, 0); }
    // deterministic flush: the partial results are stored camera-major - camera c owns rows [cam_row_start[c], cam_row_start[c+1]),
    // one per window that touches it, in (range, window) order; win_dst maps every entry of win_cams to its row
    L.cam_row_start.assign((size_t)ncam + 1, 0);
    L.win_dst.resize(0);
    if (want_det_lists) {
      for (uint32_t cam : L.win_cams) L.cam_row_start[cam + 1]++;
      for (uint32_t k = 0; k < ncam; ++k) L.cam_row_start[k + 1] += L.cam_row_start[k];
      L.win_dst.resize(L.win_cams.size());
      std::vector<uint32_t> cursor(L.cam_row_start.begin(), L.cam_row_start.end() - 1);
      for (size_t i = 0; i < L.win_cams.size(); ++i) L.win_dst[i] = cursor[L.win_cams[i]]++;
    }
  }
  lap("ranges + windows");
  // ---- camera-major copy of the local observations + work items ----
  // Stable counting sort by camera of the point-major list, in parallel over contiguous pieces of that list.
  std::vector<uint32_t> cam_start((size_t)ncam + 1, 0);
  L.cm_uv.resize(2 * (size_t)L.nobs_local);
  L.cm_lp.resize(L.nobs_local);
  L.cm_loss.resize(d->obs_loss ? L.nobs_local : 0);
  {
    const int T = L.nobs_local < 100000 ? 1 : std::max(1, std::min(omp_get_max_threads(), 64));
    std::vector<std::vector<uint32_t>> cnt(T, std::vector<uint32_t>(ncam, 0));
    std::vector<uint32_t> lp_cut(T + 1, npl);  // piece t = landmarks [lp_cut[t], lp_cut[t+1]), balanced by observations
    lp_cut[0] = 0;
    for (int t = 1; t < T; ++t)
      lp_cut[t] = (uint32_t)(std::lower_bound(pt_start.begin(), pt_start.end(), L.nobs_local * (uint64_t)t / T) - pt_start.begin());
    for (int t = 1; t <= T; ++t) lp_cut[t] = std::max(lp_cut[t], lp_cut[t - 1]);
    // (cameras, pixels and loss classes are read back from the slot arrays filled above: they hold the observations in this very
    // order, contiguously per landmark, where d->obs_*[pm[q]] would be one cache miss per observation and array)
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t)
      for (uint32_t lp = lp_cut[t]; lp < lp_cut[t + 1]; ++lp) {
        const uint32_t* sc = L.slot_cam.data() + L.pt_slot0[lp];
        for (uint32_t k = 0; k < L.pt_cnt[lp]; ++k) cnt[t][sc[k]]++;
      }
    for (uint32_t k = 0; k < ncam; ++k) {
      uint32_t run = cam_start[k];
      for (int t = 0; t < T; ++t) { const uint32_t n = cnt[t][k]; cnt[t][k] = run; run += n; }
      cam_start[k + 1] = run;
    }
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      for (uint32_t lp = lp_cut[t]; lp < lp_cut[t + 1]; ++lp)
        for (uint32_t k = 0; k < L.pt_cnt[lp]; ++k) {
          const size_t slot = (size_t)L.pt_slot0[lp] + k, ch = slot / TILE, lane = slot % TILE;
          const uint32_t pos = cnt[t][L.slot_cam[slot]]++;
          L.cm_uv[pos] = L.slot_uv[(ch * 2 + 0) * TILE + lane];
          L.cm_uv[(size_t)L.nobs_local + pos] = L.slot_uv[(ch * 2 + 1) * TILE + lane];
          L.cm_lp[pos] = lp;
          if (d->obs_loss) L.cm_loss[pos] = L.slot_loss[slot];
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
  const int dc = 6 + ((d->opt_flags & APEX_OPT_INTRINSIC) ? K : 0);
  uint32_t W; bool staged;
  schur_plan(dc, W, staged);
  build_layout(d, nranks, rank, L, (staged ? 2 : 3) * 148, W, true);
  out->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  out->shard_block = SHARD_BLOCK; out->npts_local = L.npl; out->nobs_local = L.nobs_local;
  out->ntiles = (uint32_t)L.tiles.size(); out->nlong_tiles = (uint32_t)L.giant_tiles.size();
  out->nchunks = L.nchunks; out->nnormal_chunks = L.nnormal_chunks; out->ncam_items = (uint32_t)L.items.size();
  uint64_t maxseg = 0, covered = 0;
  for (const ChunkDesc& cd : L.chunk_desc) maxseg = std::max<uint64_t>(maxseg, cd.nwseg);
  for (uint64_t o : L.slot_obs) covered += o != UINT64_MAX;
  out->nsegments = L.nwseg_total; out->max_segments_per_chunk = (uint32_t)maxseg; out->slots_used = covered;
  out->mv_ranges = (uint32_t)L.range_win0.size() - 1; out->mv_window = W; out->mv_nwindows = (uint32_t)L.win_desc.size();
  out->mv_rows = L.win_cams.size();
  // structural self-check: every local observation sits in exactly one slot; inside a normal chunk the camera-sorted lanes are
  // a permutation of the point-major lanes, both maps are inverse to each other, cameras ascend along the sorted lanes, the
  // continuation flags mark exactly the runs cut at warp boundaries; the windows tile the ranges, the ranges tile the normal
  // chunks, a window's camera list is sorted, at most W long (or one chunk) and every lane's window index names its camera;
  // the camera-major rows of the deterministic flush cover every window entry once
  out->consistent = covered == L.nobs_local ? 1 : 0;
  for (const TileDesc& t : L.tiles) {
    if (t.nchunks != 1) continue;
    const size_t base = (size_t)t.chunk0 * TILE;
    const ChunkDesc& cd = L.chunk_desc[t.chunk0];
    uint32_t nobs_tile = 0;
    for (uint32_t i = 0; i < t.npt; ++i) nobs_tile += L.pt_cnt[t.pt0 + i];
    if (cd.nobs != nobs_tile || cd.pt0 != t.pt0 || cd.npt != t.npt) { out->consistent = 0; break; }
    uint32_t wseg = 0;
    for (uint32_t s = 0; s < nobs_tile && out->consistent; ++s) {
      const uint2 m = L.cslot_meta[base + s];
      const uint32_t pos = (m.y >> 8) & 0xFFu, ipos = (m.y >> 16) & 0xFFu;
      if (pos >= nobs_tile || ipos >= nobs_tile || ((L.cslot_meta[base + pos].y >> 16) & 0xFFu) != s || L.cslot_meta[base + pos].x != L.slot_cam[base + s] ||
          L.slot_pos[base + s] != pos || (m.y & 0xFFu) != L.slot_lp[base + s])
        out->consistent = 0;
      if (s > 0 && m.x < L.cslot_meta[base + s - 1].x) out->consistent = 0;
      if (s % 32 == 0 || m.x != L.cslot_meta[base + s - 1].x) ++wseg;
      const bool cont = s % 32 == 0 && s > 0 && m.x == L.cslot_meta[base + s - 1].x;
      if (cont != ((m.y & CONT_BIT) != 0)) out->consistent = 0;
      if (cont) {
        const bool first = !((L.cslot_meta[base + s - 32].y & CONT_BIT) && L.cslot_meta[base + s - 32].x == m.x);
        if (first != ((m.y & CONT_FIRST_BIT) != 0)) out->consistent = 0;
        if (first) {
          uint32_t len = 1;
          while (s + 32 * len < nobs_tile && (L.cslot_meta[base + s + 32 * len].y & CONT_BIT) && L.cslot_meta[base + s + 32 * len].x == m.x) ++len;
          if (((m.y >> CONT_LEN_SHIFT) & 7u) != len) out->consistent = 0;
        }
      } else if (m.y >> 24) out->consistent = 0;
    }
    if (wseg != cd.nwseg) out->consistent = 0;
    for (uint32_t s = nobs_tile; s < (uint32_t)TILE; ++s)
      if (L.cslot_meta[base + s].x != PAD_CAM) out->consistent = 0;
    if (!out->consistent) break;
  }
  {
    const uint32_t nn = L.nnormal_chunks, P = (uint32_t)L.range_win0.size() - 1;
    uint32_t next_chunk = 0, next_row = 0;
    for (uint32_t r = 0; r < P && out->consistent; ++r) {
      if (nn && next_chunk != (uint32_t)((uint64_t)nn * r / P)) out->consistent = 0;
      for (uint32_t w = L.range_win0[r]; w < L.range_win0[r + 1] && out->consistent; ++w) {
        const WinDesc& wd = L.win_desc[w];
        if (wd.chunk_begin != next_chunk || wd.chunk_end <= wd.chunk_begin || wd.cam0 != next_row || wd.ncams == 0) { out->consistent = 0; break; }
        if (wd.ncams > W && wd.chunk_end - wd.chunk_begin != 1) out->consistent = 0;
        for (uint32_t i = 1; i < wd.ncams; ++i) if (L.win_cams[wd.cam0 + i] <= L.win_cams[wd.cam0 + i - 1]) out->consistent = 0;
        for (uint32_t ch = wd.chunk_begin; ch < wd.chunk_end && out->consistent; ++ch)
          for (uint32_t s = 0; s < L.chunk_desc[ch].nobs; ++s) {
            const uint32_t wi = L.cslot_widx[(size_t)ch * TILE + s];
            if (wi >= wd.ncams || L.win_cams[wd.cam0 + wi] != L.cslot_meta[(size_t)ch * TILE + s].x) { out->consistent = 0; break; }
          }
        next_chunk = wd.chunk_end; next_row = wd.cam0 + wd.ncams;
      }
    }
    if (next_chunk != nn || next_row != L.win_cams.size()) out->consistent = 0;
  }
  if (L.cam_row_start[d->ncam] != L.win_cams.size() || L.win_dst.size() != L.win_cams.size()) out->consistent = 0;
  else {
    std::vector<uint8_t> seen(L.win_cams.size(), 0);
    std::vector<uint32_t> last(d->ncam, 0);
    for (size_t i = 0; i < L.win_cams.size() && out->consistent; ++i) {   // every row once, inside its camera's block, ascending with i
      const uint32_t cam = L.win_cams[i], row = L.win_dst[i];
      if (row < L.cam_row_start[cam] || row >= L.cam_row_start[cam + 1] || seen[row]++ || (last[cam] && row < last[cam])) out->consistent = 0;
      last[cam] = row + 1;
    }
  }
  return APEX_OK;
}

// Camera-pair blocks of the explicit reduced camera system (K7). S_(ci,cj) -= sum over the landmarks seen by both cameras of
// (Jc_i^T Jp_i) Hpp^-1 (Jp_j^T Jc_j): the (observation i, observation j) pairs of one landmark with cam_j <= cam_i are grouped by
// block here, once per upload and only when an explicit solve asks for it, so that the kernel adds each block's pairs in
// registers, in a fixed order, and writes it once - instead of one FP64 reduction per element and pair. Parallel over cameras:
// camera ci's pairs are collected through its slots and sorted (stably) by cam_j.
apex_status build_schur_pairs(Ctx& c) {
  if (c.pairs_ready) return APEX_OK;
  if (!c.staging) { c.err = "no layout"; return APEX_ERR_INVALID_STATE; }
  HostLayout& L = *static_cast<HostLayout*>(c.staging.get());
  const size_t nslots = c.nslots;
  const uint32_t ncam = c.ncam;
  std::vector<uint32_t> slot_lpg(nslots, 0xFFFFFFFFu);
  for (const TileDesc& t : L.tiles)
    for (uint32_t ch = 0; ch < t.nchunks; ++ch)
      for (int lane = 0; lane < TILE; ++lane) {
        const size_t slot = ((size_t)t.chunk0 + ch) * TILE + lane;
        if (L.slot_cam[slot] != PAD_CAM) slot_lpg[slot] = t.pt0 + L.slot_lp[slot];
      }
  std::vector<uint32_t> cam_start((size_t)ncam + 1, 0), cam_slots;
  for (size_t sl = 0; sl < nslots; ++sl) if (L.slot_cam[sl] != PAD_CAM) cam_start[L.slot_cam[sl] + 1]++;
  for (uint32_t k = 0; k < ncam; ++k) cam_start[k + 1] += cam_start[k];
  cam_slots.resize(cam_start[ncam]);
  {
    std::vector<uint32_t> cur(cam_start.begin(), cam_start.end() - 1);
    for (size_t sl = 0; sl < nslots; ++sl) if (L.slot_cam[sl] != PAD_CAM) cam_slots[cur[L.slot_cam[sl]]++] = (uint32_t)sl;
  }
  std::vector<uint64_t> pair_off((size_t)ncam + 1, 0);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t ci = 0; ci < (int64_t)ncam; ++ci) {
    uint64_t np = 0;
    for (uint32_t e = cam_start[ci]; e < cam_start[ci + 1]; ++e) {
      const uint32_t lp = slot_lpg[cam_slots[e]];
      for (uint32_t sj = L.pt_slot0[lp]; sj < L.pt_slot0[lp] + L.pt_cnt[lp]; ++sj) np += L.slot_cam[sj] <= (uint32_t)ci;
    }
    pair_off[ci + 1] = np;
  }
  for (uint32_t k = 0; k < ncam; ++k) pair_off[k + 1] += pair_off[k];
  const uint64_t npairs = pair_off[ncam];
  if (npairs > 0xFFFFFFF0ull) { c.err = "too many observation pairs for the explicit Schur complement"; return APEX_ERR_UNSUPPORTED; }
  HostVec<uint2> pairs(npairs);
  std::vector<std::vector<PairBlock>> blk(ncam);
#pragma omp parallel
  {
    struct Ent { uint32_t cj, si, sj; };
    std::vector<Ent> tmp;
#pragma omp for schedule(dynamic, 16)
    for (int64_t ci = 0; ci < (int64_t)ncam; ++ci) {
      tmp.clear();
      for (uint32_t e = cam_start[ci]; e < cam_start[ci + 1]; ++e) {
        const uint32_t si = cam_slots[e], lp = slot_lpg[si];
        for (uint32_t sj = L.pt_slot0[lp]; sj < L.pt_slot0[lp] + L.pt_cnt[lp]; ++sj)
          if (L.slot_cam[sj] <= (uint32_t)ci) tmp.push_back({L.slot_cam[sj], si, sj});
      }
      std::stable_sort(tmp.begin(), tmp.end(), [](const Ent& a, const Ent& b) { return a.cj < b.cj; });
      const uint64_t off = pair_off[ci];
      for (size_t k = 0; k < tmp.size(); ++k) {
        pairs[off + k] = make_uint2(tmp[k].si, tmp[k].sj);
        if (k == 0 || tmp[k].cj != tmp[k - 1].cj) {
          if (!blk[ci].empty()) blk[ci].back().end = (uint32_t)(off + k);
          blk[ci].push_back(PairBlock{(uint32_t)ci, tmp[k].cj, (uint32_t)(off + k), 0});
        }
      }
      if (!blk[ci].empty()) blk[ci].back().end = (uint32_t)(off + tmp.size());
    }
  }
  std::vector<PairBlock> blocks;
  for (uint32_t k = 0; k < ncam; ++k) blocks.insert(blocks.end(), blk[k].begin(), blk[k].end());
  c.npair_blocks = (uint32_t)blocks.size();
  cudaStream_t s = c.stream;
  APEX_CUDA_TRY(c, upload_vec(c.pair_blocks, blocks, s));
  APEX_CUDA_TRY(c, upload_vec(c.pair_slots, pairs, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_lpg, slot_lpg, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_cs8, c.slot_pos, s));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));   // the host vectors die with this scope
  c.pairs_ready = true;
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
  c.pairs_ready = false;
  c.linearized = false;
  c.have_step = false;
  if (c.pcg_graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)c.pcg_graph_exec); c.pcg_graph_exec = nullptr; }
  c.model = d->camera_model; c.K = K; c.opt = d->opt_flags;
  c.opt_intr = (d->opt_flags & APEX_OPT_INTRINSIC) != 0;
  c.shared_intr = c.opt_intr && (d->opt_flags & APEX_OPT_SHARED_INTRINSICS) != 0;
  c.intr_vars = d->intr_vars_present != 0;
  c.dc = 6 + (c.opt_intr ? K : 0);
  c.np = 2 * (c.dc + 3);
  c.ncam = d->ncam; c.npts = d->npts; c.nobs = d->nobs;
  c.cam_dof_ref = (uint64_t)c.ncam * (6 + ((c.opt_intr || c.intr_vars) ? K : 0));
  c.loss_id = d->loss_id;
  for (int i = 0; i < 4; ++i) c.loss_p[i] = d->loss_params[i];
  c.per_obs_loss = d->obs_loss != nullptr;

  if (!c.staging) {
    auto hl = std::make_shared<HostLayout>();
    // Pageable by default: page-locking the ~300 MB of staging arrays costs more (130 ms of cudaHostAlloc on the Venice shape) than
    // it returns on the first upload of a context (H2D of pageable memory: +20 ms) - measured end to end 0.76 s -> 0.54 s for
    // upload + 10 LM iterations + download. APEX_PINNED_STAGING=1 page-locks them for callers that upload to one context many times.
    { const char* e = getenv("APEX_PINNED_STAGING"); hl->set_pinned(e && atoi(e) != 0); }
    c.staging = hl;
  }
  HostLayout& L = *static_cast<HostLayout*>(c.staging.get());
  L.reset();
  APEX_TRY(schur_configure(c));   // window width, CTAs per SM of the chunk kernel -> number of ranges
  // The device buffers whose sizes follow from the chunk count are allocated by a helper thread WHILE the rest of the layout is built
  // (on a context that has never seen a problem cudaMalloc of ~1.6 GB costs 6-50 ms, sometimes more; the layout another ~25 ms from
  // that point on). On a context that already holds large enough buffers every alloc() is a no-op.
  std::thread alloc_thread;
  cudaError_t alloc_err = cudaSuccess;
  auto on_tiles = [&](uint32_t nchunks, uint32_t nnormal, uint64_t nobs_local, uint32_t npl) {
    alloc_thread = std::thread([&c, &alloc_err, nchunks, nnormal, nobs_local, npl] {
      auto A = [&](cudaError_t e) { if (alloc_err == cudaSuccess && e != cudaSuccess) alloc_err = e; };
      A(cudaSetDevice(c.device));
      const size_t nslots = (size_t)nchunks * TILE;
      A(c.J.alloc((size_t)nchunks * c.np * TILE));
      A(c.R.alloc((size_t)nchunks * 2 * TILE));
      A(c.slot_cam.alloc(nslots));
      A(c.slot_lp.alloc(nslots));
      A(c.slot_uv.alloc(nslots * 2));
      A(c.cslot_meta.alloc((size_t)nnormal * TILE));
      A(c.cslot_widx.alloc((size_t)nnormal * TILE));
      A(c.cm_uv.alloc(2 * (size_t)nobs_local));
      A(c.cm_lp.alloc(nobs_local));
      A(c.pt.alloc((size_t)npl * 3));
      A(c.hpp.alloc((size_t)npl * 6));
      A(c.gp.alloc((size_t)npl * 3));
      A(c.hinv.alloc((size_t)npl * 6));
      A(c.step_pt.alloc((size_t)npl * 3));
    });
  };
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{alloc_thread};
  build_layout(d, c.nranks, c.rank, L, c.mv_ctas_per_sm * (uint32_t)c.num_sms, c.mv_window, true, on_tiles);
  if (alloc_thread.joinable()) alloc_thread.join();
  if (alloc_err != cudaSuccess) { c.err = std::string("device allocation: ") + cudaGetErrorString(alloc_err); cudaGetLastError(); return APEX_ERR_CUDA; }
  c.shard = L.shard; c.npl = L.npl; c.nobs_local = L.nobs_local;
  c.nnormal_chunks = L.nnormal_chunks; c.nchunks = L.nchunks;
  c.mv_nranges = (uint32_t)L.range_win0.size() - 1; c.mv_nwindows = (uint32_t)L.win_desc.size(); c.mv_nrows = L.win_cams.size();
  {
    // Deterministic flush (per-window partial rows + a fixed-order second pass) unless its extra traffic (rows written and
    // read once per operator application) is more than a tenth of what the operator streams anyway, i.e. unless the
    // windows are short (a landmark order without camera locality); APEX_DETERMINISTIC=0|1 overrides.
    const double extra = 2.0 * (double)c.mv_nrows * c.dc * 8.0, stream = (double)c.nobs_local * (8.0 * c.np + 8.0);
    c.mv_det = extra <= 0.1 * stream;
    if (const char* e = getenv("APEX_DETERMINISTIC")) c.mv_det = atoi(e) != 0;
  }
  c.ntiles = (uint32_t)L.tiles.size(); c.ngiant = (uint32_t)L.giant_tiles.size(); c.nitems = (uint32_t)L.items.size();
  c.nslots = (size_t)L.nchunks * TILE;
  c.slot_obs.swap(L.slot_obs);
  c.slot_pos.swap(L.slot_pos);
  c.h_pt_cnt = L.pt_cnt;
  lap("build_layout");

  // ---- fixed masks ----
  std::vector<uint8_t> pose_fixed(c.ncam, 0), pt_fixed(c.npl, 0);
  std::vector<uint16_t> intr_fixed(c.ncam, 0);
  if (d->pose_fixed) std::copy(d->pose_fixed, d->pose_fixed + c.ncam, pose_fixed.begin());
  if (d->intr_fixed) std::copy(d->intr_fixed, d->intr_fixed + c.ncam, intr_fixed.begin());
  if (c.shared_intr) std::fill(intr_fixed.begin(), intr_fixed.end(), intr_fixed[0]);   // one variable, one mask
  if (d->pt_fixed) for (uint32_t lp = 0; lp < c.npl; ++lp) pt_fixed[lp] = d->pt_fixed[c.shard.to_global(lp)];

  // ---- to the device ----
  cudaStream_t s = c.stream;
  g_h2d_bytes = 0;
  g_bounce_ctx = &c;
  struct BounceOff { ~BounceOff() { g_bounce_ctx = nullptr; } } bounce_off;
  APEX_CUDA_TRY(c, upload_vec(c.tiles, L.tiles, s));
  APEX_CUDA_TRY(c, upload_vec(c.giant_tiles, L.giant_tiles, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_cam, L.slot_cam, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_lp, L.slot_lp, s));
  APEX_CUDA_TRY(c, upload_vec(c.slot_uv, L.slot_uv, s));
  if (c.per_obs_loss) {
    std::vector<LossSpecPod> tab(d->n_losses);
    for (int i = 0; i < d->n_losses; ++i) tab[i] = LossSpecPod{d->loss_table[i].loss_id, d->loss_table[i].params[0], d->loss_table[i].params[1]};
    APEX_CUDA_TRY(c, upload_vec(c.slot_loss, L.slot_loss, s));
    APEX_CUDA_TRY(c, upload_vec(c.cm_loss, L.cm_loss, s));
    APEX_CUDA_TRY(c, upload_vec(c.loss_tab, tab, s));
    APEX_CUDA_TRY(c, cudaStreamSynchronize(s));   // `tab` dies with this scope
  }
  APEX_CUDA_TRY(c, upload_vec(c.pt_slot0, L.pt_slot0, s));
  APEX_CUDA_TRY(c, upload_vec(c.pt_cnt, L.pt_cnt, s));
  APEX_CUDA_TRY(c, upload_vec(c.chunk_desc, L.chunk_desc, s));
  APEX_CUDA_TRY(c, upload_vec(c.cslot_meta, L.cslot_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.cpt_meta, L.cpt_meta, s));
  APEX_CUDA_TRY(c, upload_vec(c.cslot_widx, L.cslot_widx, s));
  APEX_CUDA_TRY(c, upload_vec(c.win_desc, L.win_desc, s));
  APEX_CUDA_TRY(c, upload_vec(c.range_win0, L.range_win0, s));
  APEX_CUDA_TRY(c, upload_vec(c.win_cams, L.win_cams, s));
  if (c.mv_det) {  // only the deterministic flush reads the per-camera row lists
    APEX_CUDA_TRY(c, upload_vec(c.cam_row_start, L.cam_row_start, s));
    APEX_CUDA_TRY(c, upload_vec(c.win_dst, L.win_dst, s));
    APEX_CUDA_TRY(c, c.det_partial.alloc((size_t)std::max<uint64_t>(c.mv_nrows, 1) * c.dc));
  }
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
  APEX_CUDA_TRY(c, c.xpad.alloc((size_t)c.ncam * xpad_stride(c.dc)));
  APEX_CUDA_TRY(c, c.step_cam.alloc(ncd));
  APEX_CUDA_TRY(c, c.step_pt.alloc((size_t)c.npl * 3));
  APEX_CUDA_TRY(c, c.red_scratch.alloc(8 * (size_t)std::max<uint32_t>(std::max<uint32_t>(c.nchunks, c.ncam), 1024u) + 64));
  APEX_CUDA_TRY(c, cudaMemsetAsync(c.step_cam.p, 0, ncd * sizeof(double), s));
  APEX_CUDA_TRY(c, cudaMemsetAsync(c.step_pt.p, 0, std::max<size_t>((size_t)c.npl * 3, 1) * sizeof(double), s));

  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.pose.p, d->pose, (size_t)c.ncam * 7 * sizeof(double), cudaMemcpyHostToDevice, s));
  std::vector<double> intr_rep;   // shared intrinsics: row 0 is the value of every camera's copy (lives until the sync below)
  if (c.shared_intr) {
    intr_rep.resize((size_t)c.ncam * K);
    for (uint32_t cam = 0; cam < c.ncam; ++cam) std::copy(d->intr, d->intr + K, intr_rep.begin() + (size_t)cam * K);
  }
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.intr.p, c.shared_intr ? intr_rep.data() : d->intr, (size_t)c.ncam * K * sizeof(double), cudaMemcpyHostToDevice, s));
  std::vector<double> pt_local;  // one rank owns every landmark in the caller's order: no gather needed
  if (c.nranks > 1) pt_local = gather_local_points(c, d->pt);
  if (c.npl) APEX_CUDA_TRY(c, bounce_copy(c.pt.p, c.nranks > 1 ? pt_local.data() : d->pt, (size_t)c.npl * 3 * sizeof(double), s));
  c.upload_h2d_bytes = g_h2d_bytes + ((size_t)c.ncam * (7 + K) + (size_t)c.npl * 3) * sizeof(double);
  lap("allocations + parameters");
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));  // the host vectors above die with this scope
  lap("stream sync");
  // the peer-memory buffers of the fused all-reduce were mapped at context creation; only a larger camera vector re-maps them
  // (collective: every rank sees the same ncd and the same capacity, so all take the same branch)
  if (c.nranks > 1 && !(c.p2p_ok && ncd <= c.ar_n)) APEX_TRY(setup_peer_allreduce(c, std::max(ncd, AR_CAPACITY)));
  lap("peer setup");
  c.have_problem = true;
  return APEX_OK;
}

}  // namespace apex
