// explicit.cu — K7 formation of the reduced camera system S and K8 its dense FP64 solve.
//
// sm_100a equivalents of SparseSchurComplementSolver (src/linalg/sparse/explicit_schur.rs):
//   compute_schur_complement :771-925  -> schur_form_kernel (+ diag / symmetrize_drop kernels)
//   solve_with_cholesky      :539-634  -> blocked right-looking Cholesky; the trailing update is the only
//                                          dense contraction of the path and runs on the FP64 tensor pipe
//                                          (mma.sync.m8n8k4.f64 = DMMA.8x8x4; tcgen05 has no FP64 kind)
//   solve_with_pcg           :639-756  -> scalar-Jacobi PCG on the dense S (what SchurVariant::Iterative
//                                          dispatches to today, explicit_schur.rs:1224-1225)
// S is stored dense, row-major, in the camera-major local layout (row = cam*dc + dof), padded to a multiple
// of 64 with an identity tail so no kernel needs edge handling.
#include <algorithm>
#include <cmath>

#include "apex_ctx.h"
#include "ba_device.cuh"
#include "kernels_common.cuh"

namespace apex {

constexpr int NB = 64;        // Cholesky block size
constexpr int PLD = NB + 4;   // padded leading dimension of the shared-memory panels (bank-conflict free)

// ----------------------------------------------------------------------------------------------------
// K7: S -= sum_p sum_{i,j in obs(p)} (Jc_i^T Jp_i) Hpp_p^-1 (Jp_j^T Jc_j), lower block triangle
// ----------------------------------------------------------------------------------------------------
struct FormArgs {
  const TileDesc* tiles;
  const uint32_t* slot_cam;
  const uint16_t* slot_lp;
  const uint32_t* pt_slot0;
  const uint32_t* pt_cnt;
  const uint2* cslot_meta;   // normal chunks: camera-sorted lane of each point-major lane (split slot order, apex_ctx.h)
  const double* J;
  const double* hinv;
  double* S;
  size_t ld;
  uint32_t npl;
};

template <int DC>
__global__ void __launch_bounds__(TILE) schur_form_kernel(FormArgs a) {
  constexpr int NP = 2 * (DC + 3);
  const TileDesc td = a.tiles[blockIdx.x];
  const int tid = threadIdx.x;
  for (uint32_t ch = 0; ch < td.nchunks; ++ch) {
    const size_t chunk = (size_t)td.chunk0 + ch;
    const size_t slot = chunk * TILE + tid;
    const uint32_t cam_i = a.slot_cam[slot];
    if (cam_i == PAD_CAM) continue;
    const uint32_t lp = td.pt0 + a.slot_lp[slot];
    const uint32_t s0 = a.pt_slot0[lp], cnt = a.pt_cnt[lp];
    const bool split = td.nchunks == 1;  // camera half of a normal chunk sits at the observation's camera-sorted lane
    double jall[NP];
    {
      const double2* Ji = reinterpret_cast<const double2*>(a.J) + chunk * (NP / 2) * TILE;
      const int clane = split ? (int)((a.cslot_meta[slot].y >> 8) & 0xFFu) : tid;
#pragma unroll
      for (int m = 0; m < DC; ++m) { const double2 v = Ji[(size_t)m * TILE + clane]; jall[2 * m] = v.x; jall[2 * m + 1] = v.y; }
#pragma unroll
      for (int m = DC; m < DC + 3; ++m) { const double2 v = Ji[(size_t)m * TILE + tid]; jall[2 * m] = v.x; jall[2 * m + 1] = v.y; }
    }
    const double* jc = jall;
    const double* jp = jall + 2 * DC;
    const size_t n = a.npl;
    const double h00 = a.hinv[0 * n + lp], h01 = a.hinv[1 * n + lp], h02 = a.hinv[2 * n + lp];
    const double h11 = a.hinv[3 * n + lp], h12 = a.hinv[4 * n + lp], h22 = a.hinv[5 * n + lp];
    double G[2][3];  // Jp_i Hpp^-1
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      G[r][0] = jp[r * 3] * h00 + jp[r * 3 + 1] * h01 + jp[r * 3 + 2] * h02;
      G[r][1] = jp[r * 3] * h01 + jp[r * 3 + 1] * h11 + jp[r * 3 + 2] * h12;
      G[r][2] = jp[r * 3] * h02 + jp[r * 3 + 1] * h12 + jp[r * 3 + 2] * h22;
    }
    for (uint32_t j = s0; j < s0 + cnt; ++j) {
      const uint32_t cam_j = a.slot_cam[j];
      if (cam_j > cam_i) continue;
      const double2* Jch = reinterpret_cast<const double2*>(a.J) + (size_t)(j / TILE) * (NP / 2) * TILE;
      const double2* Jj = Jch + (j % TILE);
      const double2* Jcj = Jch + (split ? ((a.cslot_meta[j].y >> 8) & 0xFFu) : (j % TILE));
      double pj[6];
#pragma unroll
      for (int m = 0; m < 3; ++m) { const double2 v = Jj[(size_t)(DC + m) * TILE]; pj[2 * m] = v.x; pj[2 * m + 1] = v.y; }
      double M[2][2];  // Jp_i Hpp^-1 Jp_j^T
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) M[r][cc] = G[r][0] * pj[cc * 3] + G[r][1] * pj[cc * 3 + 1] + G[r][2] * pj[cc * 3 + 2];
      double cj[2 * DC];
#pragma unroll
      for (int m = 0; m < DC; ++m) { const double2 v = Jcj[(size_t)m * TILE]; cj[2 * m] = v.x; cj[2 * m + 1] = v.y; }
      double* Srow = a.S + ((size_t)cam_i * DC) * a.ld + (size_t)cam_j * DC;
#pragma unroll
      for (int p = 0; p < DC; ++p) {
        const double a0 = jc[p] * M[0][0] + jc[DC + p] * M[1][0];
        const double a1 = jc[p] * M[0][1] + jc[DC + p] * M[1][1];
#pragma unroll
        for (int q = 0; q < DC; ++q) red_add(Srow + (size_t)p * a.ld + q, -fma(a0, cj[q], a1 * cj[DC + q]));
      }
    }
  }
}

// K7 by camera-pair blocks (build_schur_pairs, problem.cu): one warp per block (cam_i >= cam_j) of the lower block triangle adds the
// block's (observation i, observation j) pairs in their fixed order and stores the dc x dc block once - no reductions into memory,
// bitwise reproducible. Lane l = g*DC + q owns column q and the rows p = g, g + G, ... (G = 32 / DC groups of lanes); per pair every lane
// recomputes the 2x2 middle  Jp_i Hpp^-1 Jp_j^T  (broadcast loads), reads its two entries of Jc_j and the rows' entries of Jc_i.
// The pairs of a block name observations anywhere in the problem: read from the chunk planes, the 2*dc + 6 doubles of one observation
// sit in dc + 3 different 4 KB-apart planes, i.e. one 32-byte sector per 16 useful bytes and 2 x (dc + 3) sectors per pair (ncu on the
// Kannala-Brandt 2 000-camera shape: 44.8 GB of DRAM reads for 1.7 GB of Jacobians, long scoreboard 73 % of the stall samples). So the
// explicit solve first packs the linearisation observation-major - row `slot` = [camera half 2*dc | landmark half 6], contiguous -
// and a pair reads two rows: (2*dc + 6) / 4 sectors each.
template <int DC>
__global__ void __launch_bounds__(TILE) pack_obs_major_kernel(const double* __restrict__ J, const uint8_t* __restrict__ slot_cs8, double* __restrict__ rows) {
  constexpr int NPAIR = DC + 3, RS = 2 * DC + 6;
  const size_t chunk = blockIdx.x, slot = chunk * TILE + threadIdx.x;
  const uint32_t cs = slot_cs8[slot];   // lane of the camera half inside the chunk (split slot order)
  const double2* J2 = reinterpret_cast<const double2*>(J) + chunk * NPAIR * TILE;
  double2* row = reinterpret_cast<double2*>(rows + slot * RS);
#pragma unroll
  for (int m = 0; m < DC; ++m) row[m] = ld_stream2(J2 + (size_t)m * TILE + cs);
#pragma unroll
  for (int m = 0; m < 3; ++m) row[DC + m] = ld_stream2(J2 + (size_t)(DC + m) * TILE + threadIdx.x);
}

struct PairArgs {
  const PairBlock* blocks;
  const uint2* pairs;
  const uint32_t* slot_lpg;
  const double* rows;    // observation-major linearisation (pack_obs_major_kernel)
  const double* hinv;
  double* S;
  size_t ld;
  uint32_t npl, nblocks;
};

template <int DC>
__global__ void __launch_bounds__(256, 4) schur_form_pairs_kernel(PairArgs a) {
  constexpr int G = 32 / DC, RPL = (DC + G - 1) / G;
  const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= a.nblocks) return;
  const int lane = threadIdx.x & 31, g = lane / DC, q = lane % DC;
  const bool act = g < G;
  const PairBlock blk = a.blocks[b];
  double acc[RPL];
#pragma unroll
  for (int r = 0; r < RPL; ++r) acc[r] = 0.0;
  constexpr int RS = 2 * DC + 6;
  const size_t n = a.npl;
  // the pair's indices (pair -> slots -> landmark) are fetched one pair ahead: one of the two dependent memory round trips per
  // pair is then off the critical path
  uint32_t nsi = 0, nsj = 0, nlp = 0;
  auto fetch = [&](uint32_t e) {
    const uint2 pr = __ldg(a.pairs + e);
    nsi = pr.x; nsj = pr.y;
    nlp = __ldg(a.slot_lpg + nsi);
  };
  if (blk.begin < blk.end) fetch(blk.begin);
  for (uint32_t e = blk.begin; e < blk.end; ++e) {
    const uint32_t si = nsi, sj = nsj, lp = nlp;
    if (e + 1 < blk.end) fetch(e + 1);
    const double* Ri = a.rows + (size_t)si * RS, *Rj = a.rows + (size_t)sj * RS;
    double pi[6], pj[6];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const double2 u = __ldg(reinterpret_cast<const double2*>(Ri + 2 * DC) + m), v = __ldg(reinterpret_cast<const double2*>(Rj + 2 * DC) + m);
      pi[2 * m] = u.x; pi[2 * m + 1] = u.y; pj[2 * m] = v.x; pj[2 * m + 1] = v.y;
    }
    const double h00 = __ldg(a.hinv + 0 * n + lp), h01 = __ldg(a.hinv + 1 * n + lp), h02 = __ldg(a.hinv + 2 * n + lp);
    const double h11 = __ldg(a.hinv + 3 * n + lp), h12 = __ldg(a.hinv + 4 * n + lp), h22 = __ldg(a.hinv + 5 * n + lp);
    double M[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double g0 = pi[r * 3] * h00 + pi[r * 3 + 1] * h01 + pi[r * 3 + 2] * h02;
      const double g1 = pi[r * 3] * h01 + pi[r * 3 + 1] * h11 + pi[r * 3 + 2] * h12;
      const double g2 = pi[r * 3] * h02 + pi[r * 3 + 1] * h12 + pi[r * 3 + 2] * h22;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) M[r][cc] = g0 * pj[cc * 3] + g1 * pj[cc * 3 + 1] + g2 * pj[cc * 3 + 2];
    }
    if (!act) continue;
    // element k of a camera half: row 0 = 0..DC-1, row 1 = DC..2DC-1
    const double cj0 = __ldg(Rj + q), cj1 = __ldg(Rj + DC + q);
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      const int p = g + r * G;
      if (p < DC) {
        const double ci0 = __ldg(Ri + p), ci1 = __ldg(Ri + DC + p);
        const double a0 = ci0 * M[0][0] + ci1 * M[1][0], a1 = ci0 * M[0][1] + ci1 * M[1][1];
        acc[r] -= fma(a0, cj0, a1 * cj1);
      }
    }
  }
  if (act) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      const int p = g + r * G;
      if (p < DC) a.S[((size_t)blk.cam_i * DC + p) * a.ld + (size_t)blk.cam_j * DC + q] = acc[r];
    }
  }
}

// S diagonal blocks += H_cc + lambda I (explicit_schur.rs:1186-1205, :784-792); identity on the padded tail
__global__ void schur_diag_kernel(double* S, size_t ld, const double* hcc, const DevState* st, uint32_t ncam, int dc, uint32_t n, uint32_t npad,
                                  int add_hcc) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nblk = (size_t)ncam * dc * dc;
  if (gid < nblk) {
    if (!add_hcc) return;
    const uint32_t cam = (uint32_t)(gid / (dc * dc));
    const int a = (int)((gid / dc) % dc), b = (int)(gid % dc);
    S[((size_t)cam * dc + a) * ld + (size_t)cam * dc + b] += hcc[gid] + (a == b ? st->damping : 0.0);
  } else if (gid < nblk + (npad - n)) {
    const size_t i = n + (gid - nblk);
    S[i * ld + i] = add_hcc ? 1.0 : 0.0;
  }
}

// explicit_schur.rs:903-921: symmetrise and drop |S_ij| <= 1e-12. Only the lower triangle was accumulated
// (the reference accumulates both and averages them); mirror it into the upper triangle.
__global__ void symmetrize_drop_kernel(double* S, size_t ld, uint32_t nt) {
  __shared__ double tile[32][33];
  // linear index -> (bi >= bj)
  const uint32_t t = blockIdx.x;
  uint32_t bi = (uint32_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((uint64_t)bi * (bi + 1) / 2 > t) --bi;
  while ((uint64_t)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const uint32_t bj = t - (uint32_t)((uint64_t)bi * (bi + 1) / 2);
  (void)nt;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const size_t row = (size_t)bi * 32 + r, col = (size_t)bj * 32 + tx;
    double v = S[row * ld + col];
    if (bi == bj && tx > r) v = 0.0;  // upper part of a diagonal tile: filled from the mirror below
    if (fabs(v) <= 1e-12) v = 0.0;
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const size_t row = (size_t)bi * 32 + r, col = (size_t)bj * 32 + tx;
    double v = tile[r][tx];
    if (bi == bj && tx > r) v = tile[tx][r];
    S[row * ld + col] = v;
    if (bi != bj) {
      const size_t mrow = (size_t)bj * 32 + r, mcol = (size_t)bi * 32 + tx;
      S[mrow * ld + mcol] = tile[tx][r];
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// K8: blocked right-looking Cholesky, lower, in place
// ----------------------------------------------------------------------------------------------------
// factor the NB x NB diagonal block at k0 (one CTA, 256 threads as 16 x 16; thread (ty, tx) owns the 4 x 4 elements
// rows ty + 16 a, columns tx + 16 b of the block, so the trailing update of a column needs no index arithmetic and the
// rows / columns still alive are spread over all threads). Two barriers per column: every thread takes the square root
// of the pivot itself, the pivot's owner stores it, the threads of the column scale their entry.
__global__ void __launch_bounds__(256) chol_potrf_kernel(double* A, size_t ld, uint32_t k0, DevState* st) {
  __shared__ double a[NB][NB + 1];
  __shared__ int bad;
  if (st->chol_fail) return;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  if (tid == 0) bad = 0;
  for (int e = tid; e < NB * NB; e += 256) { const int r = e >> 6, cidx = e & 63; a[r][cidx] = A[((size_t)k0 + r) * ld + k0 + cidx]; }
  __syncthreads();
  for (int j = 0; j < NB; ++j) {
    double d = a[j][j];
    const bool ok = d > 0.0;
    if (!ok) d = 1.0;
    const double rs = rsqrt(d);   // the column is serial in (pivot -> scale -> update): 1/sqrt + multiplies instead of sqrt + divisions
    __syncthreads();  // everybody has read the pivot
    if (tid == j) { a[j][j] = d * rs; if (!ok) bad = j + 1; }
    else if (tid > j && tid < NB) a[tid][j] *= rs;
    __syncthreads();
    // trailing update of the lower triangle: rows i > j, columns j < k <= i
    double cj[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) cj[b] = a[tx + 16 * b][j];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ty + 16 * r;
      if (i <= j) continue;
      const double ri = a[i][j];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int k = tx + 16 * b;
        if (k > j && k <= i) a[i][k] -= ri * cj[b];
      }
    }
    // no barrier here: the next column's pivot read is behind the barrier at the top of the loop only after all updates
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) { const int r = e >> 6, cidx = e & 63; if (cidx <= r) A[((size_t)k0 + r) * ld + k0 + cidx] = a[r][cidx]; }
  if (tid == 0 && bad) atomicCAS(&st->chol_fail, 0, (int)k0 + bad);
}

// panel: rows [r0 + 64*blockIdx.x, +64): X <- X * L11^-T by forward substitution, four threads per row (lanes 4r..4r+3 of
// a warp): each takes every fourth term of the dot product, two shuffles add the partial sums
__global__ void __launch_bounds__(256) chol_trsm_kernel(double* A, size_t ld, uint32_t k0) {
  extern __shared__ double trsm_smem[];
  double (*l)[PLD] = reinterpret_cast<double (*)[PLD]>(trsm_smem);
  double (*x)[PLD] = reinterpret_cast<double (*)[PLD]>(trsm_smem + NB * PLD);
  const int tid = threadIdx.x;
  const size_t r0 = (size_t)k0 + NB + (size_t)blockIdx.x * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e >> 6, cidx = e & 63;
    l[r][cidx] = A[((size_t)k0 + r) * ld + k0 + cidx];
    x[r][cidx] = A[(r0 + r) * ld + k0 + cidx];
  }
  __syncthreads();
  if (tid < NB) l[tid][tid] = 1.0 / l[tid][tid];  // reciprocal pivots once, off the serial path
  __syncthreads();
  const int row = tid >> 2, q = tid & 3;
  for (int j = 0; j < NB; ++j) {
    double s = 0.0;
    for (int k = q; k < j; k += 4) s = fma(x[row][k], l[j][k], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (q == 0) x[row][j] = (x[row][j] - s) * l[j][j];
    __syncwarp();
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) { const int r = e >> 6, cidx = e & 63; A[(r0 + r) * ld + k0 + cidx] = x[r][cidx]; }
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Trailing update  A[r0+.., c0+..] -= L[.., k0..k0+kdepth) L[.., k0..k0+kdepth)^T  on 64x64 tiles of the lower block
// triangle: tile (bi, bj) = rows r0+64*bi, columns r0+64*bj, bi >= bj, bj < ncol_tiles. 4 warps, each a 32x32 sub-tile =
// 4x4 DMMA.8x8x4 accumulators kept in registers across the whole k depth (the C tile is read and written once per
// call, which is what makes the two-level blocking of dense_cholesky pay); the two 64x64 panel tiles of each 64-wide
// k slice are staged in shared memory.
__global__ void __launch_bounds__(128) chol_syrk_kernel(double* A, size_t ld, uint32_t k0, uint32_t kdepth, uint32_t r0) {
  extern __shared__ double smem[];
  double* Pa = smem;
  double* Pb = smem + NB * PLD;
  const uint32_t bi = blockIdx.x, bj = blockIdx.y;
  if (bi < bj) return;
  const size_t ri = (size_t)r0 + (size_t)bi * NB, rj = (size_t)r0 + (size_t)bj * NB;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int lr = lane >> 2, lc = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
  for (uint32_t kc = 0; kc < kdepth; kc += NB) {
    if (kc) __syncthreads();
    for (int e = tid; e < NB * NB / 2; e += 128) {
      const int r = e / (NB / 2), c2 = (e % (NB / 2)) * 2;
      const double2 va = *reinterpret_cast<const double2*>(A + (ri + r) * ld + k0 + kc + c2);
      Pa[r * PLD + c2] = va.x; Pa[r * PLD + c2 + 1] = va.y;
      const double2 vb = *reinterpret_cast<const double2*>(A + (rj + r) * ld + k0 + kc + c2);
      Pb[r * PLD + c2] = vb.x; Pb[r * PLD + c2 + 1] = vb.y;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < NB; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) af[mi] = Pa[(wm + mi * 8 + lr) * PLD + kk + lc];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) bf[ni] = Pb[(wn + ni * 8 + lr) * PLD + kk + lc];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2* p = reinterpret_cast<double2*>(A + (ri + wm + mi * 8 + lr) * ld + rj + wn + ni * 8 + lc * 2);
      double2 v = *p;
      v.x -= acc[mi][ni][0];
      v.y -= acc[mi][ni][1];
      *p = v;
    }
}

// The deep (k = 256) trailing update on 128x128 tiles: 64x64 tiles move 64 KB through L2 per 64-deep slice for 0.5 MFLOP
// (8 flop/B - at the DMMA rate that is more than L2 delivers); a 128x128 tile doubles the intensity. 512 threads = 16
// warps as 4 x 4, each warp a 32x32 sub-tile = 4x4 DMMA.8x8x4 accumulators (four warps per scheduler: with two the tensor
// pipe idled a quarter of the time on fixed-latency waits); the two
// 128 x 32 panel slices of a k step are staged with cp.async (16-byte copies) into a three-stage shared-memory ring
// (stride 36 doubles: fragment loads hit 16 distinct 8-byte banks per half warp), so the loads of slice s+1 are in
// flight while slice s feeds the tensor pipe.
constexpr int SB = 128, SKC = 32, SPLD = SKC + 4, SYRK_STAGES = 3, SYRK_THREADS = 512;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// Tiles of the lower block triangle in column tiles [bj0, bj0 + ny) of a trailing matrix with `rem` row tiles, numbered column by
// column: column c holds the R - c tiles on and below the diagonal (R = rem - bj0), so tile t lives in the column c with
// c R - c (c - 1) / 2 <= t. (A rem x ny grid with an early return for the upper triangle launched nearly twice the CTAs, and with
// one 221 KB CTA per SM an empty CTA holds its SM for a launch + exit.)
__host__ __device__ inline uint32_t syrk128_tiles(uint32_t rem, uint32_t bj0, uint32_t ny) {
  const uint64_t R = rem - bj0;
  return (uint32_t)((uint64_t)ny * R - (uint64_t)ny * (ny - 1) / 2);
}
__global__ void __launch_bounds__(SYRK_THREADS, 1) chol_syrk128_kernel(double* A, size_t ld, uint32_t k0, uint32_t kdepth, uint32_t r0, uint32_t bj0, uint32_t rem) {
  extern __shared__ __align__(16) double smem[];
  uint32_t bi, bj;
  {
    const double R = (double)(rem - bj0), t = (double)blockIdx.x;
    int64_t c = (int64_t)floor(((2.0 * R + 1.0) - sqrt((2.0 * R + 1.0) * (2.0 * R + 1.0) - 8.0 * t)) * 0.5);
    const int64_t Ri = (int64_t)(rem - bj0), ti = (int64_t)blockIdx.x;
    if (c < 0) c = 0;
    while ((c + 1) * Ri - (c + 1) * c / 2 <= ti) ++c;
    while (c * Ri - c * (c - 1) / 2 > ti) --c;
    bj = bj0 + (uint32_t)c;
    bi = bj + (uint32_t)(ti - (c * Ri - c * (c - 1) / 2));
  }
  const size_t ri = (size_t)r0 + (size_t)bi * SB, rj = (size_t)r0 + (size_t)bj * SB;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;   // 16 warps as 4 x 4, each a 32 x 32 sub-tile = 4 x 4 DMMA accumulators
  const int lr = lane >> 2, lc = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
  const int nslices = (int)(kdepth / SKC);
  auto stage = [&](int sl) {  // slice sl -> buffer sl % 3: 2 panels x 128 rows x 32 doubles = 2 x 2048 16-byte copies (nothing past the end)
    if (sl < nslices) {
      double* Pa = smem + (size_t)(sl % SYRK_STAGES) * 2 * SB * SPLD;
      double* Pb = Pa + SB * SPLD;
      const uint32_t kc = (uint32_t)sl * SKC;
      for (int e = tid; e < SB * SKC / 2; e += SYRK_THREADS) {
        const int r = e / (SKC / 2), c2 = (e % (SKC / 2)) * 2;
        cp_async16(Pa + r * SPLD + c2, A + (ri + r) * ld + k0 + kc + c2);
        cp_async16(Pb + r * SPLD + c2, A + (rj + r) * ld + k0 + kc + c2);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one group per call, empty or not: the wait below counts groups
  };
  // Three-stage ring, ONE barrier per slice, armed in the MIDDLE of the slice's eight k steps: there every thread's copies of
  // slice s+1 have landed (wait_group 0; they were issued a slice ago); the barrier makes them visible to everybody and proves
  // that everybody is past slice s-1, whose buffer then takes slice s+2. The fragments of a k step are loaded one step ahead into
  // a second register set - across the slice boundary as well, which the mid-slice barrier allows - so no warp starts a step
  // with a shared-memory round trip in front of its first DMMA (with the barrier at the top of a slice all sixteen warps did so
  // at the same moment and the tensor pipe idled).
  const int fr = lr * SPLD + lc;
  double af[2][4], bf[2][4];
  auto load = [&](int buf, const double* Pa, const double* Pb, int kk) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) af[buf][mi] = Pa[(wm + mi * 8) * SPLD + fr + kk];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) bf[buf][ni] = Pb[(wn + ni * 8) * SPLD + fr + kk];
  };
  // The slice barrier is split (mbarrier, 512 arrivals per phase): a thread ARRIVES in the middle of slice s ("my copies of slice
  // s+1 have landed and I am past slice s-1") and WAITS only at the slice's last step, in front of its first read of slice s+1 and
  // of the copies that overwrite slice s-1's buffer: three k steps of slack, so the warps of a CTA drift against each other
  // instead of meeting eight times per 256 k.
  __shared__ __align__(8) unsigned long long slice_bar;
  const uint32_t sbar = (uint32_t)__cvta_generic_to_shared(&slice_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sbar), "r"((uint32_t)SYRK_THREADS) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  stage(0);
  stage(1);
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncthreads();
  load(0, smem, smem + SB * SPLD, 0);
  uint32_t phase = 0;
  for (int sidx = 0; sidx < nslices; ++sidx) {
    const double* Pa = smem + (size_t)(sidx % SYRK_STAGES) * 2 * SB * SPLD;
    const double* Pb = Pa + SB * SPLD;
    const double* Na = smem + (size_t)((sidx + 1) % SYRK_STAGES) * 2 * SB * SPLD;
    const double* Nb = Na + SB * SPLD;
#pragma unroll
    for (int st = 0; st < SKC / 4; ++st) {
      if (st == SKC / 8) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("{\n .reg .b64 t;\n mbarrier.arrive.shared::cta.b64 t, [%0];\n}" ::"r"(sbar) : "memory");
      }
      if (st + 1 < SKC / 4) load((st + 1) & 1, Pa, Pb, 4 * (st + 1));
      else {
        for (;;) {   // every thread of the CTA has arrived for this slice
          uint32_t ok;
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(sbar), "r"(phase) : "memory");
          if (ok) break;
        }
        phase ^= 1u;
        stage(sidx + 2);
        if (sidx + 1 < nslices) load((st + 1) & 1, Na, Nb, 0);
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], af[st & 1][mi], bf[st & 1][ni]);
    }
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2* p = reinterpret_cast<double2*>(A + (ri + wm + mi * 8 + lr) * ld + rj + wn + ni * 8 + lc * 2);
      double2 v = *p;
      v.x -= acc[mi][ni][0];
      v.y -= acc[mi][ni][1];
      *p = v;
    }
}

// triangular solves with the factor: forward L y = b, backward L^T x = y, 64-row blocks
__global__ void __launch_bounds__(NB) trsv_diag_kernel(const double* A, size_t ld, uint32_t k0, double* v, int transpose) {
  __shared__ double l[NB][NB + 1];
  __shared__ double y[NB];
  const int tid = threadIdx.x;
  for (int r = 0; r < NB; ++r) l[r][tid] = A[((size_t)k0 + r) * ld + k0 + tid];
  y[tid] = v[k0 + tid];
  __syncthreads();
  if (!transpose) {
    for (int j = 0; j < NB; ++j) {
      if (tid == j) y[j] /= l[j][j];
      __syncthreads();
      if (tid > j) y[tid] -= l[tid][j] * y[j];
      __syncthreads();
    }
  } else {
    for (int j = NB - 1; j >= 0; --j) {
      if (tid == j) y[j] /= l[j][j];
      __syncthreads();
      if (tid < j) y[tid] -= l[j][tid] * y[j];
      __syncthreads();
    }
  }
  v[k0 + tid] = y[tid];
}

// forward: v[i] -= sum_c L[i][k0+c] y[k0+c] for rows i >= k0+64 ; 8 threads per row
__global__ void __launch_bounds__(256) trsv_update_fwd_kernel(const double* A, size_t ld, uint32_t k0, double* v, uint32_t npad) {
  __shared__ double y[NB];
  const int tid = threadIdx.x;
  if (tid < NB) y[tid] = v[k0 + tid];
  __syncthreads();
  const size_t row = (size_t)k0 + NB + (size_t)blockIdx.x * 32 + (tid >> 3);
  const int sub = tid & 7;
  double s = 0.0;
  if (row < npad) {
    const double* Lr = A + row * ld + k0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s = fma(Lr[sub + 8 * c], y[sub + 8 * c], s);
  }
  s += __shfl_down_sync(0xffffffffu, s, 4, 8);
  s += __shfl_down_sync(0xffffffffu, s, 2, 8);
  s += __shfl_down_sync(0xffffffffu, s, 1, 8);
  if (sub == 0 && row < npad) v[row] -= s;
}

// backward: v[j] -= sum_r L[k0+r][j] x[k0+r] for columns j < k0 ; one thread per column
__global__ void __launch_bounds__(256) trsv_update_bwd_kernel(const double* A, size_t ld, uint32_t k0, double* v) {
  __shared__ double x[NB];
  const int tid = threadIdx.x;
  if (tid < NB) x[tid] = v[k0 + tid];
  __syncthreads();
  const size_t col = (size_t)blockIdx.x * 256 + tid;
  if (col >= k0) return;
  double s = 0.0;
#pragma unroll 8
  for (int r = 0; r < NB; ++r) s = fma(A[((size_t)k0 + r) * ld + col], x[r], s);
  v[col] -= s;
}

__global__ void add_diag_kernel(double* A, size_t ld, uint32_t n, double reg) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(size_t)i * ld + i] += reg;
}

// trace and max |diag| of S (solve_with_cholesky's regularisation base, explicit_schur.rs:575-590)
__global__ void __launch_bounds__(1024) diag_stats_kernel(const double* A, size_t ld, uint32_t n, double* out2) {
  __shared__ double sh[1024];
  double tr = 0.0, mx = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) { const double d = A[(size_t)i * ld + i]; tr += d; mx = dmax(mx, fabs(d)); }
  tr = block_reduce_sum(tr, sh);
  sh[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sh[threadIdx.x] = dmax(sh[threadIdx.x], sh[threadIdx.x + s]); __syncthreads(); }
  if (threadIdx.x == 0) { out2[0] = tr; out2[1] = sh[0]; }
}

// ---- dense PCG pieces (solve_with_pcg, explicit_schur.rs:639-756) -----------------------------------
// y = S p, one warp per row of the full symmetric matrix
__global__ void __launch_bounds__(256) dense_symv_kernel(const double* __restrict__ S, size_t ld, const double* __restrict__ p, double* __restrict__ y,
                                                         uint32_t n, const DevState* st) {
  if (st->pcg_done) return;
  const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* Sr = S + (size_t)row * ld;
  double s = 0.0;
  for (uint32_t c = lane; c < n; c += 32) s = fma(Sr[c], p[c], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) y[row] = s;
}

// scalar Jacobi preconditioner: 1/d if |d| > 1e-12 else 1 (explicit_schur.rs:655-668)
__global__ void jacobi_diag_kernel(const double* S, size_t ld, uint32_t n, double* pinv) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = S[(size_t)i * ld + i];
  pinv[i] = fabs(d) > 1e-12 ? 1.0 / d : 1.0;
}

__global__ void __launch_bounds__(1024) dense_pcg_init_kernel(const double* __restrict__ b, const double* __restrict__ pinv, double* x, double* r,
                                                              double* z, double* p, DevState* st, uint32_t n, int max_it, double cg_tol) {
  __shared__ double sh[1024];
  double bb = 0.0, rz = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) {
    const double v = b[i], zv = pinv[i] * v;
    r[i] = v; x[i] = 0.0; z[i] = zv; p[i] = zv;
    bb += v * v; rz += v * zv;
  }
  bb = block_reduce_sum(bb, sh);
  rz = block_reduce_sum(rz, sh);
  if (threadIdx.x == 0) {
    const double b_norm = sqrt(bb);
    st->b_norm = b_norm; st->pcg_tol = cg_tol * dmax(b_norm, 1.0); st->rz_old = rz; st->r_norm = b_norm;
    st->pcg_iters = 0; st->pcg_max = max_it; st->pcg_done = max_it <= 0 ? 1 : 0;
  }
}

__global__ void __launch_bounds__(1024) dense_pcg_step_kernel(const double* __restrict__ ap, const double* __restrict__ pinv, double* x, double* r,
                                                              double* z, double* p, DevState* st, uint32_t n) {
  __shared__ double sh[1024];
  if (st->pcg_done) return;
  const int iters = st->pcg_iters + 1;
  const double rz_old = st->rz_old, tol = st->pcg_tol;
  const int max_it = st->pcg_max;
  double v = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) v = fma(p[i], ap[i], v);
  const double p_ap = block_reduce_sum(v, sh);
  if (fabs(p_ap) < 1e-30) { if (threadIdx.x == 0) { st->pcg_iters = iters; st->pcg_done = 1; } return; }
  const double alpha = rz_old / p_ap;
  v = 0.0;
  double w = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) {
    x[i] += alpha * p[i];
    const double rv = r[i] - alpha * ap[i];
    r[i] = rv;
    v = fma(rv, rv, v);
    const double zv = pinv[i] * rv;
    z[i] = zv;
    w = fma(rv, zv, w);
  }
  const double r_norm = sqrt(block_reduce_sum(v, sh));
  const double rz_new = block_reduce_sum(w, sh);
  if (r_norm < tol || fabs(rz_old) < 1e-30) { if (threadIdx.x == 0) { st->pcg_iters = iters; st->pcg_done = 1; st->r_norm = r_norm; } return; }
  const double beta = rz_new / rz_old;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) p[i] = z[i] + beta * p[i];
  if (threadIdx.x == 0) { st->rz_old = rz_new; st->r_norm = r_norm; st->pcg_iters = iters; if (iters >= max_it) st->pcg_done = 1; }
}

// ----------------------------------------------------------------------------------------------------
// Shared intrinsics (APEX_OPT_SHARED_INTRINSICS): the graph [pose_k, landmarks, intrinsics] of the reference's calibration tests is
// the per-camera-intrinsics problem restricted to "every camera's copy is equal": x_un = P x_sh with P replicating the K
// intrinsics columns, so J_sh = J_un P and, the landmark elimination not touching camera columns, the reduced system is the
// Galerkin projection  S_sh = P^T (S_un - lambda I) P + lambda I,  b_sh = P^T b_un  (lambda I lives in the shared space). The
// kernels above build S_un / b_un as always; these project, and the solved step is replicated back.
// ----------------------------------------------------------------------------------------------------
// one CTA per row of S_sh: pose rows copy their pose columns and sum the intrinsics columns over the cameras; intrinsics rows sum
// over their cameras' rows as well. Writes the full symmetric matrix (identity on the padding).
__global__ void __launch_bounds__(256) shared_project_matrix_kernel(const double* __restrict__ Sun, size_t ldu, double* __restrict__ Ssh, size_t lds,
                                                                    uint32_t ncam, int dc, int K, uint32_t npad_sh, const DevState* st) {
  __shared__ double sh[256];
  const uint32_t np = 6 * ncam, nsh = np + (uint32_t)K, i = blockIdx.x;
  const double lambda = st->damping;
  if (i >= nsh) { for (uint32_t j = threadIdx.x; j < npad_sh; j += 256) Ssh[(size_t)i * lds + j] = i == j ? 1.0 : 0.0; return; }
  if (i < np) {
    const size_t ru = (size_t)(i / 6) * dc + i % 6;
    for (uint32_t j = threadIdx.x; j < np; j += 256) Ssh[(size_t)i * lds + j] = Sun[ru * ldu + (size_t)(j / 6) * dc + j % 6];
    for (int k = 0; k < K; ++k) {
      double s = 0.0;
      for (uint32_t c2 = threadIdx.x; c2 < ncam; c2 += 256) s += Sun[ru * ldu + (size_t)c2 * dc + 6 + k];
      s = block_reduce_sum(s, sh);
      if (threadIdx.x == 0) Ssh[(size_t)i * lds + np + k] = s;
    }
  } else {
    const int k = (int)(i - np);
    for (uint32_t j = threadIdx.x; j < np; j += 256) {   // column j of the pose part: sum over the cameras' intrinsics rows k
      const size_t cu = (size_t)(j / 6) * dc + j % 6;
      double s = 0.0;
      for (uint32_t c1 = 0; c1 < ncam; ++c1) s += Sun[((size_t)c1 * dc + 6 + k) * ldu + cu];
      Ssh[(size_t)i * lds + j] = s;
    }
    for (int l = 0; l < K; ++l) {
      double s = 0.0;
      for (uint32_t e = threadIdx.x; e < ncam * ncam; e += 256) s += Sun[((size_t)(e / ncam) * dc + 6 + k) * ldu + (size_t)(e % ncam) * dc + 6 + l];
      s = block_reduce_sum(s, sh);
      if (threadIdx.x == 0) Ssh[(size_t)i * lds + np + l] = s - (k == l ? (double)(ncam - 1) * lambda : 0.0);   // P^T (lambda I) P = ncam lambda there
    }
  }
  for (uint32_t j = nsh + threadIdx.x; j < npad_sh; j += 256) Ssh[(size_t)i * lds + j] = 0.0;
}
// v_sh = P^T v_un (sign * v_un): pose entries copied, intrinsics entries summed over the cameras in camera order
__global__ void shared_project_vector_kernel(const double* __restrict__ vun, double* __restrict__ vsh, uint32_t ncam, int dc, int K) {
  const uint32_t np = 6 * ncam, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) vsh[i] = vun[(size_t)(i / 6) * dc + i % 6];
  else if (i < np + (uint32_t)K) {
    double s = 0.0;
    for (uint32_t cam = 0; cam < ncam; ++cam) s += vun[(size_t)cam * dc + 6 + (i - np)];
    vsh[i] = s;
  }
}
// x_un = P x_sh
__global__ void shared_expand_vector_kernel(const double* __restrict__ vsh, double* __restrict__ vun, uint32_t ncam, int dc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncam * (uint32_t)dc) return;
  const uint32_t cam = i / dc, a = i % dc;
  vun[i] = a < 6 ? vsh[6 * cam + a] : vsh[6 * ncam + (a - 6)];
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
static uint32_t tri_count(uint32_t nb) { return (uint32_t)((uint64_t)nb * (nb + 1) / 2); }

// in-place Cholesky of the npad x npad matrix at L; returns through st->chol_fail
static apex_status dense_cholesky(Ctx& c, double* L, uint32_t npad) {
  cudaStream_t s = c.stream;
  const size_t ld = npad;
  const int smem = 2 * NB * PLD * (int)sizeof(double);
  const int trsm_smem_bytes = 2 * NB * PLD * (int)sizeof(double);
  const int syrk128_smem = SYRK_STAGES * 2 * SB * SPLD * (int)sizeof(double);
  if (!c.chol_attr_set) {   // per context, i.e. per device (function attributes are per device)
    APEX_CUDA_TRY(c, cudaFuncSetAttribute(chol_syrk128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, syrk128_smem));
    APEX_CUDA_TRY(c, cudaFuncSetAttribute(chol_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    APEX_CUDA_TRY(c, cudaFuncSetAttribute(chol_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsm_smem_bytes));
    c.chol_attr_set = true;
  }
  // Three-level right-looking blocking: 64-column steps (potrf, trsm) inside inner panels of NBI = 256 columns inside outer panels
  // of NBO columns (default 1024). A 64-wide step updates only its inner panel's remaining columns (all rows below, k depth 64:
  // bound by HBM traffic, not by the tensor pipe); a finished inner panel updates the remaining columns of its OUTER panel with k
  // depth 256 on the 128x128 tensor-pipe tiles; the matrix to the right of the outer panel is updated ONCE per outer panel with k
  // depth NBO, so a trailing tile is read and written n/NBO times and the pipeline fill + C read-modify-write of a tile (3-4 us) is
  // paid once per 2*128*128*NBO flops. Measured at n = 28 032 with two levels: NBO 256 / 512 / 768 = 295.6 / 280.3 / 277.0 ms (the
  // 64-deep in-panel work grows with the panel: 1.5 NBO / n of the flops at a third of the rate); with the middle level the wide
  // panel no longer pays that.
  const uint32_t NBI = 4 * NB;
  uint32_t NBO = 16 * NB;
  if (const char* e = getenv("APEX_CHOL_PANEL")) { const int v = atoi(e); if (v >= 256 && v <= 4096 && v % 256 == 0) NBO = (uint32_t)v; }   // A/B: outer panel width
  const bool deep128 = !getenv("APEX_CHOL_SYRK64");
  if (!deep128) NBO = NBI;
  const uint32_t nblk = npad / NB;
  cudaEvent_t last_rest = nullptr;
  bool have_rest = false;
  for (uint32_t ko = 0; ko < npad; ko += NBO) {
    const uint32_t kend = std::min(ko + NBO, npad);
    for (uint32_t ki = ko; ki < kend; ki += NBI) {
      const uint32_t kiend = std::min(ki + NBI, kend);
      for (uint32_t k0 = ki; k0 < kiend; k0 += NB) {
        chol_potrf_kernel<<<1, 256, 0, s>>>(L, ld, k0, c.state.p);
        c.launches++;
        const uint32_t rem = nblk - k0 / NB - 1;  // 64-row blocks below the diagonal block
        if (rem == 0) break;
        chol_trsm_kernel<<<rem, 256, trsm_smem_bytes, s>>>(L, ld, k0);
        c.launches++;
        const uint32_t ncol = (kiend - k0) / NB - 1;  // inner-panel column tiles still to update
        if (ncol) {
          chol_syrk_kernel<<<dim3(rem, ncol), 128, smem, s>>>(L, ld, k0, NB, k0 + NB);
          c.launches++;
        }
      }
      if (kiend < kend) {   // the rest of the outer panel: k depth 256, rows from kiend down, column tiles [kiend, kend)
        const uint32_t remm = (npad - kiend) / SB, nym = (kend - kiend) / SB;
        chol_syrk128_kernel<<<syrk128_tiles(remm, 0, nym), SYRK_THREADS, syrk128_smem, s>>>(L, ld, ki, kiend - ki, kiend, 0, remm);
        c.launches++;
      }
    }
    if (kend < npad) {
      if ((npad - kend) % SB == 0 && (kend - ko) % SKC == 0 && deep128) {
        // Look-ahead: the deep update is split by columns. The 256 columns of the NEXT panel are updated on the main
        // stream, which then goes straight on to factor that panel (potrf / trsm are latency bound on a handful of
        // CTAs); everything to the right is updated on a second, low-priority stream at the same time. Events keep the
        // order: "rest" update j waits for panel j (E_j); the next panel's columns wait for the previous rest update (R_j-1).
        const uint32_t rem = (npad - kend) / SB, head = std::min<uint32_t>(rem, NBO / SB);
        const bool lookahead = !getenv("APEX_CHOL_NO_LOOKAHEAD") && rem > head;
        if (lookahead) {
          if (!c.stream2) {
            int lo = 0, hi = 0;
            APEX_CUDA_TRY(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            APEX_CUDA_TRY(c, cudaStreamCreateWithPriority(&c.stream2, cudaStreamNonBlocking, lo));
          }
          const size_t pj = ko / NBO;
          while (c.chol_events.size() < 2 * (pj + 1)) { cudaEvent_t e = nullptr; APEX_CUDA_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c.chol_events.push_back(e); }
          cudaEvent_t Ej = c.chol_events[2 * pj], Rj = c.chol_events[2 * pj + 1];
          APEX_CUDA_TRY(c, cudaEventRecord(Ej, s));
          if (have_rest) APEX_CUDA_TRY(c, cudaStreamWaitEvent(s, last_rest, 0));
          chol_syrk128_kernel<<<syrk128_tiles(rem, 0, head), SYRK_THREADS, syrk128_smem, s>>>(L, ld, ko, kend - ko, kend, 0, rem);
          APEX_CUDA_TRY(c, cudaStreamWaitEvent(c.stream2, Ej, 0));
          chol_syrk128_kernel<<<syrk128_tiles(rem, head, rem - head), SYRK_THREADS, syrk128_smem, c.stream2>>>(L, ld, ko, kend - ko, kend, head, rem);
          APEX_CUDA_TRY(c, cudaEventRecord(Rj, c.stream2));
          last_rest = Rj; have_rest = true;
          c.launches++;
        } else {
          if (have_rest) { APEX_CUDA_TRY(c, cudaStreamWaitEvent(s, last_rest, 0)); have_rest = false; }
          chol_syrk128_kernel<<<syrk128_tiles(rem, 0, rem), SYRK_THREADS, syrk128_smem, s>>>(L, ld, ko, kend - ko, kend, 0, rem);
        }
      } else {
        const uint32_t rem = (npad - kend) / NB;
        chol_syrk_kernel<<<dim3(rem, rem), 128, smem, s>>>(L, ld, ko, kend - ko, kend);
      }
      c.launches++;
    }
  }
  if (have_rest) APEX_CUDA_TRY(c, cudaStreamWaitEvent(s, last_rest, 0));  // the main stream continues behind the last rest update
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

// Measurement aid (apex_dense_cholesky_bench): factor a synthetic SPD matrix (symmetric hash noise in [-1, 1), diagonal n:
// strictly diagonally dominant) `reps` times and report the average time of the factorisation alone.
__global__ void chol_bench_fill_kernel(double* A, uint32_t n) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const uint32_t i = (uint32_t)(idx / n), j = (uint32_t)(idx % n);
  const uint32_t lo = min(i, j), hi = max(i, j);
  uint64_t h = ((uint64_t)hi << 32 | lo) * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  A[idx] = i == j ? (double)n : (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
apex_status dense_cholesky_bench(Ctx& c, uint32_t n, int reps, double* ms_out) {
  const uint32_t npad = (n + 2 * NB - 1) / (2 * NB) * (2 * NB);  // multiple of 128: the deep trailing update works on 128x128 tiles
  APEX_CUDA_TRY(c, c.S.alloc((size_t)npad * npad));
  cudaEvent_t e0, e1;
  APEX_CUDA_TRY(c, cudaEventCreate(&e0));
  APEX_CUDA_TRY(c, cudaEventCreate(&e1));
  double total = 0.0;
  for (int r = -1; r < reps; ++r) {  // one warm-up
    chol_bench_fill_kernel<<<(unsigned)(((size_t)npad * npad + 255) / 256), 256, 0, c.stream>>>(c.S.p, npad);
    APEX_CUDA_TRY(c, cudaMemsetAsync(&c.state.p->chol_fail, 0, sizeof(int32_t), c.stream));
    APEX_CUDA_TRY(c, cudaEventRecord(e0, c.stream));
    APEX_TRY(dense_cholesky(c, c.S.p, npad));
    APEX_CUDA_TRY(c, cudaEventRecord(e1, c.stream));
    APEX_CUDA_TRY(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    APEX_CUDA_TRY(c, cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 0) total += ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  APEX_TRY(sync_state(c));
  if (c.h_state->chol_fail) { c.err = "benchmark matrix not positive definite?"; return APEX_ERR_FACTORIZATION_FAILED; }
  if (ms_out) *ms_out = total / std::max(reps, 1);
  return APEX_OK;
}

static apex_status dense_cholesky_solve(Ctx& c, const double* L, uint32_t npad, double* v) {
  cudaStream_t s = c.stream;
  const size_t ld = npad;
  const uint32_t nblk = npad / NB;
  for (uint32_t kb = 0; kb < nblk; ++kb) {
    const uint32_t k0 = kb * NB;
    trsv_diag_kernel<<<1, NB, 0, s>>>(L, ld, k0, v, 0);
    c.launches++;
    const uint32_t rows = npad - k0 - NB;
    if (rows) { trsv_update_fwd_kernel<<<(rows + 31) / 32, 256, 0, s>>>(L, ld, k0, v, npad); c.launches++; }
  }
  for (uint32_t kb = nblk; kb-- > 0;) {
    const uint32_t k0 = kb * NB;
    trsv_diag_kernel<<<1, NB, 0, s>>>(L, ld, k0, v, 1);
    c.launches++;
    if (k0) { trsv_update_bwd_kernel<<<(k0 + 255) / 256, 256, 0, s>>>(L, ld, k0, v); c.launches++; }
  }
  APEX_CUDA_TRY(c, cudaGetLastError());
  return APEX_OK;
}

// SparseSchurComplementSolver::solve_augmented_equation steps 2-8 (explicit_schur.rs:1162-1234) on the current
// linearization. Leaves the camera step in c.step_cam and the landmark step in c.step_pt.
apex_status solve_explicit(Ctx& c, bool use_pcg, int cg_max_it, double cg_tol) {
  cudaStream_t s = c.stream;
  uint32_t n = c.ncam * c.dc;
  uint32_t npad = (n + 2 * NB - 1) / (2 * NB) * (2 * NB);  // multiple of 128: the deep trailing update works on 128x128 tiles
  size_t ld = npad;
  size_t nn = (size_t)npad * npad;
  APEX_CUDA_TRY(c, c.S.alloc(nn * ((use_pcg && !c.shared_intr) ? 1 : 2)));  // [S | factor workspace]
  APEX_CUDA_TRY(c, c.dvec.alloc((size_t)npad * 2 + 8));
  double* S = c.S.p;
  APEX_CUDA_TRY(c, cudaMemsetAsync(S, 0, nn * sizeof(double), s));
  // --- S ---
  cudaEvent_t* evf = c.prof ? prof_pair(c.ev_form, c.ev_form_used++) : nullptr;
  if (evf) cudaEventRecord(evf[0], s);
  const bool by_pairs = !getenv("APEX_SCHUR_FORM_ATOMIC");   // (the reduction-per-element kernel stays for A/B measurements)
  if (c.ntiles && by_pairs) {
    APEX_TRY(build_schur_pairs(c));
    const unsigned grid = (c.npair_blocks + 7) / 8;
    if (grid) {
      APEX_CUDA_TRY(c, c.E.alloc((size_t)c.nchunks * TILE * (2 * (size_t)c.dc + 6)));   // observation-major rows
      switch (c.dc) {
        case 6: pack_obs_major_kernel<6><<<c.nchunks, TILE, 0, s>>>(c.J.p, c.slot_cs8.p, c.E.p); break;
        case 9: pack_obs_major_kernel<9><<<c.nchunks, TILE, 0, s>>>(c.J.p, c.slot_cs8.p, c.E.p); break;
        case 10: pack_obs_major_kernel<10><<<c.nchunks, TILE, 0, s>>>(c.J.p, c.slot_cs8.p, c.E.p); break;
        case 11: pack_obs_major_kernel<11><<<c.nchunks, TILE, 0, s>>>(c.J.p, c.slot_cs8.p, c.E.p); break;
        case 12: pack_obs_major_kernel<12><<<c.nchunks, TILE, 0, s>>>(c.J.p, c.slot_cs8.p, c.E.p); break;
        case 14: pack_obs_major_kernel<14><<<c.nchunks, TILE, 0, s>>>(c.J.p, c.slot_cs8.p, c.E.p); break;
        case 15: pack_obs_major_kernel<15><<<c.nchunks, TILE, 0, s>>>(c.J.p, c.slot_cs8.p, c.E.p); break;
        default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
      }
      c.launches++;
      PairArgs pa{c.pair_blocks.p, c.pair_slots.p, c.slot_lpg.p, c.E.p, c.hinv.p, S, ld, c.npl, c.npair_blocks};
      switch (c.dc) {
        case 6: schur_form_pairs_kernel<6><<<grid, 256, 0, s>>>(pa); break;
        case 9: schur_form_pairs_kernel<9><<<grid, 256, 0, s>>>(pa); break;
        case 10: schur_form_pairs_kernel<10><<<grid, 256, 0, s>>>(pa); break;
        case 11: schur_form_pairs_kernel<11><<<grid, 256, 0, s>>>(pa); break;
        case 12: schur_form_pairs_kernel<12><<<grid, 256, 0, s>>>(pa); break;
        case 14: schur_form_pairs_kernel<14><<<grid, 256, 0, s>>>(pa); break;
        case 15: schur_form_pairs_kernel<15><<<grid, 256, 0, s>>>(pa); break;
        default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
      }
      c.launches++;
    }
  } else if (c.ntiles) {
    FormArgs fa{c.tiles.p, c.slot_cam.p, c.slot_lp.p, c.pt_slot0.p, c.pt_cnt.p, c.cslot_meta.p, c.J.p, c.hinv.p, S, ld, c.npl};
    switch (c.dc) {
      case 6: schur_form_kernel<6><<<c.ntiles, TILE, 0, s>>>(fa); break;
      case 9: schur_form_kernel<9><<<c.ntiles, TILE, 0, s>>>(fa); break;
      case 10: schur_form_kernel<10><<<c.ntiles, TILE, 0, s>>>(fa); break;
      case 12: schur_form_kernel<12><<<c.ntiles, TILE, 0, s>>>(fa); break;
      case 11: schur_form_kernel<11><<<c.ntiles, TILE, 0, s>>>(fa); break;
      case 14: schur_form_kernel<14><<<c.ntiles, TILE, 0, s>>>(fa); break;
      case 15: schur_form_kernel<15><<<c.ntiles, TILE, 0, s>>>(fa); break;
      default: c.err = "unsupported dc"; return APEX_ERR_UNSUPPORTED;
    }
    c.launches++;
  }
  {
    const size_t total = (size_t)c.ncam * c.dc * c.dc + (npad - n);
    schur_diag_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(S, ld, c.hcc.p, c.state.p, c.ncam, c.dc, n, npad, c.rank == 0 ? 1 : 0);
    c.launches++;
  }
  APEX_CUDA_TRY(c, cudaGetLastError());
  APEX_TRY(allreduce_sum(c, S, nn));
  {
    const uint32_t nt = npad / 32;
    symmetrize_drop_kernel<<<tri_count(nt), dim3(32, 8), 0, s>>>(S, ld, nt);
    c.launches++;
  }
  if (evf) cudaEventRecord(evf[1], s);
  // --- reduced gradient ---
  APEX_TRY(launch_reduced_gradient(c, c.vb.p));
  APEX_CUDA_TRY(c, cudaGetLastError());
  const double* bsrc = c.vb.p;
  if (c.shared_intr) {
    // project the per-camera-intrinsics system onto the shared variable; from here on (n, npad, ld) are the shared system's
    const uint32_t nsh = 6 * c.ncam + (uint32_t)c.K, npsh = (nsh + 2 * NB - 1) / (2 * NB) * (2 * NB);
    double* tmp = S + nn;
    shared_project_matrix_kernel<<<npsh, 256, 0, s>>>(S, ld, tmp, npsh, c.ncam, c.dc, c.K, npsh, c.state.p);
    APEX_CUDA_TRY(c, cudaMemcpyAsync(S, tmp, (size_t)npsh * npsh * sizeof(double), cudaMemcpyDeviceToDevice, s));
    APEX_CUDA_TRY(c, c.sh_vec.alloc(2 * (size_t)nsh));
    shared_project_vector_kernel<<<(nsh + 255) / 256, 256, 0, s>>>(c.vb.p, c.vr.p, c.ncam, c.dc, c.K);          // b_sh (vr is free: direct solve)
    shared_project_vector_kernel<<<(nsh + 255) / 256, 256, 0, s>>>(c.gc, c.sh_vec.p, c.ncam, c.dc, c.K);        // gradient in the shared layout
    c.launches += 3;
    APEX_CUDA_TRY(c, cudaGetLastError());
    bsrc = c.vr.p;
    n = nsh; npad = npsh; ld = npsh; nn = (size_t)npsh * npsh;
  }

  if (!use_pcg) {
    double* L = S + nn;
    double* v = c.dvec.p;
    double reg = 0.0, base = 0.0;
    bool solved = false;
    for (int attempt = -1; attempt < 5 && !solved; ++attempt) {
      if (attempt >= 0) {
        if (attempt == 0) {
          diag_stats_kernel<<<1, 1024, 0, s>>>(S, ld, n, v + npad);
          c.launches++;
          double h2[2];
          APEX_CUDA_TRY(c, cudaMemcpyAsync(h2, v + npad, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
          APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
          base = std::max(std::max(h2[0] / (double)n, h2[1]), 1.0);
        }
        reg = base * std::pow(10.0, attempt - 4);
      }
      APEX_CUDA_TRY(c, cudaMemcpyAsync(L, S, nn * sizeof(double), cudaMemcpyDeviceToDevice, s));
      if (reg != 0.0) { add_diag_kernel<<<(n + 255) / 256, 256, 0, s>>>(L, ld, n, reg); c.launches++; }
      APEX_CUDA_TRY(c, cudaMemsetAsync(&c.state.p->chol_fail, 0, sizeof(int32_t), s));
      cudaEvent_t* evc = c.prof ? prof_pair(c.ev_chol, c.ev_chol_used++) : nullptr;
      if (evc) cudaEventRecord(evc[0], s);
      APEX_TRY(dense_cholesky(c, L, npad));
      if (evc) cudaEventRecord(evc[1], s);
      c.chol_n = npad;
      APEX_TRY(sync_state(c));
      if (c.h_state->singular_landmark) { c.err = "Landmark block singular"; return APEX_ERR_SINGULAR_MATRIX; }
      solved = c.h_state->chol_fail == 0;
    }
    if (!solved) { c.err = "Schur complement singular after 5 regularization attempts"; return APEX_ERR_SINGULAR_MATRIX; }
    APEX_CUDA_TRY(c, cudaMemsetAsync(v, 0, (size_t)npad * sizeof(double), s));
    APEX_CUDA_TRY(c, cudaMemcpyAsync(v, bsrc, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    APEX_TRY(dense_cholesky_solve(c, L, npad, v));
    if (c.shared_intr) {   // keep the shared step for the norms, replicate it into every camera's block
      APEX_CUDA_TRY(c, cudaMemcpyAsync(c.sh_vec.p + n, v, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
      shared_expand_vector_kernel<<<(c.ncam * c.dc + 255) / 256, 256, 0, s>>>(v, c.step_cam.p, c.ncam, c.dc);
      c.launches++;
    } else APEX_CUDA_TRY(c, cudaMemcpyAsync(c.step_cam.p, v, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    c.last_pcg_iters = 0;
  } else {
    double* dinv = c.dvec.p;
    jacobi_diag_kernel<<<(n + 255) / 256, 256, 0, s>>>(S, ld, n, dinv);
    dense_pcg_init_kernel<<<1, 1024, 0, s>>>(c.vb.p, dinv, c.step_cam.p, c.vr.p, c.vz.p, c.vp.p, c.state.p, n, cg_max_it, cg_tol);
    c.launches += 2;
    const int BATCH = 10;
    int enq = 0;
    while (enq < cg_max_it) {
      const int nb = std::min(BATCH, cg_max_it - enq);
      for (int i = 0; i < nb; ++i) {
        dense_symv_kernel<<<(n + 7) / 8, 256, 0, s>>>(S, ld, c.vp.p, c.vy.p, n, c.state.p);
        dense_pcg_step_kernel<<<1, 1024, 0, s>>>(c.vy.p, dinv, c.step_cam.p, c.vr.p, c.vz.p, c.vp.p, c.state.p, n);
        c.launches += 2;
      }
      APEX_CUDA_TRY(c, cudaGetLastError());
      enq += nb;
      APEX_TRY(sync_state(c));
      if (c.h_state->pcg_done) break;
    }
    if (cg_max_it <= 0) APEX_TRY(sync_state(c));
    c.last_pcg_iters = c.h_state->pcg_iters;
    if (c.h_state->singular_landmark) { c.err = "Landmark block singular"; return APEX_ERR_SINGULAR_MATRIX; }
  }
  // --- back-substitution (explicit_schur.rs:980-1029) ---
  APEX_TRY(launch_schur_tiles(c, MODE_BACKSUB, c.step_cam.p, nullptr, 0));
  return APEX_OK;
}

}  // namespace apex
