// comm.cu — NVLink peer-memory plumbing for the fused "all-reduce + dot product" step of the PCG loop.
//
// The per-iteration exchange of the sharded Schur operator is tiny (ncam*dc doubles: 128 KB for Venice-1778) and
// latency-bound: an NCCL all-reduce costs 30-40 us at 8 ranks, as much as the operator kernel itself on a
// 1/8 shard. Instead every rank keeps its partial result in a buffer that all peers map (cudaIpc over
// NVLink / NVSwitch); ar_reduce_pap_kernel (schur.cu) publishes a sequence flag to the peers, waits for theirs
// and sums the nranks partial vectors in rank order straight from peer memory - bitwise identical on every rank -
// fused with the p.Ap dot product of the PCG iteration. NCCL stays for the large once-per-LM-iteration reductions.
#include <cstdlib>
#include <cstring>

#include "apex_ctx.h"

namespace apex {

void release_peer_allreduce(Ctx& c) {
  for (size_t r = 0; r < c.peer_buf.size(); ++r) {
    if ((int)r == c.rank) continue;
    if (c.peer_buf[r]) cudaIpcCloseMemHandle(c.peer_buf[r]);
    if (c.peer_flags[r]) cudaIpcCloseMemHandle(c.peer_flags[r]);
  }
  c.peer_buf.clear();
  c.peer_flags.clear();
  c.p2p_ok = false;
}

// Collective over all ranks (called from apex_problem_upload). On any failure every rank falls back to NCCL.
apex_status setup_peer_allreduce(Ctx& c, size_t n) {
  const bool had_peers = c.p2p_ok;
  release_peer_allreduce(c);
  if (c.nranks <= 1 || getenv("APEX_NO_P2P")) return APEX_OK;
  cudaStream_t s = c.stream;
  if (had_peers) {
    // re-upload: every rank has closed its mappings of the peers' buffers above; nobody may free a buffer a peer still maps
    // (CUDA IPC rule), so meet here - a one-element all-reduce and a stream sync - before the exported buffers are released
    APEX_TRY(allreduce_sum(c, c.state.p->agree, 1));
    APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  }
  int ok = 1;
  // fresh allocations: an IPC handle refers to the allocation it was taken from
  c.arbuf.release();
  c.arflags.release();
  if (c.arbuf.alloc(2 * n) != cudaSuccess || c.arflags.alloc((size_t)c.nranks) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  c.ar_n = n;
  cudaIpcMemHandle_t hb, hf;
  std::memset(&hb, 0, sizeof(hb));
  std::memset(&hf, 0, sizeof(hf));
  if (ok) {
    cudaMemsetAsync(c.arbuf.p, 0, 2 * n * sizeof(double), s);
    cudaMemsetAsync(c.arflags.p, 0, (size_t)c.nranks * sizeof(unsigned long long), s);
    cudaMemsetAsync(&c.state.p->ar_seq, 0, sizeof(unsigned long long), s);
    if (cudaIpcGetMemHandle(&hb, c.arbuf.p) != cudaSuccess || cudaIpcGetMemHandle(&hf, c.arflags.p) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  // exchange the two 64-byte handles of every rank: bytes widened to doubles and summed into a zero-filled table
  const size_t per = 2 * sizeof(cudaIpcMemHandle_t);  // 128
  std::vector<double> table((size_t)c.nranks * per + 1, 0.0);
  const unsigned char* pb = reinterpret_cast<const unsigned char*>(&hb);
  const unsigned char* pf = reinterpret_cast<const unsigned char*>(&hf);
  for (size_t i = 0; i < sizeof(cudaIpcMemHandle_t); ++i) {
    table[(size_t)c.rank * per + i] = pb[i];
    table[(size_t)c.rank * per + sizeof(cudaIpcMemHandle_t) + i] = pf[i];
  }
  table[(size_t)c.nranks * per] = ok ? 0.0 : 1.0;  // number of ranks that failed so far
  DevBuf<double> dt;
  if (dt.alloc(table.size()) != cudaSuccess) { c.err = "peer all-reduce setup: out of memory"; return APEX_ERR_CUDA; }
  APEX_CUDA_TRY(c, cudaMemcpyAsync(dt.p, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  APEX_TRY(allreduce_sum(c, dt.p, table.size()));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(table.data(), dt.p, table.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  dt.release();
  if (table[(size_t)c.nranks * per] != 0.0) return APEX_OK;  // someone failed: NCCL path everywhere
  c.peer_buf.assign(c.nranks, nullptr);
  c.peer_flags.assign(c.nranks, nullptr);
  int opened = 1;
  for (int r = 0; r < c.nranks; ++r) {
    if (r == c.rank) { c.peer_buf[r] = c.arbuf.p; c.peer_flags[r] = c.arflags.p; continue; }
    cudaIpcMemHandle_t rb, rf;
    unsigned char* qb = reinterpret_cast<unsigned char*>(&rb);
    unsigned char* qf = reinterpret_cast<unsigned char*>(&rf);
    for (size_t i = 0; i < sizeof(cudaIpcMemHandle_t); ++i) {
      qb[i] = (unsigned char)table[(size_t)r * per + i];
      qf[i] = (unsigned char)table[(size_t)r * per + sizeof(cudaIpcMemHandle_t) + i];
    }
    void *vb = nullptr, *vf = nullptr;
    if (cudaIpcOpenMemHandle(&vb, rb, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&vf, rf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
    c.peer_buf[r] = static_cast<double*>(vb);
    c.peer_flags[r] = static_cast<unsigned long long*>(vf);
  }
  // agree on the outcome
  double fail = opened ? 0.0 : 1.0;
  DevBuf<double> df;
  if (df.alloc(1) != cudaSuccess) { c.err = "peer all-reduce setup: out of memory"; return APEX_ERR_CUDA; }
  APEX_CUDA_TRY(c, cudaMemcpyAsync(df.p, &fail, sizeof(double), cudaMemcpyHostToDevice, s));
  APEX_TRY(allreduce_sum(c, df.p, 1));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(&fail, df.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  df.release();
  if (fail != 0.0) { release_peer_allreduce(c); return APEX_OK; }
  APEX_CUDA_TRY(c, c.d_peer_buf.alloc((size_t)c.nranks));
  APEX_CUDA_TRY(c, c.d_peer_flags.alloc((size_t)c.nranks));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.d_peer_buf.p, c.peer_buf.data(), (size_t)c.nranks * sizeof(double*), cudaMemcpyHostToDevice, s));
  APEX_CUDA_TRY(c, cudaMemcpyAsync(c.d_peer_flags.p, c.peer_flags.data(), (size_t)c.nranks * sizeof(unsigned long long*), cudaMemcpyHostToDevice, s));
  APEX_CUDA_TRY(c, cudaStreamSynchronize(s));
  c.p2p_ok = true;
  return APEX_OK;
}

}  // namespace apex
