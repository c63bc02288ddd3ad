"""Read an .ncu-rep here (no GPU): headline metrics, stall breakdown and the hottest SASS lines with their stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_matvec_cur.ncu-rep"
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
for k in ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "dram__bytes_read.sum.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct"):
    if k in m: print(f"{k:75s} {m[k]} {u[k]}")
st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v)) for h, v in zip(hdr, vals) if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued")]
tot = sum(v for _, v in st)
print("stalls:", ", ".join(f"{h} {v / tot * 100:.1f}%" for h, v in sorted(st, key=lambda x: -x[1])[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2, data = rows[1], rows[2:]
iS, isrc = h2.index("# Samples"), h2.index("Source")
cols = {k: h2.index(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_lg", "L1 Wavefronts Shared") if k in h2}
tot = sum(int(r[iS]) for r in data)
print("samples", tot)
for i, r in sorted(sorted(enumerate(data), key=lambda x: -int(x[1][iS]))[:ntop]):
    print(i, r[isrc][:64].ljust(64), r[iS], {k.replace("stall_", ""): r[c] for k, c in cols.items() if r[c] not in ("0", "")})
