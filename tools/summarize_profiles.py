"""Turn the ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed).
usage: python tools/summarize_profiles.py <round-tag> [launches.csv] [prof.ncu-rep]"""
import collections, csv, json, re, subprocess, sys, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/prof_matvec_cur.ncu-rep"
os.makedirs("profiles", exist_ok=True)
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, data = r, rows[i + 1:]
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r"^void |apex::", "", re.sub(r"\(.*", "", r[ki]))
        v = float(r[vi].replace(",", "")) * {"ms": 1000.0, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 python bench.py --steps 2 --warmup 1 --cpu-baseline 0\n")
        f.write(f"# per-launch times under ncu are cold-cache and serialised: compare SHARES. {sum(a[0] for a in agg.values())} launches, {tot:.0f} us total\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{a[1] / tot * 100:6.2f}%  n={a[0]:4d}  avg={a[1] / a[0]:9.1f} us  {k}\n")
    print(open(f"profiles/{tag}_launches_summary.txt").read())
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
    out = {}
    for k in keys:
        if k in hdr:
            out[k] = (r[hdr.index(k)], units[hdr.index(k)])
    def tobytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic = tobytes(*out["dram__bytes_read.sum"]) + tobytes(*out["dram__bytes_write.sum"])
    with open(f"profiles/{tag}_matvec_ncu_summary.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:schur_chunk_kernel -c 1 python bench.py --steps 1 --warmup 0 (Venice-1778 shape)\n")
        for k, (v, u) in out.items():
            f.write(f"{k:70s} {v} {u}\n")
        f.write(f"dram traffic per launch (read+write) = {traffic / 1e9:.4f} GB; algorithmic bytes per launch = 1.1077 GB\n")
        stalls = [(k, r[i]) for i, k in enumerate(hdr) if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct")]
        for k, v in sorted(stalls, key=lambda kv: -float(kv[1] or 0))[:8]:
            f.write(f"{k:70s} {v} %\n")
    json.dump({"dram_bytes_per_launch": traffic, "source": f"profiles/{tag}_matvec_ncu_summary.txt", "kernel": out["Kernel Name"][0]}, open("profiles/matvec_traffic.json", "w"))
    print(open(f"profiles/{tag}_matvec_ncu_summary.txt").read())
