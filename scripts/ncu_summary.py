"""Summarise an .ncu-rep: headline metrics per profiled launch + hot SASS regions (needs -lineinfo + --import-source)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
def page(rep, name):
    """rep: an .ncu-rep, or the prefix of `<prefix>.raw.csv` / `<prefix>.source.csv.gz` exported on the GPU box (scripts/ncu_export.sh)"""
    import gzip, os
    if rep.endswith(".ncu-rep"):
        return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    f = rep + "." + name + ".csv"
    return gzip.open(f + ".gz", "rt").read() if os.path.exists(f + ".gz") else open(f).read()
raw = page(rep, "raw")
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in hdr:
            print(f"  {w} = {r[hdr.index(w)]}  [{rows[1][hdr.index(w)]}]")
src = page(rep, "source")
rows = list(csv.reader(src.splitlines()))
h = rows[1]
iex, ith, ismp, isrc = h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("# Samples"), h.index("Source")
sec = []
for r in rows[2:]:
    if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
        if sec: break
        continue
    sec.append(r)
tot = sum(int(r[iex]) for r in sec)
b = collections.Counter()
for r in sec:
    b[min(32, int(float(r[ith]) // 4 * 4))] += int(r[iex])
print("warp-instructions by avg-active-threads bucket:", sorted((k, round(v / tot, 3)) for k, v in b.items()), "total", tot)
smp = sum(int(r[ismp]) for r in sec)
print("top sampled instructions:")
for r in sorted(sec, key=lambda r: -int(r[ismp]))[:14]:
    print(f"   {int(r[ismp]) / smp:6.3f}  thr {r[ith]:>3}  exec {r[iex]:>10}  {r[isrc].strip()[:80]}")
