"""Summarise ncu artefacts into profiles/: python tools/ncu_summary.py launches <csv> | metrics <ncu-rep>"""
import collections, csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
    agg = collections.defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > vi and r[mi] == "gpu__time_duration.sum":
            agg[r[ki][:80]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# kernel launches (gpu__time_duration.sum, ns; cold-cache serialised ncu pass) total {tot/1e3:.1f} us")
    print(f"{'kernel':80s} {'n':>5s} {'total_us':>10s} {'mean_us':>9s} {'min_us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:80s} {len(v):5d} {sum(v)/1e3:10.1f} {sum(v)/len(v)/1e3:9.2f} {min(v)/1e3:9.2f} {100*sum(v)/tot:6.1f}%")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__cluster_max_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]

def metrics(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    ki = H.index("Kernel Name")
    for r in rows[2:]:
        print(f"## {r[ki][:100]}")
        for w in WANT:
            for i, h in enumerate(H):
                if h == w or h.endswith("." + w):
                    print(f"{w:75s} {rows[1][i]:>12s} {r[i]}")
                    break
        print()

if __name__ == "__main__":
    {"launches": launches, "metrics": metrics}[sys.argv[1]](sys.argv[2])
