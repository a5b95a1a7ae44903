"""Turn gpurun_out/*.ncu-rep and launches.csv into the tracked summaries under profiles/.
   python tools/summarize_ncu.py r01        (needs the `ncu` CLI; no GPU)"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tensor", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


summary = {}
batch = 1024
for kern, name in (("detect", "detect_tc_kernel"), ("code_parts", "code_parts_tc_kernel"),
                   ("spread_codes", "spread_codes_kernel"), ("spread_edges", "spread_edges_kernel"),
                   ("threshold_blocks", "threshold_blocks_kernel")):
    rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, kern))
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
    lines = ["kernel: %s" % kname, "command: python tools/prof_stage.py %s --iters 2   (B=1024 utterances, BASELINE config shapes)" % kern,
             "capture: ncu --set full --clock-control none --import-source on -k regex:%s -s 1 -c 1" % kern, ""]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if any(h == k or h.endswith(k) for k in KEYS):
            lines.append("%-95s %-12s %s" % (h, u, v))
            d[h] = (v, u)
    rd = [to_bytes(v, u) for h, (v, u) in d.items() if h == "dram__bytes_read.sum"]
    wr = [to_bytes(v, u) for h, (v, u) in d.items() if h == "dram__bytes_write.sum"]
    dur = [float(v) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1) for h, (v, u) in d.items() if h == "gpu__time_duration.sum"]
    summary[name] = {"batch": batch, "dram_bytes": (rd[0] + wr[0]) if rd and wr else None,
                     "duration_s_under_ncu": dur[0] if dur else None}
    open(os.path.join(ROOT, "profiles", "%s_%s_ncu_full.txt" % (tag, kern)), "w").write("\n".join(lines) + "\n")
json.dump(summary, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1)

# launch list: per-kernel totals and shares
lc = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lc):
    text = open(lc).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    tot = {}
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)   # -> microseconds
        k = r["Kernel Name"].split("(")[0]
        c = tot.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
    allus = sum(v for _, v in tot.values())
    with open(os.path.join(ROOT, "profiles", "%s_launch_list.txt" % tag), "w") as f:
        f.write("command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu\n")
        f.write("(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)\n\n")
        f.write("%-60s %8s %14s %8s\n" % ("kernel", "launches", "total us", "share"))
        for k, (n, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
            f.write("%-60s %8d %14.1f %7.1f%%\n" % (k[:60], n, v, 100 * v / allus))
    import shutil
    shutil.copy(lc, os.path.join(ROOT, "profiles", "%s_launches.csv" % tag))
print(json.dumps(summary, indent=1))
