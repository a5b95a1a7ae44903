"""Turn gpurun_out/<tag>_*.ncu-rep and <tag>_launches.csv into the tracked summaries under profiles/.
   python tools/summarize_ncu.py r02        (needs the `ncu` CLI; no GPU)"""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
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
# report suffix -> command that was profiled
REPORTS = [
    ("detect", "python tools/prof_stage.py detect --iters 2   (B=1024 utterances, config 3; INT8X2 = detect_tn_kernel, the pipeline default)"),
    ("detect_x3", "python tools/prof_stage.py detect --iters 2 --planes 4   (the exact 24-bit variant, detect_tc_kernel)"),
    ("code_parts", "python tools/prof_stage.py code_parts --iters 2   (B=1024 utterances, 100 parts)"),
    ("code_parts_int8", "python tools/prof_stage.py code_parts --iters 2 --impl-code 2   (the int8 form: code_parts_tc_kernel<13, true> + code_parts_refine_kernel)"),
    ("code_parts_int8_200", "python tools/prof_stage.py code_parts --iters 2 --impl-code 2 --parts 200   (200 parts: N = 208, two 256-column accumulators)"),
    ("spread_codes", "python tools/prof_stage.py spread_codes --iters 2"),
    ("spread_edges", "python tools/prof_stage.py spread_edges --iters 2"),
    ("threshold_blocks", "python tools/prof_stage.py threshold_blocks --iters 2"),
    ("detect_float", "python tools/prof_stage.py detect_float --iters 2   (64 utterances x 300 frames x 320 real features, 8 x 40-frame templates)"),
    ("em", "python bench.py --workload em --steps 2 --warmup 3 --no-cpu   (config 4: N=100000, D=12800, K=16; one EM iteration's kernels)"),
]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("tsr::", "")
    if n.startswith("code_parts_tc_kernel") and "true" in n:     # <NSTEPS, true> = the int8 form
        return "code_parts_tc_kernel_int8"
    return n.split("<")[0]


summary_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
for suffix, command in REPORTS:
    rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, suffix))
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = ["command: %s" % command,
             "capture: ncu --set full --clock-control none --import-source on   (after a plain run of the same command exited 0)", ""]
    seen = set()
    for vals in rows[2:]:
        kname = vals[hdr.index("Kernel Name")]
        if kname in seen:
            continue
        seen.add(kname)
        lines.append("kernel: %s" % kname)
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if any(h == k or h.endswith(k) for k in KEYS):
                lines.append("  %-95s %-12s %s" % (h, u, v))
                d[h] = (v, u)
        rd = [to_bytes(v, u) for h, (v, u) in d.items() if h == "dram__bytes_read.sum"]
        wr = [to_bytes(v, u) for h, (v, u) in d.items() if h == "dram__bytes_write.sum"]
        dur = [float(v) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1) for h, (v, u) in d.items() if h == "gpu__time_duration.sum"]
        summary[short(kname) + ("_200parts" if suffix.endswith("_200") else "")] = {"batch": 1024 if suffix != "em" else None, "dram_bytes": (rd[0] + wr[0]) if rd and wr else None,
                                 "duration_s_under_ncu": dur[0] if dur else None, "capture": "%s_%s" % (tag, suffix)}
        lines.append("")
    open(os.path.join(ROOT, "profiles", "%s_%s_ncu_full.txt" % (tag, suffix)), "w").write("\n".join(lines) + "\n")
json.dump(summary, open(summary_path, "w"), indent=1)

# launch list of the default bench command: per-kernel totals, and the shares inside one pipeline step
lc = os.path.join(ROOT, "gpurun_out", "%s_launches.csv" % tag)
if os.path.exists(lc):
    text = open(lc).read()
    rows = list(csv.DictReader(io.StringIO(text[text.index('"ID"'):])))
    seq = []
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1)   # -> microseconds
        seq.append((short(r["Kernel Name"]), v))
    tot = {}
    for k, v in seq:
        c = tot.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
    step_kernels = ("spread_edges_line_kernel", "code_parts_tc_kernel", "spread_codes_kernel", "detect_tn_kernel", "detect_tc_kernel",
                    "detect_decode_kernel")
    em_kernels = [k for k in tot if k.startswith("bitgemm") or k.startswith("em_")]
    with open(os.path.join(ROOT, "profiles", "%s_launch_list.txt" % tag), "w") as f:
        f.write("command: ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv python bench.py --steps 2 --warmup 3 --no-cpu\n")
        f.write("(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes; the raw list is %s_launches.csv)\n\n" % tag)
        for title, names in (("config-3 pipeline step (median per launch; a step = one launch of each)", [k for k in step_kernels if k in tot]),
                             ("config-4 EM iteration (median per launch; E and M contractions = two bitgemm launches)", em_kernels)):
            # median per launch (the stream workload of the same command launches the detection kernel on 720 000 frames)
            med = lambda k: sorted(v for kk, v in seq if kk == k)[len([1 for kk, _ in seq if kk == k]) // 2]   # noqa: E731
            sub = {k: med(k) * (2 if k.startswith("bitgemm") else 1) for k in names}
            allus = sum(sub.values()) or 1.0
            f.write("%s\n%-44s %10s %8s\n" % (title, "kernel", "us", "share"))
            for k, v in sorted(sub.items(), key=lambda x: -x[1]):
                f.write("%-44s %10.1f %7.1f%%\n" % (k, v, 100 * v / allus))
            f.write("%-44s %10.1f\n\n" % ("sum", allus))
        f.write("all launches of the command\n%-60s %8s %14s\n" % ("kernel", "launches", "total us"))
        for k, (n, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
            f.write("%-60s %8d %14.1f\n" % (k[:60], n, v))
    shutil.copy(lc, os.path.join(ROOT, "profiles", "%s_launches.csv" % tag))
print(json.dumps(summary, indent=1))
