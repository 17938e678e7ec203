"""Turn gpurun_out/*.ncu-rep + launches csv into the small, tracked summaries under profiles/.

  python tools/summarize_profiles.py <tag> <launches.csv> <full.ncu-rep> [...more .ncu-rep]
"""
import collections, csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# 1. launch list -> per-kernel mean/share table
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg.setdefault(row["Kernel Name"], []).append(v)
own = {k: v for k, v in agg.items() if "at::" not in k}
per_step = {k: sum(v) / len(v) * (2 if ("Onesweep" in k or "Histogram" in k or "ExclusiveSum" in k) else 1) for k, v in own.items()}
with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: per-kernel device time (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n\n")
    f.write("command: `ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline`\n\n")
    f.write("| kernel | launches captured | mean us |\n|---|---|---|\n")
    for k, v in own.items():
        f.write(f"| `{k[:110]}` | {len(v)} | {sum(v)/len(v):.1f} |\n")
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.writelines(lines)

# 2. full captures -> selected raw metrics + per-instruction hot spots
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
with open(os.path.join(out_dir, f"{tag}_ncu_full.md"), "w") as f:
    f.write(f"# {tag}: ncu --set full --clock-control none --import-source on (one launch per kernel)\n\n")
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            f.write(f"## {r[ix['Kernel Name']][:100]}  ({os.path.basename(rep)})\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in ix:
                    f.write(f"| {w} | {r[ix[w]]} | {units[ix[w]]} |\n")
            f.write("\n")
        # source page: instruction totals per kernel
        names = sorted({r[ix["Kernel Name"]].split("(")[0].split("<")[0].split()[-1] for r in rows[2:]})
        for nm in names:
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{nm}"], capture_output=True, text=True).stdout
            srows = list(csv.reader(io.StringIO(src)))
            if len(srows) < 3:
                continue
            sh = srows[1]; six = {h: i for i, h in enumerate(sh)}
            if "Instructions Executed" not in six:
                continue
            ie = six["Instructions Executed"]
            body = [r for r in srows[2:] if len(r) > ie and r[ie].isdigit()]
            tot = sum(int(r[ie]) for r in body)
            if tot == 0:
                continue
            ops = collections.Counter()
            for r in body:
                s = r[six["Source"]].strip().split()
                op = (s[1] if len(s) > 1 and s[0].startswith("@") else (s[0] if s else "?")).split(".")[0]
                ops[op] += int(r[six["Instructions Executed"]])
            f.write(f"### {nm}: {tot} warp-instructions executed; top opcodes: " + ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in ops.most_common(12)) + "\n\n")
# 3. profiles/ncu_traffic.json: what bench.py's roofline.traffic / issue_active_pct read — keyed by kernel, valid only for the workload
#    it was captured on and for the exact version of the kernel's source file (sha256 prefix)
import hashlib, json
# (substring of the demangled kernel name, key in ncu_traffic.json, source file) — first match wins
PATTERNS = [("preprocess_fwd_kernel<0, 1", "preprocess_fwd_scatter_kernel", "preprocess_fwd.cu"),
            ("preprocess_bwd_tma_kernel<1>", "preprocess_bwd_gather_kernel", "preprocess_bwd.cu"),
            ("blend_bwd2_kernel", "blend_bwd2_kernel", "blend_bwd2.cu"), ("blend_fwd_kernel", "blend_fwd_kernel", "blend_fwd.cu"),
            ("preprocess_fwd_kernel", "preprocess_fwd_kernel", "preprocess_fwd.cu"),
            ("preprocess_bwd_tma_kernel", "preprocess_bwd_tma_kernel", "preprocess_bwd.cu"),
            ("preprocess_bwd_kernel", "preprocess_bwd_kernel", "preprocess_bwd.cu"),
            ("compose_fwd_kernel", "compose_fwd_kernel", "compose.cu"), ("compose_bwd_kernel", "compose_bwd_kernel", "compose.cu"),
            ("emit_pairs_kernel", "emit_pairs_kernel", "binning.cu"), ("emit_big_kernel", "emit_big_kernel", "binning.cu"),
            ("tile_ranges_kernel", "tile_ranges_kernel", "binning.cu"), ("count_compact_kernel", "count_compact_kernel", "binning.cu"),
            ("count_tiles_kernel", "count_tiles_kernel", "binning.cu")]
traffic_path = os.path.join(out_dir, "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
workload = os.environ.get("SGR_PROFILE_WORKLOAD", "C")
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    conv = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    units = rows[1]
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        hit = next(((k, src) for pat, k, src in PATTERNS if pat in name), None)
        if hit is None or "dram__bytes_read.sum" not in ix:
            continue
        key, src_file = hit
        rd = float(r[ix["dram__bytes_read.sum"]]) * conv.get(units[ix["dram__bytes_read.sum"]], 1.0)
        wr = float(r[ix["dram__bytes_write.sum"]]) * conv.get(units[ix["dram__bytes_write.sum"]], 1.0)
        # the capture belongs to the source as it was when it was taken: SGR_PROFILE_COMMIT names that commit (default: the working tree)
        commit = os.environ.get("SGR_PROFILE_COMMIT")
        if commit:
            blob = subprocess.run(["git", "-C", ROOT, "show", f"{commit}:street_gaussians_b200/csrc/{src_file}"], capture_output=True).stdout
        else:
            blob = open(os.path.join(ROOT, "street_gaussians_b200", "csrc", src_file), "rb").read()
        sha = hashlib.sha256(blob).hexdigest()[:16]
        traffic[key] = dict(workload=workload, source_sha16=sha, dram_bytes=rd + wr, dram_bytes_read=rd, dram_bytes_write=wr,
                            time_us=float(r[ix["gpu__time_duration.sum"]]) * (1e-3 if units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0),
                            issue_active_pct=float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]), capture=os.path.basename(rep),
                            commit=os.environ.get("SGR_PROFILE_COMMIT", "working tree"),
                            kernel=name[:120])
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("wrote profiles/", tag, "and ncu_traffic.json with", sorted(traffic))
