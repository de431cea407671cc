#!/usr/bin/env python
"""Turn the raw ncu artefacts in gpurun_out/ into the tracked summaries under
profiles/ (run here, no GPU needed):

  gpurun_out/launches.csv   (ncu --metrics gpu__time_duration.sum launch list)
  gpurun_out/prof.ncu-rep   (ncu --set full capture of the dominant kernels)
  gpurun_out/bench.log      (the plain bench run of the same round)
  gpurun_out/eager_compare.json

  -> profiles/rNN_launches.csv, profiles/rNN_ncu_summary.json,
     profiles/rNN_bench.json, profiles/rNN_eager_compare.json
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

# which launch-list rows are the bench step's four kernels: (substring, grid) -> name
STEP_KERNELS = {"sym_fwd_x_a8": ("rowquant_vec_kernel<1,", "2048"), "sym_fwd_w_w4": ("rowquant_vec_kernel<1,", "2752"),
                "ste_bwd": ("ste_bwd_kernel<1, 0>", "1184")}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct"]


def to_bytes(value, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value) * mult


def main():
    os.makedirs(PROF, exist_ok=True)
    summary = {"tag": TAG, "launch_list": {}, "kernels": {}, "full_capture": []}
    # ---- launch list -----------------------------------------------------
    src = os.path.join(OUT, "launches.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, f"{TAG}_launches.csv"))
        rows = [r for r in csv.reader(open(src)) if len(r) > 10]
        hdr = rows[0]
        ik, ig, ib, iv = (hdr.index(c) for c in ("Kernel Name", "Grid Size", "Block Size", "Metric Value"))
        groups = collections.defaultdict(list)
        for r in rows[1:]:
            try:
                groups[(r[ik], r[ig], r[ib])].append(float(r[iv]) / 1e3)
            except ValueError:
                pass
        for (name, grid, block), v in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
            short = name.replace("qat::<unnamed>::", "").replace("void ", "")[:90]
            summary["launch_list"][f"{short} grid={grid} block={block}"] = {
                "launches": len(v), "mean_us": round(sum(v) / len(v), 2), "min_us": round(min(v), 2),
                "max_us": round(max(v), 2), "total_ms": round(sum(v) / 1e3, 3)}
        # the step's kernels: bwd x / bwd W share one kernel+grid; split by duration
        step = {}
        for key, (sub, grid) in STEP_KERNELS.items():
            v = [x for (name, g, _), vals in groups.items() if sub in name and g.strip("() ,").split(",")[0] == grid
                 for x in vals]
            if not v:
                continue
            if key == "ste_bwd":
                v.sort()
                mid = (v[0] + v[-1]) / 2
                lo, hi = [x for x in v if x < mid], [x for x in v if x >= mid]
                step["ste_bwd_x"] = sum(lo) / max(len(lo), 1)
                step["ste_bwd_w"] = sum(hi) / max(len(hi), 1)
            else:
                step[key] = sum(v) / len(v)
        tot = sum(step.values()) or 1.0
        summary["step_share_from_launch_list"] = {k: {"mean_us": round(v, 2), "share": round(v / tot, 4)}
                                                  for k, v in step.items()}
    # ---- full capture ------------------------------------------------------
    for rep in (os.path.join(OUT, "prof.ncu-rep"), os.path.join(OUT, "prof_gemm.ncu-rep")):
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            ent = {"kernel": r[idx["Kernel Name"]].replace("unnamed>::", "")}
            for w in WANT:
                if w in idx and r[idx[w]] != "":
                    try:
                        ent[w] = float(r[idx[w]])
                    except ValueError:
                        ent[w] = r[idx[w]]
                    if w.startswith("dram__bytes"):
                        ent[w] = to_bytes(r[idx[w]], units[idx[w]])
            ent["dram_bytes_per_launch"] = ent.get("dram__bytes_read.sum", 0) + ent.get("dram__bytes_write.sum", 0)
            summary["full_capture"].append(ent)
    if summary["full_capture"]:
        # map onto bench.py's kernel names by order of appearance: x, W, bwd x, bwd W
        by = collections.defaultdict(list)
        for e in summary["full_capture"]:
            by[e["kernel"].split("(")[0]].append(e)
        for kname, ents in by.items():
            ents_sorted = sorted(ents, key=lambda e: e.get("gpu__time_duration.sum", 0))
            if "rowquant_vec_kernel<1" in kname and "1, 0>" in kname:
                small = [e for e in ents if e.get("launch__grid_size") == 2048.0]
                big = [e for e in ents if e.get("launch__grid_size") == 2752.0]
                if small:
                    summary["kernels"]["sym_fwd_x_a8"] = small[0]
                if big:
                    summary["kernels"]["sym_fwd_w_w4"] = big[0]
            elif "ste_bwd_kernel<1, 0>" in kname:
                summary["kernels"]["ste_bwd_x"] = ents_sorted[0]
                summary["kernels"]["ste_bwd_w"] = ents_sorted[-1]
            elif "qlinear_i8_kernel" in kname:
                summary["kernels"]["qlinear_i8"] = ents_sorted[0]
    # ---- bench line + eager comparison of the same round -------------------
    for name in ("bench.log", "eager_compare.json"):
        src = os.path.join(OUT, name)
        if os.path.exists(src):
            dst = os.path.join(PROF, f"{TAG}_{'bench.json' if name == 'bench.log' else name}")
            if name == "bench.log":
                line = open(src).read().strip().splitlines()[-1]
                json.dump(json.loads(line), open(dst, "w"), indent=1)
            else:
                shutil.copy(src, dst)
    with open(os.path.join(PROF, f"{TAG}_ncu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary.get("step_share_from_launch_list"), indent=1))
    print({k: (v.get("gpu__time_duration.sum"), v.get("dram_bytes_per_launch")) for k, v in summary["kernels"].items()})


if __name__ == "__main__":
    main()
