#!/usr/bin/env python
"""Turn ncu reports / launch lists under gpurun_out/ into small tracked summaries under profiles/.

    python tools/summarise_profiles.py r01        # prefix for this round
"""
import csv
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def raw_summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        item = {"kernel": d.get("Kernel Name", "")[:120]}
        for k in KEEP:
            if k in d and d[k] != "":
                item[k] = "%s %s" % (d[k], u.get(k, ""))
        stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[k])
                  for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                  and "not_issued" not in k and d[k] not in ("", None)}
        item["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        res.append(item)
    return res


def launch_summary(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        agg.setdefault(row["Kernel Name"][:100], []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    return [{"kernel": k, "launches": len(v), "mean_us": sum(v) / len(v) / 1e3, "share_of_listed_time": sum(v) / tot}
            for k, v in agg.items()]


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    for fn in sorted(os.listdir(SRC)):
        p = os.path.join(SRC, fn)
        if fn.endswith(".ncu-rep"):
            s = raw_summary(p)
            if s:
                json.dump(s, open(os.path.join(OUT, "%s_%s.json" % (prefix, fn[:-8])), "w"), indent=1)
                print("wrote", fn)
        elif fn.startswith("launches") and fn.endswith(".csv"):
            json.dump(launch_summary(p), open(os.path.join(OUT, "%s_%s.json" % (prefix, fn[:-4])), "w"), indent=1)
            print("wrote", fn)
    # DRAM traffic of one full-size K4 launch for bench.py's roofline.traffic
    pool = os.path.join(OUT, "%s_prof_pool.json" % prefix)
    if os.path.exists(pool):
        k = json.load(open(pool))[0]

        def num(key):
            val, unit = k[key].rsplit(" ", 1)
            return float(val.replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        json.dump({"kernel": k["kernel"], "panels_per_launch": 64, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "source": "profiles/%s_prof_pool.json: ncu --set full --clock-control none on `bench.py --steps 2 --warmup 3 "
                             "--no-cpu-baseline --skip-extra` with K4's tuned launch form fixed" % prefix,
                   "traffic_bytes_per_launch": rd + wr}, open(os.path.join(OUT, "roofline_traffic.json"), "w"), indent=1)
        print("wrote roofline_traffic.json", rd + wr)
    for fn in ("kernels.json", "bench.log", "bench_quick.log"):
        p = os.path.join(SRC, fn)
        if os.path.exists(p):
            open(os.path.join(OUT, "%s_%s" % (prefix, fn)), "w").write(open(p).read())


if __name__ == "__main__":
    main()
