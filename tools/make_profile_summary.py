#!/usr/bin/env python
"""Write profiles/<tag>_summary.md from a bench.py JSON line, the ncu launch list of the same command and the
ncu DRAM-traffic pass.  Usage: python tools/make_profile_summary.py r01 gpurun_out/bench.json gpurun_out/launches.csv [gpurun_out/traffic.csv]"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, bench_path, launches_path = sys.argv[1], sys.argv[2], sys.argv[3]
d = json.load(open(bench_path))
if len(sys.argv) > 4:        # ncu DRAM-traffic pass over the GEMM launches of one step -> profiles/<tag>_traffic.json (+ the csv)
    rows = [x for x in csv.reader(open(sys.argv[4])) if len(x) > 14 and x[0].isdigit()]
    rd = sum(float(x[14].replace(",", "")) for x in rows if x[12] == "dram__bytes_read.sum")
    wr = sum(float(x[14].replace(",", "")) for x in rows if x[12] == "dram__bytes_write.sum")
    ns = sum(float(x[14].replace(",", "")) for x in rows if x[12] == "gpu__time_duration.sum")
    nk = len({x[0] for x in rows})                       # kernels of one step (a strided transposed conv is 2 kernels, 1 launch scope)
    n = int(d["roofline"]["launches_per_step"])          # launch scopes of one step: what roofline.achieved is averaged over
    t = {"igemm_tc": {"dram_bytes_per_launch": int((rd + wr) / n), "launches": n, "kernels": nk, "dram_read_bytes_step": rd,
                      "dram_write_bytes_step": wr, "ncu_time_ms_step": ns / 1e6,
                      "command": "tools/prof_bench.sh (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                                 "--clock-control none -k regex:tc_igemm -s 162 -c 54)"}}
    json.dump(t, open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag), "w"), indent=1)
    open(os.path.join(ROOT, "profiles", "%s_ncu_igemm_traffic.csv" % tag), "w").write(open(sys.argv[4]).read())
    d["roofline"]["traffic"] = t["igemm_tc"]["dram_bytes_per_launch"]
    d["roofline"]["traffic_source"] = ("ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch (profiles/%s_traffic.json)" % tag)
peak = d["roofline"]["peak"]
out = []
w = out.append
w("# Round-1 profile summary (B200, one GPU, tiny.json training step, %d clips x 4 s)\n" % d["config"]["clips_per_gpu"])
w("Command: `python bench.py` (defaults: %d timed steps, %d warm-up).  Raw line: `profiles/%s_bench_n1.json`.\n" % (d["steps"], d["warmup"], tag))
w("* value (device-resident inputs): **%.1f clips/s**, %.3f ms/step; e2e (pinned host buffers in, loss read back): **%.1f clips/s**"
  % (d["value"], d["ms_per_step"], d["e2e"]["value"] if d.get("e2e") else float("nan")))
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]
    w("* CPU baseline (oracle port, %d threads, %s): %.2f clips/s" % (c["cores"], c["sample"], c["value"]))
w("* clocks during the timed region: SM %s / %s MHz, throttle reasons %s" % (d["clocks"].get("sm_mhz"), d["clocks"].get("sm_max_mhz"), d["clocks"].get("reasons")))
w("* launches of this library inside the timed region: %d (%d per step)" % (d["gpu_launches"], d["gpu_launches"] // d["steps"]))
if d.get("inference"):
    i = d["inference"]
    w("* inference: streaming %d streams %.3f ms/step -> **RTF %.0f**; offline %d x 10-s clips %.2f ms -> **RTF %.0f**"
      % (i["stream"]["streams"], i["stream"]["ms_per_step"], i["stream"]["rtf"], i["offline"]["clips"], i["offline"]["ms_per_batch"], i["offline"]["rtf"]))
r = d["roofline"]
w("* roofline of the dominant kernel family (`%s`): %.0f GB/s algorithmic = **%.3f** of the measured %.0f GB/s copy peak; "
  "ncu DRAM traffic per launch %s B vs %.0f B algorithmic" % (r["kernel"], r["achieved"], r["frac"], peak, r.get("traffic"), r.get("algorithmic_bytes_per_launch", 0)))
w("\n## Per-kernel CUDA-event times measured inside bench.py (average per step)\n")
w("Bytes are the ALGORITHMIC bytes of each launch (inputs once + outputs once, DESIGN.md section 4); GB/s = bytes / event time;")
w("peak = %.1f GB/s (MEASURED_PEAKS.json, measured copy bandwidth).\n" % peak)
w("| kernel family | launches/step | ms/step | share | GB/s (algorithmic) | frac of measured HBM peak | fp32-equivalent TFLOP/s |")
w("|---|---|---|---|---|---|---|")
tot = sum(v["ms_per_step"] for v in d["kernels"].values())
for k, v in d["kernels"].items():
    w("| %s | %d | %.3f | %.1f %% | %.0f | %.2f | %.1f |" % (k, v["launches_per_step"], v["ms_per_step"], 100 * v["ms_per_step"] / tot,
                                                         v["GBps"], v["GBps"] / peak, v["TFLOPs"]))
w("\nSum of kernel times %.2f ms vs %.2f ms per profiled step (launch gaps + torch optimizer/zero-grad kernels make up the rest).\n"
  % (tot, r["ms_per_step_profiled"]))
if os.path.exists(launches_path):
    rows = [x for x in csv.reader(open(launches_path)) if len(x) > 14 and x[0].isdigit() and x[12] == "gpu__time_duration.sum"]
    # the capture window spans about two steps: keep exactly one, from one front-end launch to the next
    starts = [i for i, x in enumerate(rows) if "frontend_kernel" in x[4]]
    if len(starts) >= 2:
        rows = rows[starts[0]:starts[1]]
    agg = collections.OrderedDict()
    for x in rows:
        if x[12] != "gpu__time_duration.sum":
            continue
        name = x[4].split("(")[0][:46]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(x[14].replace(",", "")) / (1e6 if x[13] in ("ns", "nsecond") else 1e3 if x[13] in ("us", "usecond") else 1.0)
    t = sum(a[1] for a in agg.values())
    w("## ncu launch list of the same command (`profiles/%s_ncu_launches.csv`)\n" % tag)
    w("`ncu --metrics gpu__time_duration.sum --clock-control none -s <3 steps> -c <2 steps> --csv python bench.py --steps 3 --warmup 3 "
      "--no-cpu-baseline --no-e2e --no-inference` (rows between two consecutive front-end launches = one training step; cold-cache, serialised: compare SHARES with the table above, not absolutes).\n")
    w("| kernel | launches | ms | share |")
    w("|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w("| `%s` | %d | %.3f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / t))
    w("\ntotal %.2f ms over %d launches" % (t, sum(a[0] for a in agg.values())))
open(os.path.join(ROOT, "profiles", "%s_summary.md" % tag), "w").write("\n".join(out) + "\n")
json.dump(d, open(os.path.join(ROOT, "profiles", "%s_bench_n1.json" % tag), "w"))
if os.path.exists(launches_path):
    open(os.path.join(ROOT, "profiles", "%s_ncu_launches.csv" % tag), "w").write(open(launches_path).read())
print("\n".join(out[:12]))
