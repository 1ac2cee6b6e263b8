#!/usr/bin/env python
"""Write the per-round profile artefacts under profiles/ from what tools/prof_step.sh left in gpurun_out/.

    python tools/make_profile_summary.py <tag> [<n2 bench json> ...]

Inputs  (gpurun_out/): bench_<tag>.json / .detail, launches_<tag>.csv, full_{igemm,wgrad,dw,misc}_<tag>.csv
Outputs (profiles/):   <tag>_bench_n1.json, <tag>_bench_detail.txt, <tag>_ncu_launches.csv, <tag>_ncu_full_<group>.csv (selected
                       columns of the raw page, one row per launch), <tag>_traffic.json (DRAM bytes per launch of the dominant
                       kernel family: bench.py's roofline.traffic), <tag>_summary.md
"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
d = json.load(open(os.path.join(G, "bench_%s.json" % tag)))
peak = d["roofline"]["peak"]

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def to_ms(v, unit):
    return v / (1e6 if unit.startswith("n") else 1e3 if unit.startswith("u") else 1.0 if unit.startswith("m") else 1e-3)


def to_bytes(v, unit):
    u = unit.lower()
    return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)


def load_full(group):
    path = os.path.join(G, "full_%s_%s.csv" % (group, tag))
    if not os.path.exists(path):
        return []
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(h, i) for h, i in idx.items() if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
    out = []
    keep = [k for k in KEEP if k in idx]
    with open(os.path.join(P, "%s_ncu_full_%s.csv" % (tag, group)), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(keep + ["top stalls (warps per issue)"])
        w.writerow([units[idx[k]] for k in keep] + [""])
        for r in rows[2:]:
            st = sorted(((num(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]) for h, i in stall_cols), reverse=True)
            st = [(v, n) for v, n in st if n != "selected" and v == v][:4]
            stall_txt = ", ".join("%s %.2f" % (n, v) for v, n in st)
            w.writerow([r[idx[k]] for k in keep] + [stall_txt])
            g = lambda k: num(r[idx[k]]) if k in idx else float("nan")
            out.append(dict(name=re.sub(r"^void |tru::|<unnamed>::|unnamed>::|\(.*$", "", r[idx["Kernel Name"]]),
                            ms=to_ms(g("gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]]),
                            rd=to_bytes(g("dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]]),
                            wr=to_bytes(g("dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]]),
                            dram=g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                            tensor=g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                            fma=g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                            issue=g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                            smem=g("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed"),
                            grid=r[idx["launch__grid_size"]], regs=r[idx["launch__registers_per_thread"]], stalls=stall_txt))
    return out


os.makedirs(P, exist_ok=True)
json.dump(d, open(os.path.join(P, "%s_bench_n1.json" % tag), "w"))
if os.path.exists(os.path.join(G, "bench_%s.detail" % tag)):
    lines = [ln for ln in open(os.path.join(G, "bench_%s.detail" % tag)) if ln.startswith("#")]
    open(os.path.join(P, "%s_bench_detail.txt" % tag), "w").writelines(lines)
full = {g: load_full(g) for g in ("igemm", "wgrad", "dw", "misc")}

# DRAM traffic of the dominant family, per launch SCOPE (what roofline.achieved is averaged over)
traffic = {}
fam_of = {"igemm": "igemm_tc", "wgrad": "wgrad_stream"}
for g, fam in fam_of.items():
    if full[g] and fam in d["kernels"]:
        n = int(d["kernels"][fam]["launches_per_step"])
        rd, wr = sum(k["rd"] for k in full[g]), sum(k["wr"] for k in full[g])
        traffic[fam] = {"dram_bytes_per_launch": int((rd + wr) / n), "launches": n, "kernels": len(full[g]),
                        "dram_read_bytes_step": rd, "dram_write_bytes_step": wr, "ncu_time_ms_step": sum(k["ms"] for k in full[g]),
                        "command": "tools/prof_step.sh %s (ncu --set full --clock-control none -k regex:tc_%s, one training step)" % (tag, g)}
if traffic:
    json.dump(traffic, open(os.path.join(P, "%s_traffic.json" % tag), "w"), indent=1)
    top = d["roofline"]["kernel"]
    if top in traffic:
        d["roofline"]["traffic"] = traffic[top]["dram_bytes_per_launch"]

out = []
w = out.append
w("# Profile summary %s (B200, one GPU, tiny.json training step, %d clips x 4 s)\n" % (tag, d["config"]["clips_per_gpu"]))
w("Command: `tools/prof_step.sh %s` = `python bench.py` (defaults: %d timed steps, %d warm-up) + ncu passes of "
  "`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-inference --no-stock-gpu`.  Raw line: `profiles/%s_bench_n1.json`.\n"
  % (tag, d["steps"], d["warmup"], tag))
w("* value (device-resident inputs): **%.1f clips/s**, %.3f ms/step; e2e (pinned host buffers in, loss read back every step): **%.1f clips/s**"
  % (d["value"], d["ms_per_step"], d["e2e"]["value"] if d.get("e2e") else float("nan")))
ws = d["roofline"].get("whole_step")
if ws:
    w("* whole step against SURVEY 8(d)'s algorithmic bytes (%.1f GB): %.0f GB/s = **%.3f** of the measured %.0f GB/s; bytes the kernels "
      "actually move per step (their own algorithmic bytes): %.1f GB" % (ws["algorithmic_bytes"] / 1e9, ws["achieved"], ws["frac"], peak,
                                                                         ws["bytes_moved_by_kernels"] / 1e9))
if d.get("stock_gpu"):
    s = d["stock_gpu"]
    w("* stock PyTorch on the same GPU (oracle `.cuda()`: cuDNN / cuBLAS / cuFFT, same 32 clips): fp32 **%.1f clips/s** (%.1f ms/step), "
      "TF32 allowed %.1f clips/s (%.1f ms/step); stages fp32: %s" % (s["fp32"]["clips_per_sec"], s["fp32"]["ms_per_step"],
                                                                   s["tf32"]["clips_per_sec"], s["tf32"]["ms_per_step"], s["fp32"]["stages_ms"]))
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]
    w("* CPU baseline (oracle port, %d threads, %s): %.2f clips/s; one thread: %s clips/s" % (c["cores"], c["sample"], c["value"], c.get("value_1thread")))
w("* clocks during the timed region: SM %s / %s MHz, throttle reasons %s" % (d["clocks"].get("sm_mhz"), d["clocks"].get("sm_max_mhz"), d["clocks"].get("reasons")))
w("* launches of this library inside the timed region: %d (%d per step)" % (d["gpu_launches"], d["gpu_launches"] // d["steps"]))
if d.get("inference"):
    i = d["inference"]
    if "stream" in i:
        s = i["stream"]
        w("* streaming (configs[3]): %d streams, %.3f ms per 8-ms hop device-resident -> **RTF %.0f**; with the step's frames copied in from pinned "
          "host memory and its audio copied back: %.3f ms -> RTF %.0f; %.3f of the HBM roofline on %.2f GB algorithmic bytes per step"
          % (s["streams"], s["ms_per_step"], s["rtf"], s["e2e"]["ms_per_step"], s["e2e"]["rtf"], s["roofline"]["frac"],
             s["roofline"]["algorithmic_bytes_per_step"] / 1e9))
        if "pipelined_ms_per_step" in s["e2e"]:
            w("  * the same copies on side streams, one hop ahead / behind (`util.stream_host_frames`): %.3f ms per hop -> RTF %.0f"
              % (s["e2e"]["pipelined_ms_per_step"], s["e2e"]["pipelined_rtf"]))
        g = s.get("cuda_graph")
        if g and "ms_per_step" in g:
            w("  * step replayed as ONE CUDA graph (`StreamingDenoiser(cuda_graph=True)`): %.3f ms device-resident (RTF %.0f), %.3f ms with "
              "serial copies, %.3f ms with pipelined copies (RTF %.0f)" % (g["ms_per_step"], g["rtf"], g["e2e_ms_per_step"],
                                                                          g["e2e_pipelined_ms_per_step"], g["e2e_pipelined_rtf"]))
    if "offline" in i:
        o = i["offline"]
        w("* offline (configs[4]): %d x 10-s clips per GPU on %d GPU(s), host buffers in and out: %.3f s -> **RTF %.0f** (%.0f per GPU); "
          "device time per batch of %d: %.2f ms = %.3f of the HBM roofline" % (o["clips_per_gpu"], o["gpus"], o["seconds"], o["rtf"],
                                                                             o["rtf_per_gpu"], o["batch"], o["ms_per_batch_device"], o["roofline"]["frac"]))
    if "latency" in i:
        l = i["latency"]
        w("* latency (configs[0]): one 4-s clip, host to host: GPU %.2f ms (RTF %.0f), CPU oracle %s ms on %s threads / %s ms on one; "
          "one stream, one frame: %.3f ms per 8-ms hop" % (l["gpu_ms"], l["gpu_rtf"], l.get("cpu_ms"), l.get("cpu_threads"), l.get("cpu_1thread_ms"),
                                                           l["gpu_single_frame_step_ms"]))
        if "gpu_cuda_graph_ms" in l:
            w("  * as captured CUDA graphs: the 4-s clip (`util.GraphedDenoise`) %s ms, the single-frame step %s ms"
              % (l["gpu_cuda_graph_ms"], l.get("gpu_single_frame_step_cuda_graph_ms")))
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
w("\nSum of kernel times %.2f ms vs %.2f ms per profiled step (launch gaps + torch's zero-grad kernels make up the rest).  Per launch shape: "
  "`profiles/%s_bench_detail.txt`.\n" % (tot, r["ms_per_step_profiled"], tag))

lp = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, "%s_ncu_launches.csv" % tag))
    rows = [x for x in csv.reader(open(lp)) if len(x) > 14 and x[0].isdigit() and x[12] == "gpu__time_duration.sum"]
    starts = [i for i, x in enumerate(rows) if "frontend_kernel" in x[4]]
    if len(starts) >= 2:
        rows = rows[starts[0]:starts[1]]
    agg = collections.OrderedDict()
    for x in rows:
        name = re.sub(r"^void |tru::|<unnamed>::", "", x[4]).split("(")[0][:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += to_ms(num(x[14]), x[13])
    t = sum(a[1] for a in agg.values())
    w("## ncu launch list of one training step (`profiles/%s_ncu_launches.csv`)\n" % tag)
    w("`ncu --metrics gpu__time_duration.sum --clock-control none` (rows between two consecutive front-end launches = one training step; "
      "cold-cache, serialised: compare SHARES with the table above, not absolutes).\n")
    w("| kernel | launches | ms | share |")
    w("|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w("| `%s` | %d | %.3f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / t))
    w("\ntotal %.2f ms over %d launches\n" % (t, sum(a[0] for a in agg.values())))

w("## ncu `--set full` of every launch of one step (`profiles/%s_ncu_full_<group>.csv`: one row per launch)\n" % tag)
w("Aggregated per kernel (template instance); time-weighted averages.  DRAM % / tensor % / issue % are of the peak sustained rate; "
  "stalls = warps per issue slot of the launch with the largest share of the group's time.\n")
w("| kernel | launches | ms (ncu) | DRAM read+write GB | DRAM % | tensor pipe % | FMA pipe % | issue active % | smem wavefronts % | regs | top stalls (longest launch) |")
w("|---|---|---|---|---|---|---|---|---|---|---|")
for g in ("igemm", "wgrad", "dw", "misc"):
    by = collections.OrderedDict()
    for k in full[g]:
        by.setdefault(k["name"], []).append(k)
    for name, ks in sorted(by.items(), key=lambda kv: -sum(k["ms"] for k in kv[1])):
        ms = sum(k["ms"] for k in ks)
        wavg = lambda f: sum(k[f] * k["ms"] for k in ks if k[f] == k[f]) / max(ms, 1e-12)
        big = max(ks, key=lambda k: k["ms"])
        w("| `%s` | %d | %.3f | %.2f | %.0f | %.0f | %.0f | %.0f | %.0f | %s | %s |" % (
            name[:58], len(ks), ms, sum(k["rd"] + k["wr"] for k in ks) / 1e9, wavg("dram"), wavg("tensor"), wavg("fma"), wavg("issue"),
            wavg("smem"), big["regs"], big["stalls"]))
if traffic:
    w("\nDRAM traffic per launch scope (`profiles/%s_traffic.json`): " % tag + "; ".join(
        "%s %.3f GB measured vs %.3f GB algorithmic" % (f, t["dram_bytes_per_launch"] / 1e9,
                                                        d["kernels"][f]["GBps"] * d["kernels"][f]["ms_per_step"] / d["kernels"][f]["launches_per_step"] / 1e3)
        for f, t in traffic.items()))
for extra in sys.argv[2:]:
    if os.path.exists(extra):
        e = json.load(open(extra))
        shutil.copy(extra, os.path.join(P, "%s_bench_n%d.json" % (tag, e["n_gpus"])))
        w("\n* %d GPUs (`profiles/%s_bench_n%d.json`): weak scaling %.1f clips/s (%.3f ms/step), e2e %.1f; strong scaling (global batch 256): %s; "
          "data-parallel check: %s; offline inference: %s" % (e["n_gpus"], tag, e["n_gpus"], e["value"], e["ms_per_step"], e["e2e"]["value"],
                                                              e.get("strong_scaling"), e.get("dp_check"),
                                                              {k: e["inference"]["offline"][k] for k in ("rtf", "rtf_per_gpu", "seconds")} if e.get("inference") else None))
open(os.path.join(P, "%s_summary.md" % tag), "w").write("\n".join(out) + "\n")
json.dump(d, open(os.path.join(P, "%s_bench_n1.json" % tag), "w"))
print("\n".join(out[:14]))
