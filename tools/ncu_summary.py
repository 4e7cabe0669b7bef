#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the numbers DESIGN.md / profiles/ quote: duration, DRAM traffic, pipe
utilisation, issue efficiency, stall reasons, dynamic instruction mix and the hottest SASS lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [faces_per_launch] [--traffic-json profiles/ncu_traffic_rNN.json] [--report-name profiles/x.txt] [--skip N]
(--traffic-json also writes the per-face DRAM traffic and the pipe utilisations bench.py quotes as roofline.traffic / .secondary)"""
import collections, csv, io, json, re, subprocess, sys

argv = list(sys.argv)
traffic_json = None
if "--traffic-json" in argv:
    i = argv.index("--traffic-json")
    traffic_json = argv[i + 1]
    del argv[i:i + 2]
report_name = "profiles/ncu_k_submerged_r02.txt"
if "--report-name" in argv:
    i = argv.index("--report-name")
    report_name = argv[i + 1]
    del argv[i:i + 2]
skip = []
if "--skip" in argv:  # reports with several kernels: summarise launch number N of the report
    i = argv.index("--skip")
    skip = ["--launch-skip", argv[i + 1], "--launch-count", "1"]
    del argv[i:i + 2]
rep = argv[1]
faces = float(argv[2]) if len(argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep] + skip + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k): return m.get(k, ("n/a", ""))
print("kernel:", g("Kernel Name")[0], " grid", g("launch__grid_size")[0], " block", g("launch__block_size")[0])
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second"]:
    v, u = g(k)
    print(f"  {k:82s} {v:>16s} {u}")
if faces:
    inst = float(g("smsp__inst_executed.sum")[0].replace(",", ""))
    print(f"  warp instructions per 64-face warp iteration: {inst / (faces / 64):.1f}   (per face-lane: {inst / (faces / 32):.1f})")
    rd = float(g("dram__bytes_read.sum")[0].replace(",", "")); u = g("dram__bytes_read.sum")[1]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    print(f"  DRAM bytes read per face: {rd * mult / faces:.2f}")
if traffic_json and faces:
    def num(k):
        v, u = g(k)
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "%": 0.01}.get(u, 1)
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    json.dump({"kernel": g("Kernel Name")[0],
               "command": "ncu --set full --clock-control none --import-source on -k regex:k_submerged -s 3 -c 1 python bench.py --bodies 1024 --steps 3 --warmup 1 --no-cpu",
               "faces_per_launch": int(faces), "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_face": round((rd + wr) / faces, 5), "duration_us": num("gpu__time_duration.sum"),
               "report": report_name,
               "secondary": {"issue_slot_utilisation": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active"), 3),
                             "fma_pipe": round(num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 3),
                             "xu_pipe": round(num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), 3),
                             "alu_pipe": round(num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 3),
                             "fp64_pipe": round(num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 3),
                             "dram_throughput_pct_of_peak": round(num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 3)}},
              open(traffic_json, "w"), indent=1)
print("stall reasons (warps stalled per issue-active cycle):")
st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(v) for h, (v, u) in m.items() if "issue_stalled_" in h and h.endswith("per_issue_active.ratio")}
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]:
    print(f"  {k:24s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep] + skip + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = [r for r in rows[2:] if len(r) >= len(rows[1]) and r[rows[1].index("Instructions Executed")].strip().isdigit()]  # (functions that are not inlined start sections of their own)
ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
by, samp = collections.Counter(), collections.Counter()
for r in data:
    op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
    by[op] += int(r[ia]); samp[op] += int(r[isamp])
tot, ts = sum(by.values()), sum(samp.values())
print("dynamic instruction mix (share of executed warp instructions | share of stall samples):")
for op, c in by.most_common(22):
    per = f"{c / (faces / 64):7.1f}/iter" if faces else ""
    print(f"  {op:10s} {100 * c / tot:5.1f}% {per}  | {100 * samp[op] / ts:5.1f}%")
print("hottest SASS lines by stall samples:")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:12]):
    print(f"  {100 * int(data[i][isamp]) / ts:5.1f}%  {data[i][isrc].strip()[:90]}")
