#!/bin/bash
# Sustained-clock evidence for the headline kernel: thousands of back-to-back steps (several seconds of continuous kernel time)
# with every nvidia-smi sample of the timed region kept. k_submerged is issue-bound, so its rate follows the SM clock: if the
# clock sags under sustained load the headline measured over 20 steps (0.03 s, burst clock) overstates a long simulation.
#   tools/sustain.sh [steps] [out.json]     (run on the GPU box: gpurun -- 'tools/sustain.sh 3000 gpurun_out/sustain.json')
cd "$(dirname "$0")/.."
STEPS=${1:-3000}
OUT=${2:-gpurun_out/sustain.json}
mkdir -p "$(dirname "$OUT")"
python bench.py --steps "$STEPS" --warmup 5 --no-cpu --no-extra --no-validate --clock-trace > "$OUT.tmp" || exit 1
python - "$OUT.tmp" "$OUT" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c = d["clocks"]
tr = c.get("trace", [])
out = {"what": "bench.py --steps N --no-cpu --no-extra --clock-trace (cfg4, 4096 x 50k-face hulls)", "steps": d["steps"],
       "timed_region_s": d["ms_per_step"] * d["steps"] / 1e3, "value": d["value"], "ms_per_step": d["ms_per_step"],
       "e2e_value": d["e2e"]["value"], "kernel_ms": d["roofline"]["kernel_ms"], "roofline_frac": d["roofline"]["frac"],
       "sm_mhz_median": c.get("sm_mhz"), "sm_mhz_min": c.get("sm_mhz_min"), "sm_max_mhz": c.get("sm_max_mhz"),
       "power_w_max": c.get("power_w_max"), "reasons": c.get("reasons"), "samples": c.get("samples"),
       "trace_t_s__sm_mhz__power_w": tr}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "trace_t_s__sm_mhz__power_w"}))
PY
rm -f "$OUT.tmp"
