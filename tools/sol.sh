#!/bin/bash
# Speed-of-light probe of k_submerged (task 4): the production kernel next to builds that keep ONLY the mandated
# fp32 chain (-DSFH_SOL: no fp32->fp64 conversions, no fp64 adds, no counters, no guards), scalar-pair and packed f32x2 carriers.
# Build the variants first (on the CPU box):
#   cd stonefish_b200/csrc; for v in "sol:-DSFH_SOL" "solx2:-DSFH_SOL -DSFH_PACKED_F32X2" "x2:-DSFH_PACKED_F32X2" \
#       "sol9:-DSFH_SOL -DSFH_SUB_MIN_BLOCKS=9" "solx29:-DSFH_SOL -DSFH_PACKED_F32X2 -DSFH_SUB_MIN_BLOCKS=9"; do
#     make OUT=../../ab/libsfh_${v%%:*}.so INFO=../../ab/ptxas_${v%%:*}.txt EXTRA="${v#*:}"; done
# then on the GPU box:  tools/sol.sh  ->  gpurun_out/sol.txt (+ one ncu capture per build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for lib in stonefish_b200/libsfh.so ab/libsfh_x2.so ab/libsfh_sol.so ab/libsfh_sol9.so ab/libsfh_solx2.so ab/libsfh_solx29.so; do
  [ -f $lib ] || continue
  for rep in 1 2; do
    SFH_LIB=$lib python bench.py --workload cfg4 --bodies 1024 --no-cpu --no-extra --no-validate --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib  kernel_ms %.4f  -> %.1f G faces/s  (%.3f of the HBM roofline)' % (r['kernel_ms'], 51.2e6/r['kernel_ms']/1e6, r['frac']))"
  done
done
for tag in prod:stonefish_b200/libsfh.so sol:ab/libsfh_sol.so solx2:ab/libsfh_solx2.so; do
  name=${tag%%:*}; lib=${tag#*:}
  [ -f $lib ] || continue
  SFH_LIB=$lib ncu --set full --clock-control none --import-source on -k regex:k_submerged -s 3 -c 1 -o gpurun_out/sol_$name -f \
      python bench.py --workload cfg4 --bodies 1024 --steps 3 --warmup 1 --no-cpu --no-extra --no-validate > /dev/null 2>&1
  echo "==== $name ($lib)"
  python tools/ncu_summary.py gpurun_out/sol_$name.ncu-rep 51200000 | head -48
done
} | tee gpurun_out/sol.txt
