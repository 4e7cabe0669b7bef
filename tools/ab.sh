#!/bin/bash
# A/B of kernel builds on the GPU box: tools/ab.sh ab/libsfh_X.so ab/libsfh_Y.so ...  -> gpurun_out/ab.txt
# (bench.py loads the library named by SFH_LIB; every line is a device-timed bench.py run, kernel ms from its roofline object)
# AB_WORKLOADS (default "cfg4 cfg3-flat cfg5"), AB_ARGS (extra bench.py arguments, e.g. "--bodies 1024"), AB_ABLATE=1 adds
# tools/ablate_surface.py per library.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
  for w in ${AB_WORKLOADS:-cfg4 cfg3-flat cfg5}; do
    SFH_LIB=$lib python bench.py --workload $w --no-cpu --steps 30 --warmup 5 $AB_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib $w value %.2f G  ms/step %.4f  kernel_ms %.4f  frac %.3f' % (d['value']/1e9, d['ms_per_step'], r.get('kernel_ms',0), r['frac']))"
  done
done
done | tee gpurun_out/ab.txt
if [ "${AB_ABLATE:-0}" = 1 ]; then
  for lib in "$@"; do echo "== $lib"; SFH_LIB=$lib python tools/ablate_surface.py; done | tee -a gpurun_out/ab.txt
fi
