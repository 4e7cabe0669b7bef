#!/bin/bash
# Round-2 measurement pass on ONE GPU (run on the GPU box): tests, the default bench line (all BASELINE configs + validation), the
# reference arm, the sustained-clock run, the ncu launch list of the bench command and full captures of the dominant kernels.
# Artefacts land in gpurun_out/ (copy the summaries to profiles/ afterwards: see profiles/README.md).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default_r02.json 2> gpurun_out/bench_default_r02.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err
tools/sustain.sh 3000 gpurun_out/sustain_r02.json > /dev/null
python tools/latency_small.py > gpurun_out/latency_small_r02.json 2>/dev/null
SFH_PROFILE_ROLES=1 python tools/ablate_surface.py > gpurun_out/ablate_surface_r02.txt 2>&1
# launch list of the bench command (per-launch times are cold-cache and serialised: the SHARE of the step is what must agree)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-validate > gpurun_out/launches_run.log 2>&1
# full captures: k_submerged on cfg4 (1024 bodies), the wet and the clip role of k_surface_lists and k_classify on cfg5 (16 bodies)
ncu --set full --clock-control none --import-source on -k regex:k_submerged -s 3 -c 1 -o gpurun_out/sub_r02 -f \
    python bench.py --workload cfg4 --bodies 1024 --steps 3 --warmup 1 --no-cpu --no-validate > /dev/null 2>&1
SFH_PROFILE_ROLES=1 ncu --set full --clock-control none --import-source on -k regex:"k_surface_lists|k_classify" -s 10 -c 3 -o gpurun_out/surf_r02 -f \
    python bench.py --workload cfg5 --bodies 16 --steps 2 --warmup 3 --no-cpu --no-validate > /dev/null 2>&1
tail -c 400 gpurun_out/bench_default_r02.json
