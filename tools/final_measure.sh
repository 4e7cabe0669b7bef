set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_submerged -s 3 -c 1 -o gpurun_out/sub_final -f python bench.py --bodies 1024 --steps 3 --warmup 1 --no-cpu > /dev/null 2>&1
# roofline.traffic / roofline.secondary of bench.py come from this capture
python tools/ncu_summary.py gpurun_out/sub_final.ncu-rep 51200000 --traffic-json profiles/ncu_traffic_r01.json > profiles/ncu_k_submerged_r01.txt
cp profiles/ncu_traffic_r01.json profiles/ncu_k_submerged_r01.txt gpurun_out/
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for w in cfg4-shared cfg2 cfg3 cfg3-flat cfg5; do python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
python tools/bench_aero.py > gpurun_out/aero_uniform.json 2>/dev/null
python tools/bench_aero.py --wind jetpipe --bodies 256 > gpurun_out/aero_jetpipe.json 2>/dev/null
python tools/bench_wavegen.py > gpurun_out/wavegen.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_surface -s 2 -c 1 -o gpurun_out/surf_final -f python bench.py --workload cfg5 --bodies 16 --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
tail -c 600 gpurun_out/bench_default.json
