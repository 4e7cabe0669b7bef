#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck of every kernel (prologue/scan, submerged and its aerodynamic variant,
# surface flat and on a device-generated wave field with Jet/Pipe currents, epilogue, wave producer, cables) through the C++
# host mirror demo on a small mixed scene. Run on the GPU box:
#   bash tools/sanitize.sh > gpurun_out/sanitize.log
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import sys
sys.path.insert(0, ".")
from stonefish_b200 import scenes
from tests.test_host_mirror_gpu import write_scene
sc = scenes.mixed_small(3, False, "jetpipe")
sc.states["mode"] = 3            # AUTO, as the host mirror sends it
write_scene("/tmp/sanitize_scene.bin", sc)
print("scene written:", sc.n_bodies, "bodies", sc.faces_per_step(), "faces")
PY
# both ways through the surface path: direct mode (what a scene this small takes) and, with SFH_NO_DIRECT=1, the lists
# (k_classify, wet / clip roles, segments)
for nodirect in 0 1; do
for tool in memcheck racecheck synccheck; do
  echo "== SFH_NO_DIRECT=$nodirect compute-sanitizer --tool $tool"
  SFH_NO_DIRECT=$nodirect compute-sanitizer --tool $tool --error-exitcode 9 stonefish_b200/host/demo_fluid_dynamics /tmp/sanitize_scene.bin /tmp/sanitize_out.bin 2 /tmp/sanitize_aero.bin 1.0 /tmp/sanitize_waves.bin /tmp/sanitize_cable.bin 2>&1 | tail -4
done
done
echo "== compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py tests/test_real_meshes_gpu.py tests/test_aero_gpu.py -m gpu -k 'not full and not large'"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_real_meshes_gpu.py tests/test_aero_gpu.py -m gpu -q -x -k 'not full and not large' 2>&1 | tail -4
echo "== compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -k config3   (10k .. 1M-face hulls: sub-items, second wave of direct mode, lists)"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k config3 2>&1 | tail -4
