python - <<'PY'
import sys
sys.path.insert(0, ".")
from stonefish_b200 import scenes
from tests.test_host_mirror_gpu import write_scene
sc = scenes.config1(); write_scene("/tmp/scene1.bin", sc)
sc = scenes.config2(50, 51); sc.states["mode"] = 3; sc.states["flags"] = 15
import numpy as np
from stonefish_b200 import scenes as S
lo, hi = S.world_aabb(sc.meshes[0], sc.states["T_C"][0]); sc.states["aabb_min"][0] = lo; sc.states["aabb_max"][0] = hi
write_scene("/tmp/scene2.bin", sc)
PY
stonefish_b200/host/demo_fluid_dynamics /tmp/scene1.bin /tmp/o1.bin 2000
stonefish_b200/host/demo_fluid_dynamics /tmp/scene2.bin /tmp/o2.bin 2000
