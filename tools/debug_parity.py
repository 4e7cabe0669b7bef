import sys, numpy as np
sys.path.insert(0, '.')
from stonefish_b200 import HydroContext, scenes
from oracle import pyoracle as po
from tests.parity import run_both
ctx = HydroContext(0)
for args in [(8, False, "none"), (8, False, "uniform"), (8, True, "none")]:
    sc = scenes.mixed_small(*args)
    gi, gc, ri, rc = run_both(ctx, po, sc)
    print("==", args)
    base = np.concatenate([[0], np.cumsum([sc.meshes[i].n_faces for i in sc.body_mesh])])
    for b in range(sc.n_bodies):
        same = all(np.array_equal(gi[k][b], ri[k][b]) for k in ("class_count","n_degenerate","n_quadratic","n_skin","n_faces","position"))
        cm = (gc[base[b]:base[b+1]] != rc[base[b]:base[b+1]]).sum()
        if not same or cm:
            print(b, "mode", sc.states["mode"][b], "flags", sc.states["flags"][b], "pos", gi["position"][b], ri["position"][b])
            print("   gpu", gi["class_count"][b], gi["n_degenerate"][b], gi["n_quadratic"][b], gi["n_skin"][b], gi["n_faces"][b])
            print("   ref", ri["class_count"][b], ri["n_degenerate"][b], ri["n_quadratic"][b], ri["n_skin"][b], ri["n_faces"][b])
            print("   code mismatches", cm)
