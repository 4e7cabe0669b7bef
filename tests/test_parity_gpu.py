"""Parity of the CUDA path (through the C ABI) against the CPU oracle — the first gate.

Bit-exact: per-face clipping class / degenerate / drag-branch codes and the per-body counters.
1e-9 of the force norm: the six force/torque sums, Swet, Vsub against the oracle's F64_ACC mode (identical fp32
per-face terms, fp64 sums). The reference's own fp32 sequential sums (REF_F32_SEQ) are compared with a loose
tolerance that documents the reference's rounding noise.
"""
import numpy as np
import pytest

from stonefish_b200 import scenes
from tests.parity import REL_TOL, assert_counts_equal, max_rel_err, run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("waves", [False, True])
@pytest.mark.parametrize("currents", ["none", "uniform", "jetpipe"])
@pytest.mark.parametrize("seed", [7, 8])
def test_mixed_small_bit_exact_classes_and_forces(ctx, oracle, waves, currents, seed, surface_mode):
    sc = scenes.mixed_small(seed, waves, currents)
    gi, gc, ri, rc = run_both(ctx, oracle, sc)
    assert_counts_equal(gi, ri)
    if currents == "jetpipe":
        # exp/pow differ by <=1-2 ulp between CUDA and glibc: clipping class and degenerate bits stay exact,
        # drag-branch bits may flip with probability ~1e-8 per face; none expected at this size.
        assert np.array_equal(gc & 15, rc & 15)
        assert (gc != rc).sum() <= 1
    else:
        assert np.array_equal(gc, rc)
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)
    # every position and every clipping class is reached by this scene family
    assert set(np.unique(gi["position"])) == {0, 1, 2}
    assert (gi["class_count"].sum(axis=0) > 0).all()


def test_config2_reduced_and_full(ctx, oracle):
    for seg, st in ((100, 51), (400, 251)):
        sc = scenes.config2(seg, st)
        gi, gc, ri, rc = run_both(ctx, oracle, sc)
        assert_counts_equal(gi, ri)
        assert np.array_equal(gc, rc)
        err, where = max_rel_err(gi, ri)
        assert err <= REL_TOL, (err, where)
    assert gi["n_faces"][0] == 200000


@pytest.mark.parametrize("waves", [False, True])
def test_config3_reduced_and_full(ctx, oracle, waves, surface_mode):
    # 10k and 200k faces (direct mode: 4 and 2 sub-items per chunk), 455k faces (1777 chunks: one more than the resident warps —
    # whole chunks in a second wave of warps), 1M faces (more chunks than two waves: the lists whatever the mode)
    for seg, st in ((100, 51), (400, 251), (600, 380), (1000, 501)):
        sc = scenes.config3(seg, st, waves)
        gi, gc, ri, rc = run_both(ctx, oracle, sc)
        assert_counts_equal(gi, ri)
        assert np.array_equal(gc, rc)
        err, where = max_rel_err(gi, ri)
        assert err <= REL_TOL, (err, where)
        assert gi["wrote_buoyancy"][0] == 1
        if not waves:  # Archimedes on a flat sea: Fb = -Vsub * rho * g (SolidEntity.cpp:1705)
            assert gi["Fb"][0][2] == -gi["Vsub"][0] * sc.ocean.density * sc.ocean.gravity


def test_config1_scene_full_dispatch(ctx, oracle, surface_mode):
    sc = scenes.config1()
    gi, gc, ri, rc = run_both(ctx, oracle, sc)
    assert_counts_equal(gi, ri)
    assert np.array_equal(gc, rc)
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)
    assert set(np.unique(gi["position"])) == {0, 1, 2}


def test_reference_fp32_sequential_gap_is_rounding_noise(ctx, oracle):
    sc = scenes.config2(200, 101)
    gi, _, ri, _ = run_both(ctx, oracle, sc, debug=False, acc_mode=0)
    assert_counts_equal(gi, ri)
    err, _ = max_rel_err(gi, ri)
    assert err < 1e-4  # ~1e-7*sqrt(N) relative: the reference's own accumulation error, not ours


@pytest.mark.parametrize("unique", [False, True])
def test_config4_batch(ctx, oracle, unique):
    sc = scenes.config4(96, 50, 21, unique_meshes=unique)
    gi, gc, ri, rc = run_both(ctx, oracle, sc)
    assert_counts_equal(gi, ri)
    assert np.array_equal(gc, rc)
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)


def test_config4_many_bodies_no_debug(ctx, oracle):
    sc = scenes.config4(3000, 20, 11)  # thousands of tiny bodies: several bodies per thread block
    gi, _, ri, _ = run_both(ctx, oracle, sc, debug=False)
    assert_counts_equal(gi, ri)
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)


def test_config5_reduced(ctx, oracle, surface_mode):
    sc = scenes.config5(6, 120, 61, unique_meshes=True)
    gi, gc, ri, rc = run_both(ctx, oracle, sc)
    assert_counts_equal(gi, ri)
    assert np.array_equal(gc & 15, rc & 15)
    assert (gc != rc).sum() <= 1
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)


def test_bitwise_reproducible(ctx):
    sc = scenes.mixed_small(11, True, "jetpipe")
    st = sc.bind(ctx)
    a = ctx.compute(sc.ocean, st)
    b = ctx.compute(sc.ocean, st)
    assert a.tobytes() == b.tobytes()


def test_keep_order_equals_morton_order(ctx, oracle):
    sc = scenes.config3(100, 51, True)
    ids_m = ctx.upload_mesh(sc.meshes[0], keep_order=False)
    ids_k = ctx.upload_mesh(sc.meshes[0], keep_order=True)
    st = sc.states.copy()
    st["mesh_id"] = ids_m
    a = ctx.compute(sc.ocean, st)
    st["mesh_id"] = ids_k
    b = ctx.compute(sc.ocean, st)
    assert_counts_equal(a, b)
    err, where = max_rel_err(a, b)
    assert err <= 1e-12, (err, where)


def test_edge_cases(ctx, oracle, surface_mode):
    from stonefish_b200 import Mesh, Ocean, SfhError, new_body_states

    ocean = Ocean()
    # zero bodies is a no-op
    assert ctx.compute(ocean, new_body_states(0)).shape == (0,)
    # a mesh with no faces gives zero outputs, like mesh == nullptr in the reference (SolidEntity.cpp:1726-1733)
    empty = Mesh(np.zeros((3, 3), np.float32), np.zeros((0, 3), np.uint32))
    mid = ctx.upload_mesh(empty)
    st = new_body_states(2)
    st["mesh_id"] = mid
    st["mode"] = [1, 2]
    st["flags"] = 3
    out = ctx.compute(ocean, st)
    assert out["n_faces"].tolist() == [0, 0]
    for k in ("Fb", "Tb", "Fdq", "Tdq", "Fdf", "Tdf"):
        assert not out[k].any()
    assert out["wrote_damping"].tolist() == [1, 1]
    # errors are loud
    st["mesh_id"] = 10**6
    with pytest.raises(SfhError):
        ctx.compute(ocean, st)
    wavy = Ocean()
    wavy.fft_data = None
    o = wavy.to_struct()
    o.has_waves = 1
    st["mesh_id"] = mid
    fresh = type(ctx)(0)
    try:
        with pytest.raises(SfhError):
            fresh._check(fresh.lib.sfh_step(fresh.ctx, o, st.ctypes.data, 2, np.zeros(2, out.dtype).ctypes.data), "sfh_step")
    finally:
        fresh.close()
    with pytest.raises(SfhError):
        ctx.upload_mesh(Mesh(np.zeros((3, 3), np.float32), np.array([[0, 1, 7]], np.uint32)))


def test_full_size_properties_config4(ctx):
    """BASELINE size (50k-face hulls) through size-independent properties: counts add up, zero relative flow gives
    zero drag, and doubling the velocity scales the form-drag sum by 8 and skin friction by 2 (up to fp32 rounding)."""
    sc = scenes.config4(64, 250, 101)
    st = sc.bind(ctx)
    out = ctx.compute(sc.ocean, st)
    assert (out["n_faces"] == 50000).all()
    assert (out["class_count"][:, 7] == 50000).all()
    assert (out["n_quadratic"] + out["n_degenerate"] <= 50000).all()
    # at rest in still water: nothing
    sc.ocean.DisableCurrents()
    st0 = st.copy()
    st0["v"] = 0
    st0["omega"] = 0
    out0 = ctx.compute(sc.ocean, st0)
    assert not out0["Fdq"].any() and not out0["Fdf"].any() and out0["n_quadratic"].sum() == 0
    # scaling law (exact powers of two scale every fp32 term exactly)
    st1 = st.copy()
    st1["omega"] = 0
    a = ctx.compute(sc.ocean, st1)
    st1["v"] *= 2
    b = ctx.compute(sc.ocean, st1)
    assert np.array_equal(a["n_quadratic"], b["n_quadratic"])
    np.testing.assert_allclose(b["Fdq"], 8 * a["Fdq"], rtol=1e-6, atol=0)  # vc * |vc| * (vc.n) * A: cubic in speed
    np.testing.assert_allclose(b["Fdf"], 2 * a["Fdf"], rtol=1e-6, atol=0)


def test_graph_replay_of_small_steps_tracks_inputs(ctx, oracle):
    """Steps with <=1024 bodies are captured into a CUDA graph on the second identical-shape call and replayed afterwards
    (stonefish_b200/csrc/sfh.cu:step_host). Replays must see the new body states and ocean parameters of every step."""
    sc = scenes.mixed_small(21, True, "jetpipe")
    st = sc.bind(ctx)
    first = ctx.compute(sc.ocean, st)
    for rep in range(3):  # normal run, capture, replay: bit-identical
        assert ctx.compute(sc.ocean, st).tobytes() == first.tobytes()
    # new velocities and a retuned current, same shapes: the replayed graph must pick both up
    st2 = st.copy()
    st2["v"] *= 1.5
    sc.ocean.currents[0].setVelocity((0.4, 0.0, -0.1))
    sc.ocean.currents[1].setOutletVelocity(3.0)
    got = ctx.compute(sc.ocean, st2)
    sc2 = scenes.Scene(sc.name, sc.meshes, sc.body_mesh, sc.ocean, sc.states.copy())
    sc2.states["v"] *= 1.5
    ref, _ = oracle.OracleScene(sc2.meshes, sc2.body_mesh, sc2.ocean, sc2.states).step(1, n_threads=4)
    for k in ("class_count", "n_degenerate", "n_faces", "position"):
        assert np.array_equal(got[k], ref[k]), k
    err, where = max_rel_err(got, ref)
    assert err <= REL_TOL, (err, where)
    assert got.tobytes() != first.tobytes()


@pytest.mark.parametrize("depth,waves", [(0.05, False), (0.1, True), (4.0, False), (-6.0, False)])
def test_compound_parts_sum_like_the_reference(ctx, oracle, depth, waves, surface_mode):
    """Row a10, Compound::ComputeHydrodynamicForces (Compound.cpp:291-437): parts go through the C ABI as bodies sharing
    the compound's T_CG / velocities / AABB, the host sums them in part order (hydro.sum_compound) — against the oracle's
    restatement of the whole method, for crossing (flat and waves), inside and outside compounds, with an internal part,
    a non-buoyant part and one whose buoyancy is not assigned (stale temporaries re-added)."""
    from stonefish_b200 import meshes as M
    from stonefish_b200.hydro import compound_part_states, make_transform, new_body_states, quat_to_basis, sum_compound, transform_mul
    from stonefish_b200._abi import SFH_F_BUOYANCY, SFH_F_BUOYANT, SFH_F_DAMPING, SFH_POS_AUTO

    meshes_ = [M.icosphere(0.5, 2), M.box(1.0, 0.5, 0.2, 2), M.uv_ellipsoid(0.9, 0.18, 0.18, 24, 13)]
    ocean = scenes.Ocean(1000.0, 9.81)
    if waves:
        ocean.setWaves(scenes.wave_field())
    ocean.AddVelocityField(scenes.Uniform((0.3, -0.2, 0.05)))
    ocean.EnableCurrents()
    T_CG = make_transform(quat_to_basis(scenes.euler_zyx_quat(0.4, 0.1, -0.2)), (1.0, -2.0, depth))
    T_O = transform_mul(T_CG, make_transform(None, (0.05, 0.0, -0.02)))
    # (mesh, origin in the compound frame, external, buoyant)
    layout = [(0, (0.0, 0.0, 0.0), 1, 1), (1, (1.2, 0.0, 0.0), 1, 1), (2, (0.0, 0.9, -1.5), 1, 1), (1, (-1.0, 0.0, 0.1), 1, 0),
              (0, (0.3, 0.3, 0.2), 0, 1), (1, (0.0, -0.8, 0.0), 0, 0)]
    parts = new_body_states(len(layout))
    external, part_mesh = [], []
    lo, hi = np.full(3, 1e30), np.full(3, -1e30)
    for i, (mi, org, ext, buoy) in enumerate(layout):
        T_O_part = transform_mul(T_O, make_transform(quat_to_basis(scenes.euler_zyx_quat(0.1 * i, 0, 0)), org))
        parts["T_C"][i] = transform_mul(T_O_part, make_transform(None, (0.01, 0.0, 0.0)))
        parts["T_O_basis"][i] = T_O_part[:9]
        parts["Cd"][i], parts["Cf"][i] = (1.0 + 0.1 * i, 0.9, 1.1), (0.05, 0.06 + 0.01 * i, 0.07)
        parts["flags"][i] = SFH_F_BUOYANT if buoy else 0
        a, b = scenes.world_aabb(meshes_[mi], parts["T_C"][i])
        lo, hi = np.minimum(lo, a), np.maximum(hi, b)
        external.append(ext)
        part_mesh.append(mi)
    comp = new_body_states(1)[0]
    comp["T_CG"], comp["v"], comp["omega"] = T_CG, (0.5, 0.1, -0.2), (0.05, 0.02, 0.1)
    comp["mode"], comp["flags"] = SFH_POS_AUTO, SFH_F_DAMPING | SFH_F_BUOYANCY | SFH_F_BUOYANT
    comp["aabb_min"], comp["aabb_max"] = lo, hi
    comp["volume"], comp["surface"], comp["P_CB"] = 1.234, 5.678, (0.02, 0.0, -0.03)

    st = compound_part_states(comp, parts, external)
    ids = [ctx.upload_mesh(m) for m in meshes_]
    st["mesh_id"] = [ids[mi] for mi in part_mesh]
    got = sum_compound(ctx.compute(ocean, st), parts, external, comp, ocean)
    ref = oracle.compound(meshes_, part_mesh, parts, external, ocean, comp)
    for k in ("position", "wrote_buoyancy", "wrote_damping", "class_count", "n_degenerate", "n_quadratic", "n_skin", "n_faces"):
        assert np.array_equal(got[k], ref[k]), (k, got[k], ref[k])
    err, where = max_rel_err(np.atleast_1d(got), np.atleast_1d(ref))
    assert err <= REL_TOL, (err, where)
    expected = {0.05: 2, 0.1: 2, 4.0: 1, -6.0: 0}[depth]
    assert int(got["position"]) == expected


def test_ragged_and_chunk_boundary_meshes(ctx, oracle, surface_mode):
    """Mesh sizes around the 256-face / 256-vertex chunk limits, triangle soup (vertex-limited chunks), a single face,
    all in one step next to each other, crossing a wavy surface with Jet+Pipe currents."""
    from stonefish_b200 import meshes as M
    from stonefish_b200.hydro import Mesh, make_transform, new_body_states, quat_to_basis

    base = M.uv_ellipsoid(1.0, 0.4, 0.3, 40, 21)  # 1600 faces
    meshes_ = []
    for nf in (1, 2, 85, 86, 255, 256, 257, 511, 512, 513, 1600):
        meshes_.append(Mesh(base.positions.copy(), base.faces[:nf].copy()))
    r = M.splitmix_uniform(99, 300 * 9).reshape(-1, 3).astype(np.float32) - 0.5   # soup: 300 faces, 900 private vertices
    meshes_.append(Mesh(r, np.arange(900, dtype=np.uint32).reshape(-1, 3)))
    ocean = scenes.Ocean(1000.0, 9.81)
    ocean.setWaves(scenes.wave_field())
    ocean.AddVelocityField(scenes.Jet((0.0, -3.0, 1.0), (0.1, 1.0, 0.05), 0.3, 5.0))
    ocean.AddVelocityField(scenes.Pipe((-6.0, 0.5, 1.5), (6.0, -0.5, 2.0), 2.0, 3.0, 2.0, 2.0))
    ocean.EnableCurrents()
    n = 2 * len(meshes_)
    st = new_body_states(n)
    body_mesh = np.arange(n) % len(meshes_)
    q = scenes.random_unit_quats(5, n)
    for k in range(n):
        T = make_transform(quat_to_basis(q[k]), (0.3 * k - 3.0, 0.2 * k - 2.0, 0.05 if k < len(meshes_) else 2.5))
        st["T_CG"][k], st["T_C"][k] = T, T
        st["mode"][k] = 2 if k < len(meshes_) else 1
        st["flags"][k] = 3
        st["v"][k], st["omega"][k] = (0.5, -0.2, 0.1), (0.1, 0.05, -0.2)
    sc = scenes.Scene("ragged", meshes_, body_mesh, ocean, st)
    gi, gc, ri, rc = run_both(ctx, oracle, sc)
    for k in ("class_count", "n_degenerate", "n_faces", "position", "wrote_buoyancy", "wrote_damping"):
        assert np.array_equal(gi[k], ri[k]), k
    assert np.array_equal(gc & 15, rc & 15) and (gc != rc).sum() <= 1
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)
    assert gi["n_faces"].tolist() == [meshes_[i].n_faces for i in body_mesh]


@pytest.mark.parametrize("waves", [False, True])
def test_dry_chunk_skip_and_field_masks_are_exact(ctx, oracle, waves, surface_mode):
    """The surface path drops chunks whose box lies above the highest possible surface and skips Jet/Pipe fields that
    cannot reach a chunk's sphere. Both are conservative tests, so codes and sums must stay what the oracle says — also for a
    NON-RIGID T_C (scaled, sheared basis: the sphere bound uses row / Frobenius norms, not |R| = 1), for hulls standing
    several wave heights out of the water (most chunks dropped) and for bodies straddling the edge of a jet cone and a pipe."""
    from stonefish_b200 import meshes as M
    from stonefish_b200.hydro import Jet, Ocean, Pipe, Uniform, make_transform, new_body_states, quat_to_basis

    meshes_ = [M.box(0.6, 0.4, 6.0, 4), M.uv_ellipsoid(3.0, 0.5, 0.4, 120, 61), M.icosphere(0.5, 4)]
    ocean = Ocean(1000.0, 9.81)
    if waves:
        ocean.setWaves(scenes.wave_field())
    ocean.AddVelocityField(Uniform((0.2, -0.1, 0.0)))
    ocean.AddVelocityField(Jet((0.0, -3.0, 0.5), (0.0, 1.0, 0.0), 0.3, 5.0))
    ocean.AddVelocityField(Pipe((-6.0, 4.0, 1.0), (6.0, 4.0, 1.0), 0.8, 1.2, 2.0, 2.0))
    ocean.AddVelocityField(Uniform((0.0, 0.05, 0.01)))
    ocean.EnableCurrents()
    poses = [  # (mesh, basis, origin)
        (0, np.eye(3), (0.0, 0.0, -1.0)),                                   # a spar: 4 m of it dry
        (0, quat_to_basis(scenes.euler_zyx_quat(0.3, 0.4, -0.2)), (2.0, 1.0, -0.5)),
        (1, np.diag([1.5, 0.7, 2.0]) @ quat_to_basis(scenes.euler_zyx_quat(0.2, 0.1, 0.3)), (0.3, 0.0, 0.1)),   # scaled: crosses the jet cone
        (1, np.array([[1.0, 0.3, 0.0], [0.0, 1.0, 0.2], [0.1, 0.0, 1.0]]), (0.0, 3.6, 0.2)),                    # sheared: grazes the pipe wall
        (2, 3.0 * np.eye(3), (-3.0, -2.0, 0.8)),                                                                # uniformly scaled
        (2, np.eye(3), (0.1, -1.0, 0.45)),                                                                      # inside the jet
    ]
    st = new_body_states(len(poses))
    body_mesh = []
    for k, (mi, basis, org) in enumerate(poses):
        T = make_transform(np.asarray(basis, dtype=np.float64), org)
        st["T_CG"][k], st["T_C"][k] = make_transform(None, org), T
        st["mode"][k], st["flags"][k] = 2, 3
        st["v"][k], st["omega"][k] = (0.4, -0.2, 0.1), (0.05, 0.1, -0.02)
        body_mesh.append(mi)
    sc = scenes.Scene("sphere_bounds", meshes_, np.array(body_mesh), ocean, st)
    gi, gc, ri, rc = run_both(ctx, oracle, sc)
    assert_counts_equal(gi, ri)
    assert np.array_equal(gc & 15, rc & 15) and (gc != rc).sum() <= 1  # exp/pow: <= 1 drag flag in ~1e8
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)
    assert gi["class_count"][0][0] > 0.5 * meshes_[0].n_faces  # most of the spar is dry
    assert (gi["class_count"][:, 1:7].sum(axis=1) > 0).all()   # every body really crosses the surface


def test_zero_copy_small_steps_interleaved_with_large_ones(ctx, oracle):
    """Steps of <= 64 bodies return their results through pinned host memory and a polled sequence number; larger ones through
    copies and a stream synchronisation. Alternating them, with changing inputs, every result must equal the debug entry's
    (which uses neither graphs nor zero-copy) bit for bit, and a failing step must not desynchronise the ones after it."""
    from stonefish_b200.hydro import SfhError

    small, large = scenes.mixed_small(5, True, "uniform"), scenes.config4(96, 50, 21)
    assert small.n_bodies <= 64 < large.n_bodies
    st_s, st_l = small.bind(ctx), large.bind(ctx)
    for it in range(12):
        st_s["v"] = st_s["v"] * 0.9 + 0.01 * it
        got = ctx.compute(small.ocean, st_s)
        want, _ = ctx.compute_debug(small.ocean, st_s, small.faces_per_step())
        assert got.tobytes() == want.tobytes(), it
        if it % 3 == 0:
            st_l["omega"] *= 1.1
            got = ctx.compute(large.ocean, st_l)
            want, _ = ctx.compute_debug(large.ocean, st_l, large.faces_per_step())
            assert got.tobytes() == want.tobytes(), it
        if it == 5:
            bad = st_s.copy()
            bad["mesh_id"][0] = 10**6
            with pytest.raises(SfhError):
                ctx.compute(small.ocean, bad)


def test_queued_device_steps_each_see_their_own_ocean(ctx):
    """sfh_step_device is asynchronous: several steps with DIFFERENT ocean parameters queued back to back must not share a
    staging buffer (each reads the ocean it was called with), and an unchanged ocean must still give the same answer when its
    copy is skipped."""
    import torch

    from stonefish_b200 import BODY_STATE_DTYPE, INFO_DTYPE

    sc = scenes.config4(32, 50, 21)
    st = sc.bind(ctx)
    n = st.shape[0]
    dev = torch.device("cuda", 0)
    d_states = torch.from_numpy(st.view(np.uint8).reshape(-1).copy()).to(dev)
    speeds = [0.1, 0.7, 0.7, -0.4, 1.3, 0.0, 0.1, 0.1]
    outs = [torch.zeros(n * INFO_DTYPE.itemsize, dtype=torch.uint8, device=dev) for _ in speeds]
    want = []
    for s in speeds:  # the synchronous answers first
        sc.ocean.currents[0].setVelocity((s, -0.2, 0.05))
        want.append(ctx.compute(sc.ocean, st).copy())
    torch.cuda.synchronize(dev)
    for s, o in zip(speeds, outs):  # then the same sequence queued without any synchronisation in between
        sc.ocean.currents[0].setVelocity((s, -0.2, 0.05))
        ctx.compute_device(sc.ocean.to_struct(), d_states.data_ptr(), n, o.data_ptr())
    ctx.sync()
    torch.cuda.synchronize(dev)
    for k, (o, w) in enumerate(zip(outs, want)):
        got = np.frombuffer(o.cpu().numpy().tobytes(), dtype=INFO_DTYPE)
        assert got.tobytes() == w.tobytes(), k
    assert want[0].tobytes() != want[1].tobytes()


def test_cuda_path_against_the_references_own_compiled_code(ctx, oracle, surface_mode):
    """Directly against oracle/_ref/librefsrc.so — the reference's definitions compiled from its source text (shipped with the
    repository snapshot, built where the reference is mounted): resolved position and which outputs get assigned must be
    identical; the sums differ only by the reference's own fp32 sequential-summation noise (the oracle, which reproduces that
    library bit for bit, is what the 1e-9 contract is checked against)."""
    from oracle import pyrefsrc as rs

    if not rs.available():
        pytest.skip("oracle/_ref/librefsrc.so not built (reference checkout was not mounted at build time)")
    for sc in (scenes.mixed_small(7, True, "jetpipe"), scenes.config1(), scenes.config3(100, 51, True)):
        got = ctx.compute(sc.ocean, sc.bind(ctx))
        ref, _ = rs.RefScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states).step(4)
        for k in ("position", "wrote_buoyancy", "wrote_damping"):
            assert np.array_equal(got[k], ref[k]), (sc.name, k)
        err, where = max_rel_err(got, ref)
        assert err <= 2e-4, (sc.name, err, where)


def test_device_step_hints(ctx, oracle):
    """sfh_step_device_hinted: the caller of a device-resident step says which paths its bodies can take (sfh_step sees that in its
    host-side states); excluded paths launch nothing, results are the same, and a body that takes an excluded path after all is
    reported by the next sfh_sync instead of silently getting zero forces."""
    import torch

    from stonefish_b200 import _abi
    from stonefish_b200.hydro import BODY_STATE_DTYPE, INFO_DTYPE, SfhError

    sc = scenes.config4(16, 50, 21, unique_meshes=True)          # all INSIDE
    st = sc.bind(ctx)
    dev = torch.device("cuda", 0)
    d_st = torch.from_numpy(st.view(np.uint8).reshape(-1).copy()).to(dev)
    d_out = torch.zeros(16 * INFO_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    o = sc.ocean.to_struct()

    def run(hints):
        d_out.zero_()
        ctx.compute_device(o, d_st.data_ptr(), 16, d_out.data_ptr(), hints)
        ctx.sync()
        return d_out.cpu().numpy().view(INFO_DTYPE).copy()

    plain, hinted = run(0), run(_abi.SFH_HINT_NO_SURFACE)
    assert plain.tobytes() == hinted.tobytes() and np.abs(plain["Fdq"]).max() > 0
    ref, _ = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states).step(oracle.ORC_F64_ACC, n_threads=4)
    assert_counts_equal(hinted, ref)
    with pytest.raises(SfhError, match="excluded"):
        run(_abi.SFH_HINT_NO_SUBMERGED)                          # every body IS submerged
    again = run(_abi.SFH_HINT_NO_SURFACE)                        # the context is usable afterwards
    assert again.tobytes() == plain.tobytes()


@pytest.mark.parametrize("amplitude", [1.0, 0.05, 20.0])
def test_chunk_classifier_is_conservative(ctx, oracle, surface_mode, amplitude):
    """The surface path drops chunks it proves dry and skips the clipping code for chunks it proves wet, from the chunk's box and a LOCAL
    bound of the wave lookup. A wrong proof changes per-face classes, which must stay bit-exact. Poses that stress the bound: on and next
    to the period boundaries of both wave grids (893 m, 101 m: the reference's index quirk at i0f == 1.0f, OpenGLRealOcean.cpp:291-296,
    negative coordinates), far from the origin (fp32 spacing of the coordinates grows), any orientation, CG within a few centimetres of
    the surface; wave fields 20x smaller and 20x larger than the hull."""
    from stonefish_b200 import meshes as M
    from stonefish_b200._abi import SFH_POS_CROSSING
    from stonefish_b200.hydro import Ocean, Uniform, make_transform, new_body_states, quat_to_basis, transform_mul
    from stonefish_b200.scenes import CG2C, RAW, Scene, _fill_body, random_unit_quats, wave_field

    mesh = M.uv_ellipsoid(1.5, 0.4, 0.3, 48, 25)      # 2304 faces: nine chunks with boxes of decimetres
    ocean = Ocean(1000.0, 9.81)
    f = wave_field()
    f[0, :, :, :2] *= amplitude
    ocean.setWaves(f)
    ocean.AddVelocityField(Uniform((0.2, -0.1, 0.05)))
    ocean.EnableCurrents()
    gx, gy = 893.0, 101.0
    spots = [(0.0, 0.0), (gx, 0.0), (0.0, gy), (-gx, -gy), (gx * 3, gy * 7), (0.3, -0.2), (-0.2, 0.4), (gy + 0.1, gy - 0.1), (-gy * 2 + 0.05, 0.0),
             (gx - 0.2, -gx + 0.2), (gx / 256 * 0.5, gy / 256 * 0.5), (-gx / 256 * 0.5, -gy / 256 * 1.5), (1.0e4, -7.3e3), (-2.5e4 + 0.7, 3.1e4),
             (123.456, -654.321), (50.5, 50.5)]
    n = 3 * len(spots)
    st = new_body_states(n)
    quats = random_unit_quats(91, n)
    u = M.splitmix_uniform(92, n)
    for k in range(n):
        x, y = spots[k % len(spots)]
        tilt = quats[k] if k >= len(spots) else scenes.euler_zyx_quat(0.7 * k, 0.03, -0.05)
        z = (-0.25 + 0.5 * u[k]) * (1.0 if amplitude < 5 else 8.0)
        T_CG = make_transform(quat_to_basis(tilt), (x, y, z))
        _fill_body(st, k, mesh, T_CG, transform_mul(T_CG, CG2C), (0.5, 0.1, -0.2), (0.05, -0.02, 0.1), SFH_POS_CROSSING, RAW)
    sc = Scene("classifier_stress", [mesh], np.zeros(n, dtype=np.int64), ocean, st)
    gi, gc, ri, rc = run_both(ctx, oracle, sc)
    assert np.array_equal(gc, rc)
    assert_counts_equal(gi, ri)
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)
    cls = gi["class_count"].sum(axis=0)
    assert cls[0] > 0 and cls[7] > 0 and cls[1:7].sum() > 0     # dry, wet and cut faces all occur


def test_step_without_surface_items_between_surface_steps(ctx, oracle, surface_mode):
    """AUTO bodies that all turn out INSIDE leave the surface path with NO items although the host had to allow for some: the
    consumers of that step must not walk the lists (or the control words) an earlier step left behind, and the next surface step must
    not see anything of the empty one. Both ways through the surface path, host-side and device-resident states."""
    import torch

    from stonefish_b200._abi import SFH_POS_AUTO
    from stonefish_b200.hydro import INFO_DTYPE

    crossing = scenes.mixed_small(5, True, "none")
    deep = scenes.config4(12, 50, 21, unique_meshes=True)         # far below the surface ...
    deep.ocean = crossing.ocean
    deep.states["mode"] = SFH_POS_AUTO                             # ... but the gate decides that on the device
    for round_ in range(2):
        for sc in (crossing, deep, crossing):
            gi, gc, ri, rc = run_both(ctx, oracle, sc)
            assert_counts_equal(gi, ri)
            err, where = max_rel_err(gi, ri)
            assert err <= REL_TOL, (sc.name, round_, err, where)
        assert (gi["position"] == 2).any()
    gi, _, ri, _ = run_both(ctx, oracle, deep)
    assert (gi["position"] == 1).all() and np.abs(gi["Fdq"]).max() > 0
    # device-resident states: the host knows nothing about the bodies
    st = deep.bind(ctx)
    dev = torch.device("cuda", 0)
    d_st = torch.from_numpy(st.view(np.uint8).reshape(-1).copy()).to(dev)
    d_out = torch.zeros(len(st) * INFO_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    run_both(ctx, oracle, crossing)
    ctx.compute_device(deep.ocean.to_struct(), d_st.data_ptr(), len(st), d_out.data_ptr())
    ctx.sync()
    got = d_out.cpu().numpy().view(INFO_DTYPE)
    assert_counts_equal(got, ri)
    err, where = max_rel_err(got, ri)
    assert err <= REL_TOL, (err, where)


def test_pinned_caller_buffers_take_the_same_results_every_way(ctx, oracle):
    """sfh_step chooses between a replayed graph (copies through the context's staging buffers) and launches one by one with copies
    straight from / to the caller's buffers — by body count, face count and whether the caller's buffers are pinned. Same bytes either
    way, and equal to the oracle: a mixed scene of 300 bodies (> 256: launched one by one when pinned), one of 24 bodies with 4.2 M
    faces (> 2.5 M faces), and a small one (graph in both cases)."""
    import torch

    from stonefish_b200.hydro import BODY_STATE_DTYPE, INFO_DTYPE

    big = scenes.config4(24, 700, 126, unique_meshes=False)      # 24 x 175k faces, all INSIDE
    many = scenes.config4(300, 24, 22, unique_meshes=False)      # 300 x 1008 faces
    small = scenes.mixed_small(9, True, "uniform")
    for sc in (big, many, small):
        st = sc.bind(ctx)
        n = len(st)
        pageable = ctx.compute(sc.ocean, st)
        t_st = torch.from_numpy(st.view(np.uint8).reshape(-1).copy()).pin_memory()
        t_out = torch.zeros(n * INFO_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        for _ in range(3):                                        # the third call of a shape would replay a graph
            t_out.zero_()
            pinned = ctx.compute(sc.ocean, t_st.numpy().view(BODY_STATE_DTYPE), out=t_out.numpy().view(INFO_DTYPE))
            assert pinned.tobytes() == pageable.tobytes(), sc.name
        ref, _ = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states).step(oracle.ORC_F64_ACC, n_threads=8)
        assert_counts_equal(pageable, ref)
        err, where = max_rel_err(pageable, ref)
        assert err <= REL_TOL, (sc.name, err, where)
