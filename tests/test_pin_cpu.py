"""PINS: the oracle against the reference's OWN code for this path.

The reference cannot be built as a whole here (Windows + OptiX 7.7 + cuFFT), but the sources that define the per-ray
arithmetic can: `make -C oracle ref` compiles, from /root/reference where they lie and unmodified,
  * OR/OptixModel.cpp + HalfSphere.cpp + Sphere.cpp  -> loadOBJ, placeReceiver        (oracle/_ref/ref_scene_dump)
  * OR/devicePrograms.cu                              -> raygen / closest-hit / miss   (oracle/_ref/libref_device.so)
with g++ -ffp-contract=off against host stand-ins for the OptiX device API (oracle/ref_stubs/).  Their outputs are
committed as tests/golden/ref_*.npz|json by tests/golden/make_golden.py; every test below checks the oracle (and,
where it is host code, the product's C ABI) against those vectors, and re-derives them live wherever oracle/_ref exists
(this container, and the GPU box through the shipped binaries) so the vectors cannot go stale.

What stays outside the pin, by construction: the ray-triangle search OptiX runs on RT cores and cuRAND's XORWOW stream
seeded with clock64() -- both closed, both replaced by one documented recipe on BOTH sides (oracle/ref_device_shim.cpp).
"""
import json
import os
import sys

import numpy as np
import pytest

import audiorenderingv2_b200 as arv
import oracle
from oracle import ref, scene as osc
from conftest import GOLDEN, REFERENCE

sys.path.insert(0, GOLDEN)
import make_golden as mg     # noqa: E402  (the generator holds the case definitions; it only runs under __main__)

LIVE = ref.available()
needs_ref_tree = pytest.mark.skipif(not (LIVE and ref.reference_mounted()), reason="needs /root/reference and oracle/_ref")

EPS = 2.0 ** -24


def test_loadobj_digests_equal_tinyobj_digests():
    """The reference's loadOBJ (mesh per (shape, material), (v,vn,vt) dedupe, index order) yields exactly the flat
    triangles the tinyobj-level dump gives -- which the oracle's and the product's loaders are pinned to
    (test_oracle_cpu.py::test_obj_loader_matches_reference_tinyobj, test_host_cpu.py)."""
    a = json.load(open(os.path.join(GOLDEN, "ref_loadobj.json")))
    b = json.load(open(os.path.join(GOLDEN, "meshes.json")))
    assert len(a) >= 10
    for rel, meshes in a.items():
        assert meshes == b[rel]["meshes"], rel


@needs_ref_tree
def test_loadobj_live():
    a = json.load(open(os.path.join(GOLDEN, "ref_loadobj.json")))
    for rel in ("test.obj", "assets/models/cajaConToro.obj", "assets/models/3D_U.obj", "assets/models/leftHalf.obj"):
        assert mg.digest(ref.scene_dump(os.path.join(REFERENCE, rel))) == a[rel]


def test_receiver_placement_is_bit_identical_to_placeReceiver(golden_receiver):
    """place_receiver_half (OR/OptixModel.cpp:162-196: glm::rotate(mat4(1), -radians(rot), +Y) * vec4, then + cam):
    the oracle's restatement and the product's host code (arv2_receiver_place, csrc/host/scene.cpp) reproduce the
    reference's placed vertices bit for bit."""
    g = np.load(os.path.join(GOLDEN, "ref_placement.npz"))
    recv = arv.Receiver.from_triangles(*golden_receiver)
    for k, case in enumerate(g["cases"]):
        cam, rot = case[:3], float(case[3])
        for side, tmpl in (("left", golden_receiver[0]), ("right", golden_receiver[1])):
            want = g[f"{side}_{k}"]
            got = osc.place_receiver_half(tmpl, cam, rot)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (side, case)
        pl, pr = recv.place(cam, rot)
        assert np.array_equal(pl.view(np.uint32), g[f"left_{k}"].view(np.uint32)), case
        assert np.array_equal(pr.view(np.uint32), g[f"right_{k}"].view(np.uint32)), case


@needs_ref_tree
def test_receiver_placement_live():
    g = np.load(os.path.join(GOLDEN, "ref_placement.npz"))
    cam, rot = mg.PLACEMENTS[3]
    m = ref.scene_dump(os.path.join(REFERENCE, "test.obj"), (os.path.join(REFERENCE, "assets/models/leftHalf.obj"),
                                                            os.path.join(REFERENCE, "assets/models/rightHalf.obj"), cam, rot))
    assert np.array_equal(m[-2][1], g["left_3"]) and np.array_equal(m[-1][1], g["right_3"])


def _check_shading(inp, want, got):
    """oracle_shade_hit against __closesthit__radiance.  Integer outcomes are equal; floats agree to fp32 rounding of
    the quantity they are differences of (the reference is plain a*b+c code that nvcc may or may not fuse; the oracle
    fixes where the fmas are): path length <= 8 ulp (99.9 % within 1); reflected direction (unit scale) <= 128 eps absolute, 99.9 % within
    16 eps (eps = 2^-24; the normal of a sliver triangle is a cancelling cross product); offset origin <= 32 eps of the
    coordinate scale; energies <= 1e-4 of (input energy x the largest chord, 2 m) with 99 % of the hits within 32 eps
    (the chord's discriminant b^2 - 4ac cancels near grazing incidence); bins equal except, for <= 1e-4 of the hits, the
    neighbouring bin when the path length sits on a bin edge."""
    n = len(inp["mat"])
    wall = inp["mat"] >= 0
    assert np.array_equal(want["depth"], got["depth"])
    assert np.array_equal(want["depth"][wall], np.full(wall.sum(), 4)) and np.all(want["depth"][~wall] == -1)
    assert np.array_equal(want["ndep"], got["ndep"])
    for k in range(2):
        live = want["ndep"] > k
        assert np.array_equal(want["ear"][live, k], got["ear"][live, k])
        d_bin = np.abs(want["idx"][live, k] - got["idx"][live, k])                # bin and bin + delay
        assert d_bin.max() <= 1 and (d_bin != 0).mean() <= 1e-4                   # a 1-ulp path length next to a bin edge
    assert (want["ndep"][~wall] == 2).mean() > 0.9
    e_in = inp["prd"][:, 0].astype(np.float64)
    ulp = np.abs(want["prd"][:, 1].view(np.int32).astype(np.int64) - got["prd"][:, 1].view(np.int32))
    assert ulp.max() <= 8 and (ulp <= 1).mean() >= 0.999 and (ulp == 0).mean() > 0.95
    d_dir = np.abs(want["prd"][wall, 5:8].astype(np.float64) - got["prd"][wall, 5:8]).max(axis=1)
    assert d_dir.max() <= 128 * EPS and np.percentile(d_dir, 99.9) <= 16 * EPS      # sliver triangles: cross product cancels
    scale = np.maximum(1.0, np.abs(want["prd"][wall, 2:5]).max(axis=1))
    d_pos = np.abs(want["prd"][wall, 2:5].astype(np.float64) - got["prd"][wall, 2:5]).max(axis=1) / scale
    assert d_pos.max() <= 32 * EPS and np.percentile(d_pos, 99.9) <= 8 * EPS
    d_e = np.abs(want["prd"][:, 0].astype(np.float64) - got["prd"][:, 0]) / (2.0 * e_in)
    assert d_e.max() <= 1e-4 and np.percentile(d_e, 99) <= 32 * EPS
    assert np.array_equal(want["prd"][wall, 0], got["prd"][wall, 0])              # energy * (1 - absorption): one multiply
    for k in range(2):
        live = want["ndep"] > k
        d_v = np.abs(want["val"][live, k].astype(np.float64) - got["val"][live, k]) / (2.0 * e_in[live])
        assert d_v.max() <= 1e-4 and np.percentile(d_v, 99) <= 32 * EPS
    return n


def test_shading_matches_reference_closesthit():
    g = np.load(os.path.join(GOLDEN, "ref_shading.npz"))
    inp = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    want = {k[4:]: g[k] for k in g.files if k.startswith("out_")}
    got = mg.run_shading(oracle.lib().oracle_shade_hit, inp)
    assert _check_shading(inp, want, got) == 4096


@pytest.mark.skipif(not LIVE, reason="needs oracle/_ref/libref_device.so")
def test_shading_live_100k_hits():
    """The same on 10^5 fresh hits, both sides run now; also the golden file is what the library produces today."""
    g = np.load(os.path.join(GOLDEN, "ref_shading.npz"))
    inp = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    again = mg.run_shading(ref.lib().ref_closesthit, inp)
    for k, v in again.items():
        assert np.array_equal(v, g["out_" + k]), k
    inp = mg.shading_inputs(100_000, seed=12345)
    for mono, hrtf, fs, ir_len in ((False, 0.9, 48000, 96000), (True, 0.25, 16000, 16000)):
        want = mg.run_shading(ref.lib().ref_closesthit, inp, fs, hrtf, mono, ir_len)
        got = mg.run_shading(oracle.lib().oracle_shade_hit, inp, fs, hrtf, mono, ir_len)
        if mono:
            assert want["ndep"].max() == 1
        _check_shading(inp, want, got) if not mono else _check_mono(inp, want, got)


def _check_mono(inp, want, got):
    assert np.array_equal(want["ndep"], got["ndep"]) and np.array_equal(want["depth"], got["depth"])
    live = want["ndep"] > 0
    assert np.array_equal(want["ear"][live, 0], got["ear"][live, 0])
    assert (want["idx"][live, 0] != got["idx"][live, 0]).mean() <= 1e-4
    assert (~live[inp["mat"] < 0]).mean() > 0.3                # @16 kHz, 1 s IR: hits beyond the IR deposit nothing


def _oracle_render(case, golden_scenes, golden_receiver):
    c = mg.RENDER_CASES[case]
    key = {"c1": "test", "caja": "caja"}[case]
    names = [str(n) for n in golden_scenes[key + "_names"]]
    model = osc.Model(meshes=[osc.Mesh(names[i], golden_scenes[key + "_verts"][golden_scenes[key + "_mesh"] == i]) for i in range(len(names))])
    flat = osc.flatten(model, osc.ReceiverTemplate(*golden_receiver), c["center"], c["yaw"], [])
    flat.absorption[:] = c["absorption"]
    p = oracle.make_params(rays=c["rays"], emitter=c["emitter"], sphere_center=c["center"], base_power=c["base_power"],
                           max_bounces=c["max_bounces"], hrtf=c["hrtf"], sample_rate=c["sample_rate"], ir_length=c["ir_length"], seed=c["seed"])
    return flat, oracle.trace(p, flat, use_bvh=False)


def _check_render(want, o, ir_length, energy0):
    """Per-ray outcome of the reference's raygen + programs against the oracle's.  The two sides start each ray in
    directions that differ by float rounding (the reference takes theta = 2 pi u1 and phi = acos(2 u2 - 1) through a
    float32 u1 and libm; the oracle takes them exactly), so a ray may graze an edge differently or land next to a bin
    boundary: >= 99.9 % of the rays take the same number of segments AND end in the same ear AND the same bin; the
    deposited energies of those rays agree to 1e-3 of the start energy; the IR's total energy to 1e-5."""
    n = len(want["bin"])
    o_dep = (o["ear"] > 0) & (o["bin"] < ir_length)
    o_bin = np.where(o_dep, o["bin"], -1); o_ear = np.where(o_dep, o["ear"], 0)
    same = (want["bin"] == o_bin) & (want["ear"] == o_ear) & (want["nseg"] == o["nseg"])
    assert same.mean() >= 0.999, same.mean()
    both = same & o_dep
    assert both.sum() > 0.01 * n
    # the chord weight sqrt(b^2 - 4ac) magnifies the 1e-7 direction difference near grazing incidence, so the
    # deposits are compared in units of the ray's start energy: all within 1e-3, 99 % within 1e-4
    d_e = np.abs(want["energy"][both].astype(np.float64) - o["energy"][both, 0]) / energy0
    assert d_e.max() <= 1e-3 and np.percentile(d_e, 99) <= 1e-4
    tot_w, tot_o = want["hist_val"].sum(), o["hist"].sum()
    assert abs(tot_w - tot_o) <= 1e-5 * tot_o
    assert abs(int(want["segments"]) - o["segments"]) <= 1e-3 * o["segments"]
    return same.mean()


@pytest.mark.parametrize("case", ["c1", "caja"])
def test_oracle_trace_matches_reference_programs(case, golden_scenes, golden_receiver):
    """BASELINE config 1 (test.obj, 100k rays) and the closed box (50 bounces): the oracle's whole per-ray loop against
    the reference's own __raygen__renderFrame / __closesthit__radiance / __miss__radiance run on the CPU."""
    g = np.load(os.path.join(GOLDEN, "ref_render.npz"))
    want = {k[len(case) + 1:]: g[k] for k in g.files if k.startswith(case + "_")}
    flat, o = _oracle_render(case, golden_scenes, golden_receiver)
    c = mg.RENDER_CASES[case]
    frac = _check_render(want, o, c["ir_length"], c["base_power"] / (c["rays"][0] * c["rays"][1] * c["rays"][2] * 4.18879020478))
    if case == "c1":
        assert frac == 1.0           # the open scene: every one of the 100 000 rays agrees


@needs_ref_tree
def test_reference_programs_live(golden_scenes, golden_receiver):
    """Re-run the reference's programs now (scene through the reference's own loadOBJ + placeReceiver): the committed
    vectors are what they produce, and the flat scene the oracle traces is the reference's, triangle for triangle."""
    g = np.load(os.path.join(GOLDEN, "ref_render.npz"))
    for case in ("c1", "caja"):
        (tv, tm, ab), r = mg.run_render(mg.RENDER_CASES[case])
        flat, _ = _oracle_render(case, golden_scenes, golden_receiver)
        assert np.array_equal(np.asarray(flat.tri_verts, np.float32).reshape(-1, 3, 3), tv)
        assert np.array_equal(np.asarray(flat.tri_mat) < 0, tm < 0) and np.array_equal(np.asarray(flat.tri_mat)[tm < 0], tm[tm < 0])
        assert np.array_equal(r["bin"], g[case + "_bin"]) and np.array_equal(r["nseg"], g[case + "_nseg"])
        assert np.array_equal(r["energy"][:, 0], g[case + "_energy"])


def test_reference_direction_recipe_is_the_oracles_distribution():
    """OR/devicePrograms.cu:219-224 -- theta = 2 pi u1, phi = acos(2 u2 - 1), (sin phi cos theta, sin phi sin theta,
    cos phi) -- against the oracle's integer/polynomial evaluation of the same map for the same Philox words."""
    n = 50_000
    un = ref.uniforms_for_rays(9, 0, n).astype(np.float64)
    theta = 2.0 * np.float64(np.float32(3.141592654)) * un[:, 0]
    phi = np.arccos(2.0 * un[:, 1] - 1.0)
    want = np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)], axis=1)
    got = np.stack([oracle.ray_direction(9, i) for i in range(0, n, 25)])
    assert np.abs(got - want[::25]).max() <= 2e-6
