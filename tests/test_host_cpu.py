"""CPU suite for the product's host side: the C-ABI library loads and exports every symbol
include/arv2.h declares, the C++ front end (OBJ/MTL loader, receiver placement, config,
WAV) agrees with the golden fixtures and with the oracle's independent numpy restatement,
and everything that needs a GPU fails loudly instead of falling back."""
import hashlib
import json
import os
import re

import numpy as np
import pytest

import audiorenderingv2_b200 as arv
from audiorenderingv2_b200 import scenes, sharding
from oracle import scene as osc
from conftest import GOLDEN, REFERENCE, ROOT

HAVE_REF = os.path.isdir(os.path.join(REFERENCE, "assets", "models"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).view(np.uint32).tobytes()).hexdigest()


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "arv2.h")).read()
    declared = set(re.findall(r"\b(arv2_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 45
    L = arv.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/arv2.h but not exported"
    assert declared == set(arv.SYMBOLS), declared ^ set(arv.SYMBOLS)
    assert b"sm_100a" in L.arv2_version()


def test_no_cpu_fallback():
    """Without a CUDA device every GPU entry point reports ARV2_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    scene = arv.Scene.from_triangles(np.zeros((0, 3, 3), np.float32), np.zeros(0, np.int32), [])
    with pytest.raises(arv.Arv2Error, match=r"error -3"):
        arv.AudioRenderer(scene, 1, 16000, [], (10, 10, 1))
    with pytest.raises(arv.Arv2Error, match=r"error -3"):
        arv.ConvStream(1, 512, 4800)


def test_argument_validation():
    with pytest.raises(arv.Arv2Error, match=r"error -2"):
        arv.Scene.load_obj("/nonexistent/scene.obj")
    with pytest.raises(arv.Arv2Error):
        arv.Scene.from_triangles(np.zeros((1, 3, 3), np.float32), np.array([3], np.int32), ["a"])
    with pytest.raises(arv.Arv2Error):
        arv.parse_config("{ not json")
    with pytest.raises(arv.Arv2Error):
        arv.ConvStream(1, 500, 4800)          # block must be a power of two (checked before the device)


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_cpp_obj_loader_matches_reference_tinyobj():
    gold = json.load(open(os.path.join(GOLDEN, "meshes.json")))
    for rel, g in gold.items():
        path = os.path.join(REFERENCE, rel)
        if "error" in g:
            with pytest.raises(arv.Arv2Error, match="could not parse materials"):
                arv.Scene.load_obj(path)
            continue
        s = arv.Scene.load_obj(path)
        tv, tm = s.triangles()
        names = s.mesh_materials()
        got = [(names[i], int((tm == i).sum()), sha(tv[tm == i])) for i in range(len(names))]
        assert got == [(e["material"], e["tris"], e["sha256"]) for e in g["meshes"]], rel


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_receiver_load_matches_golden(golden_receiver):
    r = arv.Receiver.load(os.path.join(REFERENCE, "assets/models/leftHalf.obj"), os.path.join(REFERENCE, "assets/models/rightHalf.obj"))
    l, rr = r.place((0, 0, 0), 0.0)
    assert r.counts() == (510, 510)
    assert np.array_equal(l, golden_receiver[0] + np.float32(0)) and np.array_equal(rr, golden_receiver[1] + np.float32(0))


def test_receiver_placement_equals_oracle(golden_receiver):
    r = arv.Receiver.from_triangles(*golden_receiver)
    rng = np.random.default_rng(0)
    for _ in range(20):
        cam = rng.uniform(-30, 30, 3).astype(np.float32)
        rot = float(rng.uniform(0, 360))
        l, rr = r.place(cam, rot)
        assert np.array_equal(l.view(np.uint32), osc.place_receiver_half(golden_receiver[0], cam, rot).view(np.uint32))
        assert np.array_equal(rr.view(np.uint32), osc.place_receiver_half(golden_receiver[1], cam, rot).view(np.uint32))


def test_scene_roundtrip_and_bounds(golden_scenes):
    tv, tm = golden_scenes["test_verts"], golden_scenes["test_mesh"]
    s = arv.Scene.from_triangles(tv, tm, [str(n) for n in golden_scenes["test_names"]])
    assert s.counts() == (116, 4)
    tv2, tm2 = s.triangles()
    assert np.array_equal(tv, tv2) and np.array_equal(tm, tm2)
    lo, hi = s.bounds()
    assert np.allclose(lo, tv.min(axis=(0, 1))) and np.allclose(hi, tv.max(axis=(0, 1)))
    assert np.allclose(lo, [-13.453613, 0.0, -14.577106]) and np.allclose(hi[0], 18.02746)


def _bvh_stats_in_subprocess(tris, env):
    """The builder reads its tuning knobs once per process, so every setting gets a fresh interpreter."""
    import subprocess, sys, tempfile
    with tempfile.NamedTemporaryFile(suffix=".npy") as f:
        np.save(f.name, np.ascontiguousarray(tris, np.float32))
        code = ("import json, sys, numpy as np; sys.path.insert(0, %r); import audiorenderingv2_b200 as arv; t = np.load(%r); "
                "print(json.dumps(arv.Scene.from_triangles(t, np.zeros(len(t), np.int32), ['m']).bvh_stats()))") % (ROOT, f.name)
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_host_bvh_is_valid_for_ragged_inputs(golden_scenes):
    """Host SAH builder (replaces optixAccelBuild, OR/AudioRenderer.cpp:95-218): every triangle in exactly one leaf,
    child boxes contain their triangles, children after parents -- for empty, tiny, degenerate (coincident
    centroids, zero-area) and real inputs, with the SAH leaf test (default) and without it."""
    rng = np.random.default_rng(3)
    one = rng.standard_normal((1, 3, 3)).astype(np.float32)
    cases = {
        "empty": np.zeros((0, 3, 3), np.float32),
        "one": one,
        "five": rng.standard_normal((5, 3, 3)).astype(np.float32),
        "coincident": np.repeat(one, 37, axis=0),                                   # 37 copies of one triangle
        "zero_area": np.zeros((9, 3, 3), np.float32),
        "test_obj": golden_scenes["test_verts"],
        "random_2000": (rng.standard_normal((2000, 1, 3)) * 10 + rng.standard_normal((2000, 3, 3)) * 0.3).astype(np.float32),
    }
    for name, tris in cases.items():
        for env in ({}, {"ARV2_SAH_CT": "0"}, {"ARV2_LEAF_MAX": "8", "ARV2_SAH_CT": "1.5"}, {"ARV2_LEAF_MAX": "1"}, {"ARV2_SAH_SWEEP": "64"}):
            st = _bvh_stats_in_subprocess(tris, env)
            assert st["valid"] == 1, (name, env, st)
            assert st["n_tris"] == len(tris) and st["max_leaf_tris"] <= int(env.get("ARV2_LEAF_MAX", 4)) or len(tris) <= 4, (name, env, st)
            assert st["depth"] + 3 <= 64, (name, env, st)              # kTraversalStack of the kernels
    # the SAH leaf test ends the tessellated scenes in quads: fewer triangle tests for a few more nodes
    tv, _, _ = scenes.conference_room(target_tris=20_000)
    with_test, without = _bvh_stats_in_subprocess(tv, {}), _bvh_stats_in_subprocess(tv, {"ARV2_SAH_CT": "0"})
    assert with_test["valid"] == 1 and without["valid"] == 1
    assert with_test["sah_tris"] < 0.75 * without["sah_tris"] and with_test["sah_nodes"] < 1.15 * without["sah_nodes"]


CONFIG_CASES = [
    "{}",
    json.dumps({"pathtracer_parameters": {"hrtf_absorption_rate": 0.9, "ray_max_bounces": 99.5, "base_power": 3.62,
                                          "rays": {"x": 10, "y": 20, "z": 30}, "ray_distance_threshold": 2000.0,
                                          "materials": [{"name": "low", "mat_absorption": 0.1}, {"name": 3}, {"name": "x", "mat_absorption": 0.25}]},
                "renderer_parameters": {"ir_length_in_seconds": 1.5, "re_render_distance_threshold": 2.4, "width": 800.6,
                                        "write_first_ir_to_file": True, "initial_volume": 0.5},
                "scene_parameters": {"mono": True, "scene_file_path": "a/b.obj", "audio_file_path": "c.wav",
                                     "initial_receiver_pos": {"x": 1, "y": 2, "z": 3}, "initial_emitter_pos": {"x": 1, "y": 2}}}),
]


@pytest.mark.parametrize("text", CONFIG_CASES)
def test_config_matches_oracle(text):
    c = arv.parse_config(text)
    o = osc.load_config(text)
    assert c.ir_length_in_seconds == o["ir_length_in_seconds"] and c.width == o["width"] and c.height == o["height"]
    assert c.ray_max_bounces == o["ray_max_bounces"]
    assert c.hrtf_absorption_rate == pytest.approx(o["hrtf_absorption_rate"], abs=1e-7)
    assert c.base_power == pytest.approx(o["base_power"], rel=1e-7)
    assert tuple(c.rays) == o["rays"] and bool(c.mono) == o["mono"]
    assert tuple(c.initial_receiver_pos) == o["initial_receiver_pos"] and tuple(c.initial_emitter_pos) == o["initial_emitter_pos"]
    assert c.scene_file_path.decode() == o["scene_file_path"] and c.audio_file_path.decode() == o["audio_file_path"]
    assert c.re_render_distance_threshold == o["re_render_distance_threshold"]
    assert bool(c.write_first_ir_to_file) == o["write_first_ir_to_file"]
    assert [(n, pytest.approx(a)) for n, a in arv.config_materials(c)] == o["materials"]


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_shipped_config_json():
    c = arv.load_config(os.path.join(REFERENCE, "config.json"))
    assert c.hrtf_absorption_rate == 1.0 and c.ray_max_bounces == 100 and c.ir_length_in_seconds == 2
    assert c.scene_file_path.decode().endswith("3D_U.obj") and c.n_materials == 5
    assert tuple(c.rays) == (100.0, 100.0, 100.0) and abs(c.base_power - 3.62) < 1e-6


def test_material_absorption_rule():
    mats = [("red", 0.2), ("blue", 0.9)]
    assert arv.material_absorption("receiver_left", mats) == -1.0 and arv.material_absorption("receiver_right", mats) == -2.0
    assert arv.material_absorption("blue", mats) == pytest.approx(0.9) and arv.material_absorption("Amarillo", mats) == 0.5


def test_wav_read_write(tmp_path):
    d = np.load(os.path.join(GOLDEN, "wav_decode.npz"))
    pcm = d["pcm"]
    path = tmp_path / "a.wav"
    with open(path, "wb") as fh:            # mono int16 @16 kHz
        fh.write(b"RIFF" + (36 + 2 * len(pcm)).to_bytes(4, "little") + b"WAVEfmt " + (16).to_bytes(4, "little")
                 + (1).to_bytes(2, "little") + (1).to_bytes(2, "little") + (16000).to_bytes(4, "little")
                 + (32000).to_bytes(4, "little") + (2).to_bytes(2, "little") + (16).to_bytes(2, "little")
                 + b"data" + (2 * len(pcm)).to_bytes(4, "little") + pcm.astype("<i2").tobytes())
    sr, ch, a = arv.wav_read(path)
    assert (sr, ch) == (16000, 1)
    assert np.abs(a - d["text"]).max() <= 6e-7                                  # the reference's own dump
    osr, oa = osc.read_wav(str(path))
    assert np.array_equal(a, oa[0])
    out = tmp_path / "Result.wav"
    arv.wav_write_stereo_normalized(out, a, -a, sr)
    sr2, ch2, l = arv.wav_read(out)
    assert (sr2, ch2) == (16000, 2) and l.max() <= 1.0 and l.min() >= -1.0 and abs(l.max() - 1.0) < 1e-3
    with pytest.raises(arv.Arv2Error):
        arv.wav_write_stereo_normalized(out, np.zeros(10, np.float32), np.zeros(10, np.float32), sr)   # OR/main.cpp:641-643


def test_procedural_scenes_are_deterministic():
    tv, tm, names = scenes.conference_room()
    assert 330_000 <= len(tv) <= 332_000 and len(names) == 6 and tm.max() == 5
    tv2, _, _ = scenes.conference_room()
    assert np.array_equal(tv, tv2)
    assert sha(tv) == sha(tv2)
    lo, hi = tv.min(axis=(0, 1)), tv.max(axis=(0, 1))
    assert np.allclose(lo, [-0.2, -0.2, -0.2], atol=1e-6) and np.allclose(hi, [12.2, 3.2, 8.2], atol=1e-6)


def test_ray_range_partition():
    for n in (0, 1, 7, 1_000_000, 100_000_007):
        for w in (1, 2, 3, 4, 8):
            parts = [sharding.ray_range(r, w, n) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert sharding.sources_of(1, 8, 16) == [1, 9] and sum(len(sharding.sources_of(r, 3, 16)) for r in range(3)) == 16


# ----------------------------------------------------- callers either side of the path (8f rows 3-4)
def test_global_angle_convention():
    """Camera::calculate_global_angle: degrees(atan2(z, x)) in [0, 360) (OR/Camera.cpp:31-41)."""
    assert arv.global_angle(1, 0) == 0.0
    assert arv.global_angle(0, 1) == pytest.approx(90.0)
    assert arv.global_angle(-1, 0) == pytest.approx(180.0)
    assert arv.global_angle(0, -1) == pytest.approx(270.0)
    assert 0.0 <= arv.global_angle(0.3, -0.0001) < 360.0


def test_rerender_policy_matches_main_loop():
    """OR/main.cpp:470-498: distance > threshold, shortest-arc angle > threshold, or > 1 s
    (whole seconds) after the first movement; never while a render is in flight."""
    p = arv.RerenderPolicy(2.0, 5.0, (0, 0, 0), 350.0)
    assert not p.update((0, 0, 0), 350.0, 100.0)                 # nothing moved
    assert not p.update((1.9, 0, 0), 350.0, 100.2)               # below the distance threshold, timer starts
    assert not p.update((1.9, 0, 0), 350.0, 101.9)               # difftime(101, 100) = 1, not > 1
    assert p.update((1.9, 0, 0), 350.0, 102.0)                   # 2 whole seconds later
    assert not p.update((1.9, 0, 0), 350.0, 102.1)               # origin moved to the last render
    assert p.update((1.9, 0, 2.1), 350.0, 102.2)                 # distance
    assert not p.update((1.9, 0, 2.1), 354.0, 102.3)             # 4 degrees
    assert p.update((1.9, 0, 2.1), 2.0, 102.4)                   # 350 -> 2 = 12 degrees over the wrap
    assert not p.update((1.9, 0, 2.1), 10.0, 102.5, is_rendering=True)   # 8 degrees, but a render is in flight
    assert p.update((1.9, 0, 2.1), 10.0, 102.6)


def test_playback_callback_contract():
    """audioHandler (OR/main.cpp:69-97): interleaved LRLR doubles, x100 x volume, position from
    streamTime modulo the file length, the reference's i + position indexing."""
    n = 1000
    l = np.arange(n, dtype=np.float32); r = -np.arange(n, dtype=np.float32)
    out, w = arv.playback_fill(4, 0.0105, 16000, l, r, output_buffer_len=4 * n, volume=0.5)
    pos = int(0.0105 * 16000) % n
    assert w == 8
    exp = [(l if i % 2 == 0 else r)[i + pos] * 100 * 0.5 for i in range(8)]
    assert np.allclose(out, exp)
    out, w = arv.playback_fill(256, 0.0, 16000, l, r, output_buffer_len=100, volume=1.0)   # stops at output_buffer_len
    assert w == 100
    out, w = arv.playback_fill(4, (n + 3.5) / 16000.0, 16000, l, r, output_buffer_len=4 * n)  # wraps modulo the length
    assert out[0] == l[3] * 100


def test_ring_is_the_reference_circular_buffer():
    """OR/CircularBuffer.h: add() overlap-adds without advancing, get_and_reset() pops + zeroes."""
    ring = arv.Ring(8)
    ring.add([1, 2, 3, 4, 5])
    ring.add([10, 20])
    assert list(ring.get_and_reset(3)) == [11, 22, 3]
    ring.add([1, 1, 1, 1, 1, 1, 1])                 # wraps around the end
    assert list(ring.get_and_reset(8)) == [5, 6, 1, 1, 1, 1, 1, 0]
    with pytest.raises(arv.Arv2Error):
        ring.get_and_reset(9)


def test_shard_range_in_the_library_is_the_python_rule():
    for n in (0, 1, 7, 1_000_000, 100_000_007):
        for w in (1, 2, 3, 4, 8):
            for r in range(w):
                assert arv.shard_range(n, r, w) == sharding.ray_range(r, w, n)


def test_convolved_output_text_dump(tmp_path):
    """output_convolute_left.txt / _right.txt (OR/AudioRenderer.cpp:720-744): one value per line in ostream's default
    format (6 significant digits), readable the way utils/main.py:17-28 reads them."""
    rng = np.random.default_rng(2)
    l = (rng.standard_normal(1000) * 10.0 ** rng.integers(-8, 3, 1000)).astype(np.float32); r = -l[::-1].copy()
    a, b = tmp_path / "output_convolute_left.txt", tmp_path / "output_convolute_right.txt"
    arv.write_convolved_text(a, b, l, r)
    la = open(a).read().split("\n")
    assert len(la) == 1001 and la[-1] == ""
    assert la[:5] == ["%g" % v for v in l[:5]]
    assert np.allclose([float(s) for s in la[:-1]], l, rtol=1e-5, atol=0)
    assert np.allclose([float(s) for s in open(b)], r, rtol=1e-5, atol=0)


def test_nccl_binds_at_run_time_and_needs_a_device():
    """libarv2 dlopens libnccl.so.2 (no link-time dependency); a communicator needs a CUDA device like everything else."""
    import subprocess
    deps = subprocess.run(["ldd", arv.LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl" not in deps
    uid = arv.Comm.unique_id()
    assert len(uid) == arv.COMM_ID_BYTES and uid != bytes(arv.COMM_ID_BYTES)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(arv.Arv2Error, match=r"error -3"):
            arv.Comm(0, 0, 1, uid)


def test_direction_tile_rule_partitions_the_sphere_evenly():
    """The default multi-GPU shard rule (host-side statement in sharding.direction_tile_rank of direction_select_kernel):
    every direction belongs to exactly one rank, the ranks' shares of a uniform ray set are equal to a few per cent, and
    a rank's rays are spread over the whole sphere (every octant), not one patch of it."""
    from audiorenderingv2_b200 import sharding
    rng = np.random.default_rng(3)
    d = rng.standard_normal((200_000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    assert sharding.direction_tile_bits(8_000_000) == 14 and sharding.direction_tile_bits(100_000) == 8 and sharding.direction_tile_bits(100) == 6
    for world in (2, 3, 8):
        r = sharding.direction_tile_rank(d, world, 14)
        assert r.min() == 0 and r.max() == world - 1
        share = np.bincount(r, minlength=world) / len(d)
        assert np.all(np.abs(share * world - 1.0) < 0.05)
        octant = (d[:, 0] > 0).astype(int) | ((d[:, 1] > 0).astype(int) << 1) | ((d[:, 2] > 0).astype(int) << 2)
        for k in range(world):
            assert len(np.unique(octant[r == k])) == 8
    # neighbouring directions share a tile: the point of the rule is that a rank's rays stay dense in direction
    base = np.array([[0.3, 0.5, 0.81]], np.float32)
    near = base + 1e-4 * rng.standard_normal((1000, 3)).astype(np.float32)
    assert np.mean(sharding.direction_tile_rank(near, 8, 14) == sharding.direction_tile_rank(base, 8, 14)[0]) > 0.95
