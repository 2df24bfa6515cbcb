"""CPU suite: the oracle against the golden fixtures and analytic known answers
(SURVEY.md 8c).  No GPU, no compute call into libarv2."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

import oracle
from oracle import scene as osc
from conftest import GOLDEN, REFERENCE

HAVE_REF = os.path.isdir(os.path.join(REFERENCE, "assets", "models"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).view(np.uint32).tobytes()).hexdigest()


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_obj_loader_matches_reference_tinyobj():
    gold = json.load(open(os.path.join(GOLDEN, "meshes.json")))
    assert len(gold) >= 15
    for rel, g in gold.items():
        path = os.path.join(REFERENCE, rel)
        if "error" in g:
            with pytest.raises(RuntimeError):
                osc.load_obj(path)
            continue
        m = osc.load_obj(path)
        assert m.materials == g["materials"], rel
        assert [(x.material_name, len(x.tris), sha(x.tris)) for x in m.meshes] == \
               [(e["material"], e["tris"], e["sha256"]) for e in g["meshes"]], rel


def test_golden_scene_fixtures_consistent(golden_scenes, golden_receiver):
    gold = json.load(open(os.path.join(GOLDEN, "meshes.json")))
    l, r = golden_receiver
    assert l.shape == (510, 3, 3) and r.shape == (510, 3, 3)
    assert sha(l) == gold["assets/models/leftHalf.obj"]["meshes"][0]["sha256"]
    assert sha(r) == gold["assets/models/rightHalf.obj"]["meshes"][0]["sha256"]
    assert golden_scenes["test_verts"].shape == (116, 3, 3)
    assert list(golden_scenes["test_names"]) == ["Amarillo", "Amarillo", "Rojo", "Luz"]


def test_wav_decode_rule():
    d = np.load(os.path.join(GOLDEN, "wav_decode.npz"))
    dec = d["pcm"].astype(np.float32) / np.float32(32768.0)          # AudioFile.h:1242-1245
    assert np.abs(dec - d["text"]).max() <= 6e-7                     # 6 significant digits in input.txt


def test_receiver_placement_rotation(golden_receiver):
    l, r = golden_receiver
    cam = (1.0, 2.0, 3.0)
    p0 = osc.place_receiver_half(l, cam, 0.0)
    assert np.allclose(p0, l + np.array(cam, np.float32), atol=1e-6)
    p90 = osc.place_receiver_half(l, (0, 0, 0), 90.0)                # R_y(-90): x' = -z, z' = x
    assert np.allclose(p90[..., 0], -l[..., 2], atol=1e-6) and np.allclose(p90[..., 2], l[..., 0], atol=1e-6)
    assert np.allclose(p90[..., 1], l[..., 1], atol=1e-6)
    # rigid: radii about the centre are preserved
    assert np.allclose(np.linalg.norm(osc.place_receiver_half(r, (0, 0, 0), 37.0), axis=-1), np.linalg.norm(r, axis=-1), atol=1e-5)


def test_config_defaults_and_quirks():
    c = osc.load_config("{}")
    assert c["ir_length_in_seconds"] == 2 and c["ray_max_bounces"] == 10 and c["base_power"] == 100.0
    assert c["hrtf_absorption_rate"] == 0.9 and c["rays"] == (100.0, 100.0, 100.0)
    assert c["initial_receiver_pos"] == (-2.5, 10.0, 0.0) and c["scene_file_path"].endswith("1D_U.obj")
    c = osc.load_config(json.dumps({"pathtracer_parameters": {"hrtf_absorption_rate": 0.9, "ray_max_bounces": 99.5,
                                                              "ray_distance_threshold": 2000.0},
                                    "renderer_parameters": {"ir_length_in_seconds": 1.5, "re_render_distance_threshold": 2.4}}))
    assert c["hrtf_absorption_rate"] == 1.0          # Context.cpp:143-145 round()
    assert c["ray_max_bounces"] == 100 and c["ir_length_in_seconds"] == 2 and c["re_render_distance_threshold"] == 2.0


def test_material_lookup():
    mats = [("red", 0.2), ("blue", 0.9)]
    assert osc.material_absorption("receiver_left", mats) == -1.0
    assert osc.material_absorption("receiver_right", mats) == -2.0
    assert osc.material_absorption("blue", mats) == 0.9
    assert osc.material_absorption("Amarillo", mats) == 0.5


# ------------------------------------------------------------------------- tracer
def free_field(golden_receiver, D, n=(200, 100, 10), fs=16000, ir_s=1, hrtf=1.0, base_power=100.0, yaw=0.0):
    flat = osc.flatten(osc.Model(), osc.ReceiverTemplate(*golden_receiver), (D, 0, 0), yaw, [])
    p = oracle.make_params(rays=n, emitter=(0, 0, 0), sphere_center=(D, 0, 0), base_power=base_power, max_bounces=10,
                           hrtf=hrtf, sample_rate=fs, ir_length=ir_s * fs, seed=3)
    return oracle.trace(p, flat), p


def test_rng_is_philox4x32_10_known_answer():
    # Random123 known-answer vector: counter = key = 0
    assert [hex(int(v)) for v in oracle.philox(0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


def test_direction_distribution():
    d = np.array([oracle.ray_direction(11, i) for i in range(40000)], np.float64)
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-6
    assert np.abs(d.mean(0)).max() < 0.01                      # isotropic
    assert abs((d[:, 2] ** 2).mean() - 1 / 3) < 0.01           # cos(phi) uniform on [-1,1]
    assert not np.array_equal(oracle.ray_direction(11, 5), oracle.ray_direction(12, 5))


def test_free_field_energy_and_support(golden_receiver):
    """no walls => sum(L+R) -> P/(4 pi D^2); support of the direct sound =
    [round((D-1) fs/343), round(sqrt(D^2-1) fs/343)] (+-1 bin for the facetted ball)."""
    D, fs = 6.0, 16000
    o, p = free_field(golden_receiver, D, n=(500, 400, 5))
    total = o["hist"].sum()
    expect = 100.0 / (4 * math.pi * D * D)
    # the receiver is two facetted half-balls with a 0.085 m gap, weighted by the chord of
    # the ideal unit ball: -4 % (gap towards the source) .. +2 % (side-on) of the ideal
    assert abs(total - expect) < 0.08 * expect
    ir = o["hist"].sum(axis=(0, 1))
    nz = np.nonzero(ir)[0]
    assert abs(nz[0] - round((D - 1) * fs / 343)) <= 1
    # a few rays slip into the gap between the halves and land on the far cap (up to D+1);
    # everything else ends at the tangent distance sqrt(D^2-1)
    last = round(math.sqrt(D * D - 1) * fs / 343) + 1
    assert ir[: last + 1].sum() > 0.98 * ir.sum() and nz[-1] <= round((D + 1) * fs / 343) + 1
    assert (o["nseg"] == 1).all() and o["segments"] == 500 * 400 * 5


def test_reference_output_ir_fixture_shape(golden_receiver):
    """prebuild/obj_raytracer/output_ir.txt: the direct-sound run 389..434 is what a
    receiver at D = 9.34 m gives at 16 kHz (D-1 = 8.34 m -> bin 389, sqrt(D^2-1) -> 433);
    its total is the one-ear share of P/(4 pi D^2) with the shipped base_power... the
    dump's config is unknown, so only the support and the order of magnitude are pinned."""
    g = json.load(open(os.path.join(GOLDEN, "output_ir.json")))
    idx = np.array([i for i, _ in g["nonzero"]]); val = np.array([v for _, v in g["nonzero"]])
    direct = idx < 500
    assert g["length"] == 64000 and idx[direct].min() == 389 and idx[direct].max() in (433, 434)
    D = 1 + 389.0 * 343 / 16000
    o, _ = free_field(golden_receiver, D, n=(200, 200, 10), ir_s=4, hrtf=1.0, base_power=100.0, yaw=0.0)
    ir = o["hist"].sum(axis=(0, 1))
    nz = np.nonzero(ir)[0]
    assert abs(nz[0] - 389) <= 1 and ir[:435].sum() > 0.98 * ir.sum()
    ear = o["hist"][0, 0].sum()                                # one ear, like the dumped left IR
    assert 0.4 < val[direct].sum() / ear < 2.5


def test_hrtf_delay_and_mono(golden_receiver):
    o, p = free_field(golden_receiver, 5.0, n=(100, 100, 4), hrtf=0.75)
    left_hits = o["ear"] == 1
    e = o["energy"][:, 0].astype(np.float64)
    # every left-ear hit deposits e in L[bin] and e*(1-hrtf) in R[bin+7]   (devicePrograms.cu:125-148)
    L = np.zeros(16000); R = np.zeros(16000)
    for b, ear, en in zip(o["bin"], o["ear"], e):
        if ear == 1:
            L[b] += en; R[b + 7] += np.float32(en) * np.float32(0.25)
        elif ear == 2:
            R[b] += en; L[b + 7] += np.float32(en) * np.float32(0.25)
    assert np.allclose(L, o["hist"][0, 0], rtol=1e-12) and np.allclose(R, o["hist"][1, 0], rtol=1e-12)
    assert left_hits.any()
    l, r = oracle.finalize_ir(o["hist"], mono=True)
    assert np.array_equal(l, r)


def test_closed_box_reflections(golden_scenes, golden_receiver):
    """caja.obj (closed cube, inner half-width 25.84 m): energy of a path that reached the
    receiver after k wall hits is e0 * (1-a)^k * chord, chord <= 2; first-order image
    sources put energy near the predicted arrival bins."""
    a = 0.3
    tv, tm, names = golden_scenes["caja_verts"], golden_scenes["caja_mesh"], golden_scenes["caja_names"]
    model = osc.Model(meshes=[osc.Mesh(str(names[0]), tv)])
    em, rc = np.array([3.0, 1.0, -2.0]), np.array([-8.0, 4.0, 6.0])
    flat = osc.flatten(model, osc.ReceiverTemplate(*golden_receiver), rc, 0.0, [(str(names[0]), a)])
    n = (200, 100, 20)
    fs = 4000
    p = oracle.make_params(rays=n, emitter=em, sphere_center=rc, base_power=100.0, max_bounces=4, hrtf=1.0,
                           sample_rate=fs, ir_length=fs, seed=2)
    o = oracle.trace(p, flat)
    e0 = np.float32(100.0 / (400000 * 4.18879020478))
    hit = o["ear"] > 0
    k = o["nseg"][hit] - 1
    assert (o["energy"][hit, 0] <= e0 * (1 - a) ** k * 2.0 * 1.0001).all()
    assert (k >= 0).all() and k.max() >= 1
    inner = float(np.abs(tv).min(axis=(0, 1)).max())          # inner wall coordinate
    ir = o["hist"].sum(axis=(0, 1))
    for ax in range(3):
        for sgn in (-1, 1):
            img = em.copy(); img[ax] = 2 * sgn * inner - em[ax]
            d = np.linalg.norm(img - rc)
            b0, b1 = int((d - 1.2) * fs / 343), int(math.sqrt(d * d - 1) * fs / 343) + 2
            if b1 < fs:
                assert ir[b0:b1].sum() > 0, (ax, sgn)


def test_bvh_equals_bruteforce(golden_scenes, golden_receiver):
    tv, tm = golden_scenes["toro_verts"], golden_scenes["toro_mesh"]
    model = osc.Model(meshes=[osc.Mesh("a", tv[tm == 0]), osc.Mesh("b", tv[tm == 1])])
    flat = osc.flatten(model, osc.ReceiverTemplate(*golden_receiver), (6, 2, 3), 10.0, [])
    p = oracle.make_params(rays=(40, 40, 2), emitter=(0, 5, 0), sphere_center=(6, 2, 3), max_bounces=20, seed=9)
    a = oracle.trace(p, flat, use_bvh=False)
    b = oracle.trace(p, flat, use_bvh=True, n_threads=3)
    assert np.array_equal(a["bin"], b["bin"]) and np.array_equal(a["energy"], b["energy"]) and a["segments"] == b["segments"]
    assert np.allclose(a["hist"], b["hist"], rtol=1e-12, atol=0)
    # sharded ranges reproduce the full set
    c = oracle.trace(p, flat, ray_begin=1000, n_rays=500)
    assert np.array_equal(c["bin"], a["bin"][1000:1500])


def test_intersector_known_answers():
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    i, t, u, v = oracle.closest_hit(tri, (0.25, 0.25, 1.0), (0, 0, -1))
    assert i == 0 and t == 1.0 and u == 0.25 and v == 0.25
    assert oracle.closest_hit(tri, (0.25, 0.25, 1.0), (0, 0, 1))[0] == -1          # behind the origin
    assert oracle.closest_hit(tri, (0.25, 0.25, -1.0), (0, 0, 1))[0] == 0          # two-sided
    assert oracle.closest_hit(tri, (2.0, 2.0, 1.0), (0, 0, -1))[0] == -1
    two = np.concatenate([tri + np.float32([0, 0, 0.5]), tri])                     # nearer one wins
    assert oracle.closest_hit(two, (0.2, 0.2, 1.0), (0, 0, -1))[0] == 0
    same = np.concatenate([tri, tri])                                              # tie -> lower id
    assert oracle.closest_hit(same, (0.2, 0.2, 1.0), (0, 0, -1))[0] == 0


# -------------------------------------------------------------------- convolution
def test_direct_conv_matches_numpy():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(700).astype(np.float32); h = rng.standard_normal(300).astype(np.float32)
    assert np.allclose(oracle.direct_conv(x, h), np.convolve(x.astype(np.float64), h.astype(np.float64)), rtol=1e-12, atol=1e-12)


def test_reference_file_conv_semantics():
    """OR/kernels.cu:382-438 + AudioRenderer.cpp:702-711: unit impulse => 2x;
    circular wrap when support + fs > ir_len; whole seconds only; truncated to n."""
    fs = 1000
    rng = np.random.default_rng(1)
    x = rng.standard_normal(3 * fs + 250).astype(np.float32)
    d = np.zeros(fs, np.float32); d[0] = 1
    y = oracle.reference_file_conv(x, d, fs)
    assert np.allclose(y[: 3 * fs], 2.0 * x[: 3 * fs], atol=1e-12) and np.allclose(y[3 * fs:], 0.0)
    s = np.zeros(fs, np.float32); s[300] = 1                                        # shift wraps inside each second
    y = oracle.reference_file_conv(x, s, fs)
    exp = np.zeros(len(x))
    for k in range(3):
        exp[k * fs:(k + 1) * fs] += 2.0 * np.roll(x[k * fs:(k + 1) * fs].astype(np.float64), 300)
    assert np.allclose(y, exp, atol=1e-12)
    h = np.zeros(2 * fs, np.float32); h[:fs] = rng.standard_normal(fs)              # ir_len = 2 fs, support fs: no wrap
    y = oracle.reference_file_conv(x, h, fs)
    lin = 2.0 * oracle.direct_conv(x[: 3 * fs], h)[: len(x)]
    assert np.allclose(y, lin, atol=1e-9)
    # closed form through the DFT (what cuFFT computes), C1 shape ir_len == fs
    h1 = rng.standard_normal(fs).astype(np.float32)
    y = oracle.reference_file_conv(x[:fs], h1, fs)
    ref = np.fft.irfft(np.fft.rfft(x[:fs].astype(np.float64)) * np.fft.rfft(h1.astype(np.float64)), fs) * 2.0
    assert np.allclose(y, ref, atol=1e-9)


def test_reference_live_conv_and_upola():
    rng = np.random.default_rng(2)
    n = 3000
    hl = rng.standard_normal(n).astype(np.float32); hr = rng.standard_normal(n).astype(np.float32)
    x = rng.standard_normal(256)
    out = oracle.reference_live_conv(x, hl, hr)
    xp = np.zeros(n); xp[:256] = x
    assert np.allclose(out[0::2], 2.0 * np.fft.irfft(np.fft.rfft(xp) * np.fft.rfft(hl.astype(np.float64)), n), atol=1e-9)
    assert np.allclose(out[1::2], 2.0 * np.fft.irfft(np.fft.rfft(xp) * np.fft.rfft(hr.astype(np.float64)), n), atol=1e-9)
    xs = (0.1 * rng.standard_normal(64 * 40)).astype(np.float32)
    ol, orr, _ = oracle.upola(xs, hl[:1000], hr[:1000], 64)
    ref = oracle.direct_conv(xs, hl[:1000])[: len(xs)]
    assert np.linalg.norm(ol - ref) / np.linalg.norm(ref) < 1e-5


def test_direct_conv_window_equals_full():
    rng = np.random.default_rng(4)
    x = rng.standard_normal(3000).astype(np.float32); h = rng.standard_normal(700).astype(np.float32)
    full = oracle.direct_conv(x, h)
    for b, n in ((0, 10), (650, 900), (3000, 699), (len(full) - 1, 1)):
        assert np.array_equal(oracle.direct_conv_window(x, h, b, n), full[b:b + n])


def test_check_parity_per_bin_rule_with_flipped_rays(golden_scenes, golden_receiver):
    """tests/util.check_parity on synthetic 'CUDA' outputs derived from the oracle's own records: two rays of 20 000
    land in another bin (float intersection order) -> the per-bin 1e-4 check still runs on every bin and passes;
    a 3e-4 error in one bin no flipped ray touches, or a third flipped ray, must fail."""
    from util import Case, check_parity, deposit_hist
    case = Case(golden_scenes["test_verts"], golden_scenes["test_mesh"], golden_scenes["test_names"], golden_receiver,
                rays=(100, 100, 2), emitter=(0, 2, 0), center=(5, 2, 0), hrtf=0.7, sample_rate=16000, ir_seconds=1, seed=1)
    o = case.oracle_run()
    ir_len = o["ir_left"].shape[-1]
    # the deposit rule in numpy reproduces the oracle's histogram from its per-ray records
    everything = np.ones(len(o["bin"]), bool)
    assert np.allclose(deposit_hist(case, o, everything, ir_len), o["hist"], rtol=1e-12, atol=0)
    hits = np.nonzero((o["ear"] > 0) & (o["bin"] < ir_len - 40))[0]
    assert len(hits) > 100

    def fake(n_flip, spoil=None):
        rec = {k: v.copy() for k, v in o.items() if k in ("bin", "ear", "energy", "nseg")}
        for j in hits[:n_flip]:
            rec["bin"][j] += 3
        l, r = oracle.finalize_ir(deposit_hist(case, rec, everything, ir_len))
        if spoil is not None:
            l = l.copy(); l[0, spoil] *= np.float32(1.0 + 3e-4)
        return rec, l, r

    rec, l, r = fake(2)
    assert check_parity(rec, l, r, o["segments"], o, case=case) == 1.0 - 2 / 20000
    clean = [b for b in np.nonzero(o["ir_left"][0])[0] if b not in set(o["bin"][hits[:2]]) | set(o["bin"][hits[:2]] + 3)]
    rec, l, r = fake(2, spoil=clean[0])
    with pytest.raises(AssertionError):
        check_parity(rec, l, r, o["segments"], o, case=case)
    rec, l, r = fake(3)
    with pytest.raises(AssertionError):
        check_parity(rec, l, r, o["segments"], o, case=case)


def test_oracle_builds_are_bit_identical(golden_scenes, golden_receiver):
    """BASELINE.md section 5 wants the CPU baseline built -O3 -march=native; the oracle's results must not depend on
    that: the portable -O3 build, the -O2 build and the native build (compiled on this machine) agree bit for bit
    (every fused operation is an explicit fmaf, contraction is off)."""
    from util import Case
    case = Case(golden_scenes["test_verts"], golden_scenes["test_mesh"], golden_scenes["test_names"], golden_receiver,
                rays=(100, 50, 2), emitter=(0, 2, 0), center=(5, 2, 0), hrtf=0.7, sample_rate=16000, ir_seconds=1, seed=3)
    flat, p = case.flat(), case.params()
    tv = np.ascontiguousarray(flat.tri_verts, np.float32); tm = np.ascontiguousarray(flat.tri_mat, np.int32)
    ab = np.ascontiguousarray(flat.absorption, np.float32); sc = np.ascontiguousarray(flat.scattering, np.float32)
    import ctypes as C
    n = 10_000
    outs = []
    for L in (oracle.lib(), oracle.load_variant("o2"), oracle.load_variant("native")):
        hist = np.zeros((2, 1, 16000)); b = np.empty(n, np.int32); e = np.empty(n, np.int32); s = np.empty(n, np.int32)
        en = np.empty((n, 1), np.float32)
        segs = L.oracle_trace(C.byref(p), oracle._fp(tv), oracle._ip(tm), len(tv), oracle._fp(ab), oracle._fp(sc), len(ab), 0, n, 1, 1,
                              oracle._dp(hist), oracle._ip(b), oracle._ip(e), oracle._fp(en), oracle._ip(s))
        x = np.random.default_rng(1).standard_normal(3000).astype(np.float32); h = np.random.default_rng(2).standard_normal(500).astype(np.float32)
        y = np.zeros(3499)
        L.oracle_direct_conv(oracle._fp(x), 3000, oracle._fp(h), 500, oracle._dp(y), 1)
        outs.append((segs, hist, b, e, en, s, y))
    for o in outs[1:]:
        assert o[0] == outs[0][0]
        for a, b in zip(o[1:], outs[0][1:]):
            assert np.array_equal(a, b)
