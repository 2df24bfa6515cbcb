"""The C++ side, run for real: audiorenderingv2_b200/lib/arv2_cli (csrc/host/arv2_cli.cpp over csrc/host/audio_renderer.hpp,
the mirror of class AudioRenderer) executes the reference's export sequence (OR/main.cpp:653-718: loadContext -> loadOBJ
-> receiver halves -> AudioRenderer -> setters -> render -> convoluteAudioFile -> Result.wav) from a config.json of the
reference's schema, and its files -- output_ir_left/right.txt, output_convolute_left/right.txt, Result.wav -- are
checked against the same pipeline through the oracle.  The reference checkout is not on the GPU box, so the scene
(3D_U.obj, the shipped config's scene), the receiver halves and the dry signal are written out from tests/golden."""
import json
import os
import subprocess

import numpy as np
import pytest

import audiorenderingv2_b200 as arv
import oracle
from oracle import scene as osc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CLI = os.path.join(os.path.dirname(arv.LIB_PATH), "arv2_cli")


def write_obj(path, tris, mesh_ids, names, with_mtl=True):
    """Flat triangles -> OBJ + MTL (one `usemtl` group per mesh, 9 significant digits: float32 round-trips)."""
    base = os.path.splitext(os.path.basename(path))[0]
    with open(path, "w") as fh:
        if with_mtl:
            fh.write(f"mtllib {base}.mtl\n")
        for t in tris.reshape(-1, 3):
            fh.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in t))
        fh.write("o all\n")
        cur = None
        for i in range(len(tris)):
            if with_mtl and mesh_ids[i] != cur:
                cur = mesh_ids[i]
                fh.write(f"usemtl {names[cur]}\n")
            fh.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
    if with_mtl:
        with open(os.path.join(os.path.dirname(path), base + ".mtl"), "w") as fh:
            for n in dict.fromkeys(names):
                fh.write(f"newmtl {n}\nKd 0.5 0.5 0.5\n")


def write_wav_i16(path, pcm, fs):
    pcm = np.asarray(pcm, "<i2")
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + (36 + 2 * len(pcm)).to_bytes(4, "little") + b"WAVEfmt " + (16).to_bytes(4, "little")
                 + (1).to_bytes(2, "little") + (1).to_bytes(2, "little") + int(fs).to_bytes(4, "little")
                 + int(2 * fs).to_bytes(4, "little") + (2).to_bytes(2, "little") + (16).to_bytes(2, "little")
                 + b"data" + (2 * len(pcm)).to_bytes(4, "little") + pcm.tobytes())


@pytest.fixture
def workdir(tmp_path, golden_scenes, golden_receiver):
    """assets/ and config.json as the reference lays them out (config.json:1-61), from the golden triangles."""
    assets = tmp_path / "assets"; assets.mkdir()
    names = [str(n) for n in golden_scenes["u3d_names"]]
    write_obj(str(assets / "3D_U.obj"), golden_scenes["u3d_verts"], golden_scenes["u3d_mesh"], names)
    write_obj(str(assets / "leftHalf.obj"), golden_receiver[0], np.zeros(len(golden_receiver[0]), int), ["half"])
    write_obj(str(assets / "rightHalf.obj"), golden_receiver[1], np.zeros(len(golden_receiver[1]), int), ["half"])
    fs = 16000
    x = np.load(os.path.join(GOLDEN, "guitar_2s.npy"))
    pcm = np.round(x / np.abs(x).max() * 20000).astype(np.int16)
    write_wav_i16(str(assets / "dry.wav"), pcm, fs)
    cfg = {
        "renderer_parameters": {"initial_volume": 1, "ir_length_in_seconds": 2, "width": 1366, "height": 768,
                                "write_first_ir_to_file": True, "write_first_output_to_file": True, "re_render_distance_threshold": 2},
        "scene_parameters": {"mono": False, "audio_file_path": str(assets / "dry.wav"), "scene_file_path": str(assets / "3D_U.obj"),
                             "materials_file_path": "", "initial_receiver_pos": {"x": 2.5, "y": 9.9, "z": 0.0},
                             "initial_emitter_pos": {"x": 0.0, "y": 0.0, "z": 0.0}},
        "pathtracer_parameters": {"base_power": 3.62, "rays": {"x": 100, "y": 100, "z": 10}, "ray_distance_threshold": 2000.0,
                                  "ray_energy_threshold": 0.0, "ray_max_bounces": 100, "hrtf_absorption_rate": 0.9, "seed": 3,
                                  "materials": [{"name": "low", "mat_absorption": 0.1}, {"name": names[0], "mat_absorption": 0.3}]},
    }
    (tmp_path / "config.json").write_text(json.dumps(cfg, indent=1))
    return tmp_path, cfg, pcm.astype(np.float32) / 32768.0, fs, names


def oracle_pipeline(cfg, golden_scenes, golden_receiver, x, fs, names):
    """loadContext quirks (hrtf rounded to 1.0, OR/Context.cpp:143-145) -> trace -> reference file convolver -> export."""
    pt = cfg["pathtracer_parameters"]; sp = cfg["scene_parameters"]
    model = osc.Model(meshes=[osc.Mesh(names[i], golden_scenes["u3d_verts"][golden_scenes["u3d_mesh"] == i]) for i in range(len(names))])
    recv = tuple(sp["initial_receiver_pos"][k] for k in "xyz"); emit = tuple(sp["initial_emitter_pos"][k] for k in "xyz")
    mats = [(m["name"], m["mat_absorption"]) for m in pt["materials"]]
    flat = osc.flatten(model, osc.ReceiverTemplate(*golden_receiver), recv, 0.0, mats)
    ir_len = 2 * fs
    p = oracle.make_params(rays=(100, 100, 10), emitter=emit, sphere_center=recv, base_power=pt["base_power"], max_bounces=100,
                           hrtf=float(round(pt["hrtf_absorption_rate"])), sample_rate=fs, ir_length=ir_len, seed=pt["seed"])
    o = oracle.trace(p, flat)
    l, r = oracle.finalize_ir(o["hist"])
    yl = oracle.reference_file_conv(x, l[0], fs); yr = oracle.reference_file_conv(x, r[0], fs)
    return o, l[0], r[0], yl, yr


def read_wav_stereo_i16(path):
    raw = open(path, "rb").read()
    pos = raw.find(b"data") + 8
    a = np.frombuffer(raw[pos:], "<i2").reshape(-1, 2)
    return a[:, 0], a[:, 1]


def normalise_i16(v):
    """normalizeToRangeMinusOneToOne + AudioFile's 16-bit quantisation (OR/main.cpp:628-651, 704-717)."""
    v = np.asarray(v, np.float64)
    n = 2.0 * (v - v.min()) / (v.max() - v.min()) - 1.0
    return np.clip(np.round(n * 32767.0), -32768, 32767)


def run_cli(tmp, *extra):
    out = subprocess.run([CLI, str(tmp / "config.json"), "export", str(tmp / "Result.wav"), str(tmp / "assets"), *extra],
                         cwd=tmp, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    return out.stdout


def check_outputs(tmp, o, l, r, yl, yr):
    got_l = np.array([float(s) for s in open(tmp / "output_ir_left.txt")])          # utils/printIR.py:9-12
    got_r = np.array([float(s) for s in open(tmp / "output_ir_right.txt")])
    assert got_l.shape == l.shape and np.array_equal(got_l != 0, l != 0)
    assert np.allclose(got_l, l, rtol=2e-5, atol=0) and np.allclose(got_r, r, rtol=2e-5, atol=0)        # 6 significant digits
    cl = np.array([float(s) for s in open(tmp / "output_convolute_left.txt")])      # utils/main.py:17-28
    cr = np.array([float(s) for s in open(tmp / "output_convolute_right.txt")])
    assert len(cl) == len(yl)
    assert np.linalg.norm(cl - yl) <= 2e-5 * np.linalg.norm(yl) and np.linalg.norm(cr - yr) <= 2e-5 * np.linalg.norm(yr)
    wl, wr = read_wav_stereo_i16(tmp / "Result.wav")
    assert len(wl) == len(yl)
    assert np.abs(wl - normalise_i16(yl)).max() <= 2 and np.abs(wr - normalise_i16(yr)).max() <= 2


def test_cli_export_matches_the_oracle_pipeline(workdir, golden_scenes, golden_receiver):
    tmp, cfg, x, fs, names = workdir
    stdout = run_cli(tmp)
    o, l, r, yl, yr = oracle_pipeline(cfg, golden_scenes, golden_receiver, x, fs, names)
    assert f"({o['segments']} segments" in stdout and "Time taken just to convolute" in stdout
    check_outputs(tmp, o, l, r, yl, yr)
    assert (l != 0).sum() > 50


def test_cli_two_gpus_gives_the_same_files(workdir, golden_scenes, golden_receiver):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    tmp, cfg, x, fs, names = workdir
    stdout = run_cli(tmp, "--gpus", "2")
    o, l, r, yl, yr = oracle_pipeline(cfg, golden_scenes, golden_receiver, x, fs, names)
    assert f"({o['segments']} segments, 2 GPUs)" in stdout
    check_outputs(tmp, o, l, r, yl, yr)


def test_live_input_through_the_mirror_semantics(golden_scenes, golden_receiver):
    """convoluteLiveInput(double*, bytes, CircularBuffer*) of the C++ mirror = arv2_stream (1 source, 512-sample blocks)
    fed with the renderer's device IR + arv2_live_callback: after a render, a 4096-frame callback leaves in the ring what
    the reference's own convoluteLiveInput (restated in oracle.reference_live_conv) would."""
    from util import Case
    case = Case(golden_scenes["test_verts"], golden_scenes["test_mesh"], golden_scenes["test_names"], golden_receiver,
                rays=(100, 100, 2), emitter=(0, 2, 0), center=(5, 2, 0), hrtf=0.9, sample_rate=16000, ir_seconds=1, seed=1)
    r = case.renderer()
    r.render()
    l, rr = r.get_ir()
    st = arv.ConvStream(1, 512, r.ir_length)
    st.set_ir_device(0, *r.ir_device())
    ring = arv.Ring(2 * r.ir_length)
    x = np.random.default_rng(3).standard_normal(4096)
    arv.live_callback(st, x, ring)
    got = ring.get_and_reset(2 * 4096)
    ref = oracle.reference_live_conv(x, l[0], rr[0])[: 2 * 4096]       # support of this IR + 4096 < ir_len: no wrap
    assert np.linalg.norm(got - ref) <= 1e-5 * np.linalg.norm(ref)
