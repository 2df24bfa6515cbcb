"""Generate the fixtures under tests/golden/ from the read-only reference checkout.

Run in the authoring container only (needs /root/reference and `make -C oracle`):
    python tests/golden/make_golden.py

What is produced and where it comes from (paths relative to /root/reference):
  meshes.json          per shipped OBJ: MTL names, and for every mesh loadOBJ would emit
                       (prebuild/obj_raytracer/OptixModel.cpp:75-151) its material name,
                       triangle count and sha256 of the float32 vertex bits -- produced by
                       the REFERENCE'S OWN loader (oracle/_ref/tinyobj_dump compiles
                       prebuild/common/3rdParty/tiny_obj_loader.h where it lies).
  receiver.npz         the two receiver half-ball templates (assets/models/leftHalf.obj,
                       rightHalf.obj) as flat float32 triangles, from the same dump.
  scenes.npz           flat triangles + mesh ids of test.obj, caja.obj, 3D_U.obj,
                       cajaConToro.obj (the scenes the parity tests trace).
  wav_decode.npz       first 4096 int16 samples of experimento_entrada_16KHz.wav and the
                       matching lines of prebuild/obj_raytracer/input.txt (the reference's
                       own dump of that file) -- pins the WAV decode rule.
  output_ir.json       length and non-zero (index, value) pairs of
                       prebuild/obj_raytracer/output_ir.txt (shape-only known answer).
  guitar_2s.npy        first 2 s of guitar_sample_16k.wav (float32 mono 16 kHz).
  ref_loadobj.json     the same per-mesh digests from the reference's own loadOBJ
                       (prebuild/obj_raytracer/OptixModel.cpp:75-151 compiled where it lies,
                       oracle/_ref/ref_scene_dump): mesh split, order and (v,vn,vt) dedupe.
  ref_placement.npz    receiver triangles as the reference's own placeReceiver /
                       place_receiver_half (OptixModel.cpp:153-257, glm::rotate + mat4*vec4)
                       places them, for several (position, rotation) cases.
  ref_shading.npz      4096 hand-made hits through the reference's own
                       __closesthit__radiance (prebuild/obj_raytracer/devicePrograms.cu:62-180,
                       compiled UNMODIFIED with g++ -ffp-contract=off against oracle/ref_stubs/):
                       inputs and the PRD / deposits it leaves.
  ref_render.npz       the reference's own __raygen__renderFrame + closest-hit + miss programs
                       (devicePrograms.cu:192-254) run on the CPU over BASELINE config 1
                       (test.obj + placed receiver, 100k rays) and over caja.obj (20k rays, 50
                       bounces): per-ray bin / ear / deposited energy / optixTrace calls and the
                       IR.  The ray-triangle search OptiX does in hardware and cuRAND's uniforms are
                       supplied by oracle/ref_device_shim.cpp (documented there).
"""
import hashlib
import json
import os
import struct
import subprocess
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DUMP = os.path.join(ROOT, "oracle", "_ref", "tinyobj_dump")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PLACEMENTS = [((5, 2, 0), 30.0), ((0, 0, 0), 0.0), ((-8, 4, 6), 200.0), ((9.0, 1.4, 5.5), 63.0), ((1.5, -2.25, 3.125), 359.5),
              ((4.0, 3.0, -2.0), -45.0), ((30.0, 1.6, 15.0), 90.0)]


def dump(path):
    out = subprocess.run([DUMP, path], capture_output=True, text=True)
    if out.returncode != 0:
        return None
    mats, meshes = [], []
    for line in out.stdout.splitlines():
        if line.startswith("material "):
            mats.append(line[9:])
        elif line.startswith("mesh"):
            meshes.append([line[5:], []])
        elif line.startswith("t "):
            meshes[-1][1].append([int(x, 16) for x in line.split()[1:]])
    res = []
    for name, tris in meshes:
        a = np.array(tris, dtype=np.uint32).reshape(-1, 3, 3)
        res.append((name, a.view(np.float32)))
    return mats, res


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    files = {}
    models = sorted(os.listdir(os.path.join(REF, "assets/models")))
    paths = [os.path.join("assets/models", f) for f in models if f.endswith(".obj")] + ["test.obj", "monkey.obj"]
    flat = {}
    for rel in paths:
        d = dump(os.path.join(REF, rel))
        if d is None:
            continue
        mats, meshes = d
        if not mats:
            files[rel] = {"error": "could not parse materials ..."}
            continue
        files[rel] = {
            "materials": mats,
            "meshes": [{"material": n, "tris": int(len(t)),
                        "sha256": hashlib.sha256(np.ascontiguousarray(t).view(np.uint32).tobytes()).hexdigest()}
                       for n, t in meshes],
        }
        flat[rel] = meshes
    with open(os.path.join(HERE, "meshes.json"), "w") as fh:
        json.dump(files, fh, indent=1, sort_keys=True)

    np.savez_compressed(os.path.join(HERE, "receiver.npz"),
                        left=flat["assets/models/leftHalf.obj"][0][1], right=flat["assets/models/rightHalf.obj"][0][1])
    sc = {}
    for key, rel in [("test", "test.obj"), ("caja", "assets/models/caja.obj"), ("u3d", "assets/models/3D_U.obj"),
                     ("toro", "assets/models/cajaConToro.obj")]:
        meshes = flat[rel]
        sc[key + "_verts"] = np.concatenate([t for _, t in meshes])
        sc[key + "_mesh"] = np.concatenate([np.full(len(t), i, np.int32) for i, (_, t) in enumerate(meshes)])
        sc[key + "_names"] = np.array([n for n, _ in meshes])
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **sc)

    raw = open(os.path.join(REF, "experimento_entrada_16KHz.wav"), "rb").read()
    pos = raw.find(b"data") + 8
    pcm = np.frombuffer(raw[pos:pos + 2 * 4096], dtype="<i2")
    txt = np.loadtxt(os.path.join(REF, "prebuild/obj_raytracer/input.txt"), max_rows=4096)
    np.savez_compressed(os.path.join(HERE, "wav_decode.npz"), pcm=pcm, text=txt)

    ir = np.loadtxt(os.path.join(REF, "prebuild/obj_raytracer/output_ir.txt"))
    nz = np.nonzero(ir)[0]
    with open(os.path.join(HERE, "output_ir.json"), "w") as fh:
        json.dump({"length": int(len(ir)), "nonzero": [[int(i), float(ir[i])] for i in nz]}, fh)

    g = open(os.path.join(REF, "guitar_sample_16k.wav"), "rb").read()
    pos = g.find(b"data") + 8
    np.save(os.path.join(HERE, "guitar_2s.npy"), np.frombuffer(g[pos:pos + 4 * 32000], dtype="<f4").copy())
    pins(files)
    print("golden fixtures written to", HERE)


def digest(meshes):
    return [{"material": n, "tris": int(len(t)), "sha256": hashlib.sha256(np.ascontiguousarray(t).view(np.uint32).tobytes()).hexdigest()}
            for n, t in meshes]


def shading_inputs(n, seed=0):
    """Hand-made hits for __closesthit__radiance: a random triangle, barycentrics, a ray that arrives at the hit point
    from 0.5..30 m away, a receiver centre within 1.2 m of it (so the unit-ball chord exists), PRD before the hit.
    kind 0 / 1 / 2 = receiver_left / receiver_right / wall with a random absorption."""
    rng = np.random.default_rng(seed)
    out = dict(tri=np.empty((n, 3, 3), np.float32), mat=np.empty(n, np.float32), dir=np.empty((n, 3), np.float32),
               uv=np.empty((n, 2), np.float32), center=np.empty((n, 3), np.float32), prd=np.empty((n, 8), np.float32))
    for i in range(n):
        tri = rng.uniform(-10, 10, (3, 3)).astype(np.float32)
        u = rng.uniform(0, 1); v = rng.uniform(0, 1 - u)
        P = (1 - u - v) * tri[0] + u * tri[1] + v * tri[2]
        d = rng.standard_normal(3); d /= np.linalg.norm(d)
        prev = P - rng.uniform(0.5, 30) * d
        off = rng.standard_normal(3); off *= rng.uniform(0, 1.2) / np.linalg.norm(off)
        out["tri"][i] = tri; out["uv"][i] = (u, v); out["dir"][i] = d; out["center"][i] = P + off
        out["mat"][i] = (-1.0, -2.0, rng.uniform(0, 1))[i % 3]
        out["prd"][i] = [rng.uniform(1e-6, 1e-3), rng.uniform(0, 600), *prev, *d]
    return out


def run_shading(call, inp, sample_rate=48000, hrtf=0.9, mono=False, ir_length=96000, depth=3):
    from oracle import ref
    n = len(inp["mat"])
    prd = np.empty((n, 8), np.float32); dep = np.empty(n, np.int32); ndep = np.empty(n, np.int32)
    ear = np.empty((n, 2), np.int32); idx = np.empty((n, 2), np.int32); val = np.empty((n, 2), np.float32)
    for i in range(n):
        prd[i], dep[i], ndep[i], ear[i], idx[i], val[i] = ref.closesthit(
            call, inp["tri"][i], inp["mat"][i], inp["dir"][i], inp["uv"][i, 0], inp["uv"][i, 1], inp["center"][i], sample_rate, hrtf,
            mono, ir_length, inp["prd"][i], depth)
    return dict(prd=prd, depth=dep, ndep=ndep, ear=ear, idx=idx, val=val)


RENDER_CASES = {
    # BASELINE config 1: test.obj + the receiver at (5,2,0), 100x100x10 rays, 1 s IR @16 kHz
    "c1": dict(obj="test.obj", rays=(100, 100, 10), emitter=(0, 2, 0), center=(5, 2, 0), yaw=0.0, base_power=3.62, max_bounces=100,
               hrtf=0.9, sample_rate=16000, ir_length=16000, seed=1, absorption=0.5),
    # closed double-walled cube: every ray bounces until the receiver, the bounce limit or the length cap ends it
    "caja": dict(obj="assets/models/caja.obj", rays=(100, 100, 2), emitter=(3, 1, -2), center=(-8, 4, 6), yaw=20.0, base_power=3.62,
                 max_bounces=50, hrtf=0.9, sample_rate=16000, ir_length=32000, seed=5, absorption=0.2),
}


def render_case_scene(c):
    """Flat scene of a render case through the reference's own loadOBJ + placeReceiver."""
    from oracle import ref
    meshes = ref.scene_dump(os.path.join(REF, c["obj"]), (os.path.join(REF, "assets/models/leftHalf.obj"),
                                                         os.path.join(REF, "assets/models/rightHalf.obj"), c["center"], c["yaw"]))
    tv = np.concatenate([t for _, t in meshes])
    tm, n_wall = [], 0
    for n, t in meshes:
        if n in ("receiver_left", "receiver_right"):
            tm.append(np.full(len(t), -1 if n == "receiver_left" else -2, np.int32))
        else:
            tm.append(np.full(len(t), n_wall, np.int32)); n_wall += 1
    return tv, np.concatenate(tm), np.full(n_wall, c["absorption"], np.float32)


def run_render(c):
    from oracle import ref
    tv, tm, ab = render_case_scene(c)
    n = c["rays"][0] * c["rays"][1] * c["rays"][2]
    un = ref.uniforms_for_rays(c["seed"], 0, n)
    r = ref.render(tv, tm, ab, c["rays"], c["emitter"], c["center"], c["base_power"], 0.0, c["max_bounces"], c["hrtf"],
                   c["sample_rate"], False, c["ir_length"], un)
    return (tv, tm, ab), r


def pins(files):
    """Vectors from the reference's own code compiled where it lies (oracle/_ref, oracle/ref.py)."""
    from oracle import ref
    out = {}
    for rel in sorted(files):
        if "error" in files[rel]:
            continue
        out[rel] = digest(ref.scene_dump(os.path.join(REF, rel)))
    with open(os.path.join(HERE, "ref_loadobj.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)

    pl = {}
    for k, (cam, rot) in enumerate(PLACEMENTS):
        m = ref.scene_dump(os.path.join(REF, "test.obj"), (os.path.join(REF, "assets/models/leftHalf.obj"),
                                                          os.path.join(REF, "assets/models/rightHalf.obj"), cam, rot))
        assert [n for n, _ in m[-2:]] == ["receiver_left", "receiver_right"]
        pl[f"left_{k}"] = m[-2][1]; pl[f"right_{k}"] = m[-1][1]
    pl["cases"] = np.array([[*cam, rot] for cam, rot in PLACEMENTS], np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_placement.npz"), **pl)

    inp = shading_inputs(4096)
    res = run_shading(ref.lib().ref_closesthit, inp)
    np.savez_compressed(os.path.join(HERE, "ref_shading.npz"), **{"in_" + k: v for k, v in inp.items()}, **{"out_" + k: v for k, v in res.items()})

    rr = {}
    for name, c in RENDER_CASES.items():
        _, r = run_render(c)
        rr[name + "_bin"] = r["bin"]; rr[name + "_ear"] = r["ear"].astype(np.int8); rr[name + "_nseg"] = r["nseg"].astype(np.int16)
        rr[name + "_energy"] = r["energy"][:, 0]; rr[name + "_segments"] = np.int64(r["segments"])
        nz = np.nonzero(r["hist"].reshape(-1))[0]
        rr[name + "_hist_idx"] = nz.astype(np.int32); rr[name + "_hist_val"] = r["hist"].reshape(-1)[nz]
    np.savez_compressed(os.path.join(HERE, "ref_render.npz"), **rr)


if __name__ == "__main__":
    sys.exit(main())
