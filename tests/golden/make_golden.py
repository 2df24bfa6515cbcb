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
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    sys.exit(main())
