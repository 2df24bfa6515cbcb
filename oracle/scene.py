"""CPU ORACLE (test infrastructure, not the product) -- host-side front end.

numpy restatement of the reference's scene/config handling, independent of the
product's C++ front end (audiorenderingv2_b200/csrc/host):

* ``load_obj``       OR/OptixModel.cpp:75-151 (loadOBJ) on top of the vendored
                     tinyobjloader v2.0.0rc (prebuild/common/3rdParty/tiny_obj_loader.h:
                     parse loop :2134-2640, ear-clip triangulation :1385-1593,
                     LoadMtl flush rules).  Pinned against that very header through
                     oracle/_ref/tinyobj_dump -> tests/golden/*.mesh.
* ``place_receiver`` OR/OptixModel.cpp:153-257 (glm::rotate about +Y by -angle, then
                     translate by the camera position).
* ``load_config``    OR/Context.cpp:17-165 (defaults and rounding quirks).
* ``material_absorption`` OR/AudioRenderer.cpp:34-56.
* ``read_wav``       AudioFile.h:1242-1245 int16 -> /32768, :617-618 float32 passthrough.

PARITY UNPINNED for the tracer/convolver (see arv2_oracle.h); the mesh front end
*is* pinned (golden meshes generated from the reference's own loader).
"""
from __future__ import annotations

import json
import math
import os
import struct
from dataclasses import dataclass, field

import numpy as np

F = np.float32


# --------------------------------------------------------------------------- OBJ
@dataclass
class Mesh:
    material_name: str
    tris: np.ndarray  # float32 [n,3,3]


@dataclass
class Model:
    meshes: list = field(default_factory=list)
    materials: list = field(default_factory=list)  # MTL names in file order


def _load_mtl_names(path):
    """tinyobj LoadMtl: a material is flushed on `newmtl` only if the previous one
    has a name; the last one is always pushed (so an empty .mtl yields [''])."""
    names = []
    cur = ""
    try:
        with open(path, "r", errors="replace") as fh:
            for line in fh:
                tok = line.strip()
                if tok.startswith("newmtl") and len(tok) > 6 and tok[6] in " \t":
                    if cur != "":
                        names.append(cur)
                    cur = tok[7:].strip()
    except OSError:
        return None
    names.append(cur)
    return names


def _fix_index(idx, n):
    # tinyobj fixIndex: 1-based -> 0-based, negative = relative to the current count
    if idx > 0:
        return idx - 1
    if idx == 0:
        raise ValueError("zero index in OBJ face")
    return n + idx


def _pnpoly3(vx, vy, tx, ty):
    c = False
    j = 2
    for i in range(3):
        if (vy[i] > ty) != (vy[j] > ty):
            if tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i]:
                c = not c
        j = i
    return c


def _triangulate(face, v):
    """tiny_obj_loader.h:1385-1593 in float32 (real_t == float)."""
    n = len(face)
    if n < 3:
        return []
    if n == 3:
        return [tuple(face)]
    axes = [1, 2]
    eps = F(np.finfo(np.float32).eps)
    for k in range(n):
        p0, p1, p2 = (v[face[(k + d) % n]] for d in range(3))
        e0 = p1 - p0
        e1 = p2 - p1
        cx = abs(F(e0[1] * e1[2]) - F(e0[2] * e1[1]))
        cy = abs(F(e0[2] * e1[0]) - F(e0[0] * e1[2]))
        cz = abs(F(e0[0] * e1[1]) - F(e0[1] * e1[0]))
        if cx > eps or cy > eps or cz > eps:
            if not (cx > cy and cx > cz):
                axes[0] = 0
                if cz > cx and cz > cy:
                    axes[1] = 1
            break
    area = F(0)
    for k in range(n):
        a = v[face[k]]
        b = v[face[(k + 1) % n]]
        area = F(area + F(F(F(a[axes[0]] * b[axes[1]]) - F(a[axes[1]] * b[axes[0]])) * F(0.5)))
    out = []
    rem = list(face)
    guess = 0
    remaining_iter = n
    prev_remaining = n
    while len(rem) > 3 and remaining_iter > 0:
        m = len(rem)
        if guess >= m:
            guess -= m
        if prev_remaining != m:
            prev_remaining = m
            remaining_iter = m
        else:
            remaining_iter -= 1
        ind = [rem[(guess + k) % m] for k in range(3)]
        vx = [v[i][axes[0]] for i in ind]
        vy = [v[i][axes[1]] for i in ind]
        e0x = F(vx[1] - vx[0]); e0y = F(vy[1] - vy[0])
        e1x = F(vx[2] - vx[1]); e1y = F(vy[2] - vy[1])
        cross = F(F(e0x * e1y) - F(e0y * e1x))
        if F(cross * area) < 0:
            guess += 1
            continue
        overlap = False
        for other in range(3, m):
            idx = (guess + other) % m
            o = rem[idx]
            if _pnpoly3(vx, vy, v[o][axes[0]], v[o][axes[1]]):
                overlap = True
                break
        if overlap:
            guess += 1
            continue
        out.append(tuple(ind))
        del rem[(guess + 1) % m]
    if len(rem) == 3:
        out.append(tuple(rem))
    return out


def load_obj(path: str) -> Model:
    """loadOBJ: one mesh per (shape, material id), material ids ascending (std::set),
    faces in file order, shapes in file order."""
    mtl_dir = path[: path.rfind("/") + 1]
    v = []
    materials = []
    material_map = {}
    shapes = []  # each: list of (tri(3 vertex ids), material id)
    shape = []
    faces = []  # pending prim group: list of vertex-id lists
    material = -1

    def export():
        nonempty = len(faces) > 0
        for f in faces:
            for t in _triangulate(f, v):
                shape.append((t, material))
        return nonempty

    with np.errstate(all="ignore"), open(path, "r", errors="replace") as fh:
        for raw in fh:
            line = raw.rstrip("\n").rstrip("\r")
            tok = line.lstrip(" \t")
            if not tok or tok[0] == "#":
                continue
            if tok[0] == "v" and len(tok) > 1 and tok[1] in " \t":
                p = tok[2:].split()
                xyz = [float(p[i]) if i < len(p) else 0.0 for i in range(3)]
                v.append(np.array(xyz, dtype=np.float32))
            elif tok[0] == "f" and len(tok) > 1 and tok[1] in " \t":
                ids = []
                for w in tok[2:].split():
                    ids.append(_fix_index(int(w.split("/")[0]), len(v)))
                faces.append(ids)
            elif tok.startswith("usemtl") and len(tok) > 6 and tok[6] in " \t":
                name = tok[7:]
                new_id = material_map.get(name, -1)
                if new_id != material:
                    export()
                    faces.clear()
                    material = new_id
            elif tok.startswith("mtllib") and len(tok) > 6 and tok[6] in " \t":
                for fn in tok[7:].split(" "):
                    if not fn:
                        continue
                    names = _load_mtl_names(mtl_dir + fn)
                    if names is not None:
                        for nm in names:
                            material_map.setdefault(nm, len(materials))
                            materials.append(nm)
                        break
            elif tok[0] == "g" and len(tok) > 1 and tok[1] in " \t":
                export()
                if shape:
                    shapes.append(shape)
                shape = []
                faces.clear()
            elif tok[0] == "o" and len(tok) > 1 and tok[1] in " \t":
                if export():
                    shapes.append(shape)
                shape = []
                faces.clear()
        ret = export()
        if ret or shape:
            shapes.append(shape)

    if not materials:
        raise RuntimeError("could not parse materials ...")  # OptixModel.cpp:99-100
    model = Model(materials=materials)
    varr = np.stack(v) if v else np.zeros((0, 3), np.float32)
    for shp in shapes:
        for mid in sorted({m for _, m in shp}):
            tris = [t for t, m in shp if m == mid]
            if not tris:
                continue
            idx = np.array(tris, dtype=np.int64)
            model.meshes.append(Mesh(materials[mid] if mid >= 0 else "", varr[idx].astype(np.float32)))
    return model


# ----------------------------------------------------------------------- receiver
@dataclass
class ReceiverTemplate:
    left: np.ndarray   # float32 [510,3,3] untransformed (leftHalf.obj)
    right: np.ndarray


def load_receiver(asset_dir: str) -> ReceiverTemplate:
    """HalfSphere ctor (OR/HalfSphere.cpp): shapes[0] of leftHalf.obj / rightHalf.obj."""
    l = load_obj(os.path.join(asset_dir, "leftHalf.obj"))
    r = load_obj(os.path.join(asset_dir, "rightHalf.obj"))
    return ReceiverTemplate(l.meshes[0].tris, r.meshes[0].tris)


def c_round(x: float) -> float:
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def rotation_coeffs(rotation_deg: float):
    """glm::rotate(mat4(1), -radians(rot), (0,1,0)) -> (c, s, k=c+(1-c)) as float32.
    cos/sin are taken in fp64 and narrowed (arithmetic contract)."""
    ang = F(F(rotation_deg) * F(0.01745329251994329576923690768489))
    a = F(-ang)
    c = F(math.cos(float(a)))
    s = F(math.sin(float(a)))
    k = F(c + F(F(1) - c))
    return c, s, k


def place_receiver_half(tris: np.ndarray, cam, rotation_deg: float) -> np.ndarray:
    """OptixModel.cpp:162-196: v' = cam + M*v with glm's mat4*vec4 evaluation order
    (Mul0+Mul1)+(Mul2+Mul3)."""
    c, s, k = rotation_coeffs(rotation_deg)
    x = tris[..., 0].astype(np.float32)
    y = tris[..., 1].astype(np.float32)
    z = tris[..., 2].astype(np.float32)
    zero = F(0)
    with np.errstate(all="ignore"):
        xr = (c * x + zero * y) + (s * z + zero)
        yr = (zero * x + k * y) + (zero * z + zero)
        zr = ((-s) * x + zero * y) + (c * z + zero)
        out = np.stack([F(cam[0]) + xr, F(cam[1]) + yr, F(cam[2]) + zr], axis=-1)
    return out.astype(np.float32)


# ------------------------------------------------------------------------ config
DEFAULTS = dict(
    initial_volume=1.0, ir_length_in_seconds=2, width=1366, height=768,
    write_first_ir_to_file=False, write_first_output_to_file=False,
    re_render_distance_threshold=3.0, re_render_angle_threshold=5.0,
    mono=False, scene_file_path="../../assets/models/1D_U.obj", audio_file_path="",
    materials_file_path="", initial_receiver_pos=(-2.5, 10.0, 0.0), initial_emitter_pos=(0.0, 0.0, 0.0),
    base_power=100.0, rays=(100.0, 100.0, 100.0), ray_energy_threshold=0.0, ray_max_bounces=10,
    hrtf_absorption_rate=0.9, materials=[],
)


def _isnum(x):
    return isinstance(x, (int, float)) and not isinstance(x, bool)


def load_config(text: str) -> dict:
    """Context::loadContext (OR/Context.cpp:15-165) incl. its rounding quirks:
    ir_length_in_seconds, width, height, both re_render thresholds, ray_max_bounces
    and hrtf_absorption_rate are round()ed when present."""
    cfg = json.loads(text)
    out = dict(DEFAULTS)
    out["materials"] = []
    rp = cfg.get("renderer_parameters")
    if isinstance(rp, dict):
        if _isnum(rp.get("initial_volume")):
            out["initial_volume"] = float(F(rp["initial_volume"]))
        for key in ("ir_length_in_seconds", "width", "height"):
            if _isnum(rp.get(key)):
                out[key] = int(c_round(rp[key]))
        for key in ("write_first_ir_to_file", "write_first_output_to_file"):
            if isinstance(rp.get(key), bool):
                out[key] = rp[key]
        for key in ("re_render_distance_threshold", "re_render_angle_threshold"):
            if _isnum(rp.get(key)):
                out[key] = float(c_round(rp[key]))
    sp = cfg.get("scene_parameters")
    if isinstance(sp, dict):
        if isinstance(sp.get("mono"), bool):
            out["mono"] = sp["mono"]
        for key in ("scene_file_path", "audio_file_path", "materials_file_path"):
            if isinstance(sp.get(key), str):
                out[key] = sp[key]
        for key in ("initial_receiver_pos", "initial_emitter_pos"):
            p = sp.get(key)
            if isinstance(p, dict) and all(_isnum(p.get(a)) for a in "xyz"):
                out[key] = tuple(float(F(p[a])) for a in "xyz")
    pp = cfg.get("pathtracer_parameters")
    if isinstance(pp, dict):
        if _isnum(pp.get("base_power")):
            out["base_power"] = float(F(pp["base_power"]))
        r = pp.get("rays")
        if isinstance(r, dict) and all(_isnum(r.get(a)) for a in "xyz"):
            out["rays"] = tuple(float(F(r[a])) for a in "xyz")
        if _isnum(pp.get("ray_energy_threshold")):
            out["ray_energy_threshold"] = float(F(pp["ray_energy_threshold"]))
        if _isnum(pp.get("ray_max_bounces")):
            out["ray_max_bounces"] = int(c_round(pp["ray_max_bounces"]))
        if _isnum(pp.get("hrtf_absorption_rate")):
            out["hrtf_absorption_rate"] = float(c_round(pp["hrtf_absorption_rate"]))  # Context.cpp:143-145
        mats = pp.get("materials")
        if isinstance(mats, list):
            for m in mats:
                if isinstance(m, dict) and isinstance(m.get("name"), str) and _isnum(m.get("mat_absorption")):
                    out["materials"].append((m["name"], float(F(m["mat_absorption"]))))
    return out


def material_absorption(name: str, materials) -> float:
    """getMaterialAbsorption, OR/AudioRenderer.cpp:34-56."""
    if name == "receiver_left":
        return -1.0
    if name == "receiver_right":
        return -2.0
    for n, a in materials:
        if n == name:
            return a
    return 0.5


# -------------------------------------------------------------------- flat scene
@dataclass
class FlatScene:
    tri_verts: np.ndarray   # float32 [T,3,3]
    tri_mat: np.ndarray     # int32 [T]
    absorption: np.ndarray  # float32 [M,bands]
    scattering: np.ndarray  # float32 [M]
    n_scene_tris: int


def flatten(model: Model, receiver: ReceiverTemplate, cam, rotation_deg, materials, bands=1,
            scattering=0.0) -> FlatScene:
    """Scene meshes in loadOBJ order, then receiver_left, then receiver_right
    (placeReceiver pushes them last, OptixModel.cpp:238-254)."""
    verts, mats, absorb = [], [], []
    for mi, mesh in enumerate(model.meshes):
        verts.append(mesh.tris)
        mats.append(np.full(len(mesh.tris), mi, np.int32))
        a = material_absorption(mesh.material_name, materials)
        absorb.append([a] * bands if _isnum(a) else list(a))
    n_scene = sum(len(m.tris) for m in model.meshes)
    if receiver is not None:
        l = place_receiver_half(receiver.left, cam, rotation_deg)
        r = place_receiver_half(receiver.right, cam, rotation_deg)
        verts += [l, r]
        mats += [np.full(len(l), -1, np.int32), np.full(len(r), -2, np.int32)]
    tv = np.concatenate(verts).astype(np.float32) if verts else np.zeros((0, 3, 3), np.float32)
    tm = np.concatenate(mats).astype(np.int32) if mats else np.zeros((0,), np.int32)
    ab = np.array(absorb, dtype=np.float32).reshape(len(model.meshes), bands)
    sc = np.full(len(model.meshes), scattering, np.float32)
    return FlatScene(np.ascontiguousarray(tv), np.ascontiguousarray(tm), ab, sc, n_scene)


# --------------------------------------------------------------------------- WAV
def read_wav(path: str):
    """Minimal RIFF/WAVE reader following AudioFile.h's decode rule: int16 ->
    sample/32768 (AudioFile.h:1242-1245), IEEE float32 passthrough (:617-618).
    Returns (sample_rate, float32 [channels, frames])."""
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos = 12
    fmt = None
    pcm = None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        size = struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError("missing fmt/data chunk")
    tag, ch, rate, _, _, bits = fmt
    if tag == 1 and bits == 16:
        a = np.frombuffer(pcm[: len(pcm) // 2 * 2], dtype="<i2").astype(np.float32) / F(32768.0)
    elif tag == 3 and bits == 32:
        a = np.frombuffer(pcm[: len(pcm) // 4 * 4], dtype="<f4").astype(np.float32)
    else:
        raise ValueError(f"unsupported WAV format tag={tag} bits={bits}")
    frames = len(a) // ch
    return rate, np.ascontiguousarray(a[: frames * ch].reshape(frames, ch).T)
