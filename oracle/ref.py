"""oracle.ref -- TEST INFRASTRUCTURE: the reference's OWN sources for the hot path, compiled where they lie under
/root/reference by `make -C oracle ref` into oracle/_ref/ (see oracle/Makefile, ref_scene_dump.cpp,
ref_device_shim.cpp, ref_stubs/).  Used to PIN the oracle: tests/golden/make_golden.py writes the committed vectors
from it, and tests/test_pin_cpu.py re-runs it live wherever the reference checkout is mounted.

Nothing here is importable by the product, and nothing here exists on a machine without /root/reference (the GPU box
uses the prebuilt oracle/_ref/ files that travel with the snapshot, or the committed golden vectors).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
REFERENCE = "/root/reference"
_lib = None


def available() -> bool:
    """The compiled reference artefacts exist (built here, or shipped with the snapshot)."""
    return all(os.path.exists(os.path.join(_REF, f)) for f in ("ref_scene_dump", "libref_device.so"))


def reference_mounted() -> bool:
    return os.path.isfile(os.path.join(REFERENCE, "prebuild", "obj_raytracer", "devicePrograms.cu"))


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(_REF, "libref_device.so"))
        fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)
        L.ref_closesthit.argtypes = [fp, C.c_float, fp, C.c_float, C.c_float, fp, C.c_int, C.c_float, C.c_int, C.c_int,
                                     fp, ip, ip, ip, ip, fp]
        L.ref_render.restype = C.c_int64
        L.ref_render.argtypes = [fp, ip, C.c_int64, fp, C.c_int, C.c_int, C.c_int, fp, fp, C.c_float, C.c_float, C.c_uint,
                                 C.c_float, C.c_int, C.c_int, C.c_int, fp, C.c_int64, C.c_int64, C.c_int, dp, ip, ip, fp, ip]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def scene_dump(obj_path, receiver=None):
    """The reference's loadOBJ (OR/OptixModel.cpp:75-151) and, with receiver = (left_obj, right_obj, cam3, rotation_deg),
    its placeReceiver (:153-257).  Returns [(material_name, float32 [n,3,3])] in model->meshes order."""
    cmd = [os.path.join(_REF, "ref_scene_dump"), obj_path]
    if receiver is not None:
        left, right, cam, rot = receiver
        cmd += [left, right] + [repr(float(np.float32(v))) for v in cam] + [repr(float(np.float32(rot)))]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError(out.stderr.strip())
    meshes = []
    for line in out.stdout.splitlines():
        if line.startswith("mesh"):
            meshes.append([line[5:], []])
        elif line.startswith("t "):
            meshes[-1][1].append([int(x, 16) for x in line.split()[1:]])
    return [(n, np.array(t, dtype=np.uint32).reshape(-1, 3, 3).view(np.float32)) for n, t in meshes]


def closesthit(call, tri, mat_absorption, ray_dir, u, v, sphere_center, sample_rate, hrtf, mono, ir_length, prd8, depth):
    """One closest-hit invocation through `call` (lib().ref_closesthit or oracle.lib().oracle_shade_hit: same flat
    signature).  Returns (prd8 after, depth after, n deposits, ears[2], idx[2], val[2])."""
    tri = np.ascontiguousarray(tri, np.float32); d = np.ascontiguousarray(ray_dir, np.float32)
    c = np.ascontiguousarray(sphere_center, np.float32); p = np.array(prd8, np.float32)
    dep = C.c_int32(int(depth)); n = C.c_int32(0)
    ear = np.zeros(2, np.int32); idx = np.zeros(2, np.int32); val = np.zeros(2, np.float32)
    call(_f(tri), float(mat_absorption), _f(d), float(u), float(v), _f(c), int(sample_rate), float(hrtf), 1 if mono else 0,
         int(ir_length), _f(p), C.byref(dep), C.byref(n), _i(ear), _i(idx), _f(val))
    return p, dep.value, n.value, ear, idx, val


def render(tri_verts, tri_mat, absorption, rays, emitter, sphere_center, base_power, energy_thres, max_bounces, hrtf,
           sample_rate, mono, ir_length, uniforms, ray_begin=0, n_threads=None):
    """The reference's __raygen__renderFrame + closest-hit / miss programs over a flat scene for the launch indices
    [ray_begin, ray_begin + len(uniforms)).  uniforms: float32 [n,2] = curand_uniform's values for theta and phi."""
    tv = np.ascontiguousarray(tri_verts, np.float32); tm = np.ascontiguousarray(tri_mat, np.int32)
    ab = np.ascontiguousarray(absorption, np.float32).reshape(-1)
    un = np.ascontiguousarray(uniforms, np.float32)
    n = un.shape[0]
    e = np.ascontiguousarray(emitter, np.float32); c = np.ascontiguousarray(sphere_center, np.float32)
    hist = np.zeros((2, 1, ir_length), np.float64)
    b = np.empty(n, np.int32); ear = np.empty(n, np.int32); s = np.empty(n, np.int32); en = np.empty((n, 1), np.float32)
    calls = lib().ref_render(_f(tv), _i(tm), tv.shape[0], _f(ab), int(rays[0]), int(rays[1]), int(rays[2]), _f(e), _f(c),
                             float(base_power), float(energy_thres), int(max_bounces), float(hrtf), int(sample_rate),
                             1 if mono else 0, int(ir_length), _f(un), int(ray_begin), n, n_threads or (os.cpu_count() or 1),
                             hist.ctypes.data_as(C.POINTER(C.c_double)), _i(b), _i(ear), _f(en), _i(s))
    return dict(hist=hist, bin=b, ear=ear, energy=en, nseg=s, segments=int(calls))


def philox4x32_10(seed, ray, bounce=0, purpose=0):
    """Philox4x32-10 in numpy (Salmon et al. 2011), vectorised over `ray`: an implementation independent of the
    oracle's and the product's.  Returns uint32 [n,4]."""
    ray = np.asarray(ray, np.uint64)
    c = [(ray & np.uint64(0xFFFFFFFF)), (ray >> np.uint64(32)), np.full(ray.shape, bounce, np.uint64), np.full(ray.shape, purpose, np.uint64)]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & mask]
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF; k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def uniforms_for_rays(seed, ray_begin, n):
    """The (u1, u2) the oracle's direction recipe draws for rays [ray_begin, ray_begin+n) -- theta = 2 pi (w0 + 1/2) /
    2^32, cos(phi) = 2 u2 - 1 with u2 = ((w1 >> 8) + 1) 2^-24 -- narrowed to the float32 curand_uniform returns."""
    w = philox4x32_10(seed, np.arange(ray_begin, ray_begin + n, dtype=np.uint64))
    u1 = ((w[:, 0].astype(np.float64) + 0.5) * 2.0 ** -32).astype(np.float32)
    u2 = (((w[:, 1] >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24).astype(np.float32)
    return np.stack([u1, u2], axis=1)
