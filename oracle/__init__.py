"""CPU ORACLE -- test infrastructure, NOT the product.

Pin status: the tracer's per-ray arithmetic (raygen, closest-hit, miss: OR/devicePrograms.cu:62-254), loadOBJ and
placeReceiver are PINNED against the reference's own sources compiled where they lie (oracle/ref.py, oracle/_ref/,
tests/test_pin_cpu.py, tests/golden/ref_*).  PARITY UNPINNED for what the reference delegates to closed third-party
code absent from /root/reference: OptiX 7.7's ray-triangle search, cuRAND's clock64()-seeded stream, cuFFT (the
convolvers are anchored on the fp64 direct convolution instead).

ctypes binding of oracle/_build/liboracle.so (oracle_trace.cpp, oracle_conv.cpp;
build with `make -C oracle`).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


class Params(C.Structure):
    _fields_ = [
        ("size_x", C.c_int32), ("size_y", C.c_int32), ("size_z", C.c_int32),
        ("emitter", C.c_float * 3), ("sphere_center", C.c_float * 3),
        ("base_power", C.c_float), ("energy_thres", C.c_float), ("max_bounces", C.c_uint32),
        ("hrtf_absorption_rate", C.c_float), ("sample_rate", C.c_int32), ("is_mono", C.c_int32),
        ("ir_length", C.c_int32), ("bands", C.c_int32), ("seed", C.c_uint64),
    ]


def build(target=None):
    subprocess.check_call(["make", "-C", _HERE, "-s"] + ([target] if target else []))


def _bind(path):
    L = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    L.oracle_trace.restype = C.c_int64
    L.oracle_trace.argtypes = [C.POINTER(Params), fp, ip, C.c_int64, fp, fp, C.c_int32, C.c_int64,
                               C.c_int64, C.c_int32, C.c_int32, dp, ip, ip, fp, ip]
    L.oracle_scene_create.restype = C.c_void_p
    L.oracle_scene_create.argtypes = [fp, ip, C.c_int64, fp, fp, C.c_int32, C.c_int32, C.c_int32]
    L.oracle_scene_destroy.argtypes = [C.c_void_p]
    L.oracle_trace_scene.restype = C.c_int64
    L.oracle_trace_scene.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64, C.c_int64, C.c_int32, dp, ip, ip, fp, ip]
    L.oracle_shade_hit.argtypes = [fp, C.c_float, fp, C.c_float, C.c_float, fp, C.c_int32, C.c_float, C.c_int32, C.c_int32,
                                   fp, ip, ip, ip, ip, fp]
    L.oracle_ray_direction.argtypes = [C.c_uint64, C.c_uint64, fp]
    L.oracle_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    L.oracle_closest_hit.restype = C.c_int64
    L.oracle_closest_hit.argtypes = [fp, C.c_int64, fp, fp, fp, fp, fp]
    L.oracle_finalize_ir.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, fp, fp]
    L.oracle_direct_conv.argtypes = [fp, C.c_int64, fp, C.c_int64, dp, C.c_int32]
    L.oracle_direct_conv_window.argtypes = [fp, C.c_int64, fp, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int32]
    L.oracle_reference_file_conv.argtypes = [fp, C.c_int64, fp, C.c_int32, C.c_int32, dp, C.c_int32]
    L.oracle_reference_live_conv.argtypes = [dp, C.c_int64, fp, fp, C.c_int32, dp]
    L.oracle_upola.restype = C.c_double
    L.oracle_upola.argtypes = [fp, C.c_int64, C.c_int32, fp, fp, C.c_int32, fp, fp]
    return L


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = _bind(_LIB_PATH)
    return _lib


def load_variant(name):
    """Another build of the same sources (oracle/Makefile): "native" = -O3 -march=native, compiled on this machine;
    "o2" = -O2.  Returns the bound library without making it the default."""
    build(name)
    return _bind(os.path.join(_HERE, "_build", f"liboracle_{name}.so"))


def use_native():
    """Build the oracle with -O3 -march=native ON THIS MACHINE and make it the library every call below uses
    (bench.py's CPU-baseline legs).  Results are bit-identical to the portable build."""
    global _lib
    _lib = load_variant("native")
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def make_params(rays=(100, 100, 10), emitter=(0, 0, 0), sphere_center=(0, 0, 0), base_power=100.0,
                energy_thres=0.0, max_bounces=10, hrtf=0.9, sample_rate=16000, mono=False,
                ir_length=16000, bands=1, seed=1) -> Params:
    p = Params()
    p.size_x, p.size_y, p.size_z = (int(r) for r in rays)
    p.emitter[:] = [float(x) for x in emitter]
    p.sphere_center[:] = [float(x) for x in sphere_center]
    p.base_power = base_power
    p.energy_thres = energy_thres
    p.max_bounces = int(max_bounces)
    p.hrtf_absorption_rate = hrtf
    p.sample_rate = int(sample_rate)
    p.is_mono = 1 if mono else 0
    p.ir_length = int(ir_length)
    p.bands = int(bands)
    p.seed = int(seed)
    return p


def trace(p: Params, scene, ray_begin=0, n_rays=None, use_bvh=None, n_threads=None, records=True):
    """Run the oracle tracer over a scene.FlatScene.  Returns dict(hist, bin, ear, energy,
    nseg, segments)."""
    L = lib()
    n_total = p.size_x * p.size_y * p.size_z
    if n_rays is None:
        n_rays = n_total - ray_begin
    tv = np.ascontiguousarray(scene.tri_verts, dtype=np.float32)
    tm = np.ascontiguousarray(scene.tri_mat, dtype=np.int32)
    ab = np.ascontiguousarray(scene.absorption, dtype=np.float32)
    sc = np.ascontiguousarray(scene.scattering, dtype=np.float32)
    T = tv.shape[0]
    if use_bvh is None:
        use_bvh = T > 64
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    hist = np.zeros((2, p.bands, p.ir_length), np.float64)
    rb = re_ = rn = en = None
    if records:
        rb = np.empty(n_rays, np.int32)
        re_ = np.empty(n_rays, np.int32)
        rn = np.empty(n_rays, np.int32)
        en = np.empty((n_rays, p.bands), np.float32)
    segs = L.oracle_trace(C.byref(p), _fp(tv), _ip(tm), T, _fp(ab), _fp(sc), ab.shape[0], ray_begin, n_rays,
                          1 if use_bvh else 0, n_threads, _dp(hist), _ip(rb), _ip(re_), _fp(en), _ip(rn))
    if segs < 0:
        raise RuntimeError("oracle_trace failed")
    return dict(hist=hist, bin=rb, ear=re_, energy=en, nseg=rn, segments=int(segs))


class PreparedScene:
    """Scene + oracle BVH built once (bench.py's CPU baseline keeps the build untimed)."""

    def __init__(self, scene, bands=1, use_bvh=True):
        L = lib()
        self._tv = np.ascontiguousarray(scene.tri_verts, dtype=np.float32)
        tm = np.ascontiguousarray(scene.tri_mat, dtype=np.int32)
        ab = np.ascontiguousarray(scene.absorption, dtype=np.float32)
        sc = np.ascontiguousarray(scene.scattering, dtype=np.float32)
        self.bands = bands
        self._h = L.oracle_scene_create(_fp(self._tv), _ip(tm), self._tv.shape[0], _fp(ab), _fp(sc), ab.shape[0], bands,
                                        1 if use_bvh else 0)
        if not self._h:
            raise RuntimeError("oracle_scene_create failed")

    def trace(self, p: Params, ray_begin, n_rays, n_threads=None, want_hist=True):
        hist = np.zeros((2, p.bands, p.ir_length), np.float64) if want_hist else None
        segs = lib().oracle_trace_scene(self._h, C.byref(p), int(ray_begin), int(n_rays), n_threads or (os.cpu_count() or 1),
                                        _dp(hist), None, None, None, None)
        if segs < 0:
            raise RuntimeError("oracle_trace_scene failed")
        return int(segs), hist

    def trace_records(self, p: Params, ray_begin, n_rays, n_threads=None):
        """Same, with the per-ray records: dict(hist, bin, ear, energy, nseg, segments) like oracle.trace."""
        n_rays = int(n_rays)
        hist = np.zeros((2, p.bands, p.ir_length), np.float64)
        rb = np.empty(n_rays, np.int32); re_ = np.empty(n_rays, np.int32); rn = np.empty(n_rays, np.int32)
        en = np.empty((n_rays, p.bands), np.float32)
        segs = lib().oracle_trace_scene(self._h, C.byref(p), int(ray_begin), n_rays, n_threads or (os.cpu_count() or 1),
                                        _dp(hist), _ip(rb), _ip(re_), _fp(en), _ip(rn))
        if segs < 0:
            raise RuntimeError("oracle_trace_scene failed")
        return dict(hist=hist, bin=rb, ear=re_, energy=en, nseg=rn, segments=int(segs))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.oracle_scene_destroy(self._h)
            self._h = None


def finalize_ir(hist, mono=False):
    L = lib()
    _, bands, n = hist.shape
    l = np.empty((bands, n), np.float32)
    r = np.empty((bands, n), np.float32)
    h = np.ascontiguousarray(hist, dtype=np.float64)
    L.oracle_finalize_ir(_dp(h), bands, n, 1 if mono else 0, _fp(l), _fp(r))
    return l, r


def ray_direction(seed, ray_id):
    d = np.empty(3, np.float32)
    lib().oracle_ray_direction(int(seed), int(ray_id), _fp(d))
    return d


def philox(seed, ray_id, bounce=0, purpose=0):
    o = np.empty(4, np.uint32)
    lib().oracle_philox(int(seed), int(ray_id), int(bounce), int(purpose), o.ctypes.data_as(C.POINTER(C.c_uint32)))
    return o


def closest_hit(tri_verts, org, direction):
    tv = np.ascontiguousarray(tri_verts, dtype=np.float32)
    o = np.ascontiguousarray(org, dtype=np.float32)
    d = np.ascontiguousarray(direction, dtype=np.float32)
    t = C.c_float(); u = C.c_float(); v = C.c_float()
    i = lib().oracle_closest_hit(_fp(tv), tv.shape[0], _fp(o), _fp(d), C.byref(t), C.byref(u), C.byref(v))
    return int(i), t.value, u.value, v.value


def direct_conv(x, h, n_threads=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    h = np.ascontiguousarray(h, dtype=np.float32)
    y = np.zeros(len(x) + len(h) - 1, np.float64)
    lib().oracle_direct_conv(_fp(x), len(x), _fp(h), len(h), _dp(y), n_threads or (os.cpu_count() or 1))
    return y


def direct_conv_window(x, h, begin, count, n_threads=None):
    """Outputs [begin, begin+count) of direct_conv(x, h)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    h = np.ascontiguousarray(h, dtype=np.float32)
    assert 0 <= begin and count > 0 and begin + count <= len(x) + len(h) - 1
    y = np.zeros(count, np.float64)
    lib().oracle_direct_conv_window(_fp(x), len(x), _fp(h), len(h), int(begin), int(count), _dp(y), n_threads or (os.cpu_count() or 1))
    return y


def reference_file_conv(x, h, sample_rate, n_threads=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    h = np.ascontiguousarray(h, dtype=np.float32)
    y = np.zeros(len(x), np.float64)
    lib().oracle_reference_file_conv(_fp(x), len(x), _fp(h), len(h), int(sample_rate), _dp(y),
                                     n_threads or (os.cpu_count() or 1))
    return y


def reference_live_conv(x, ir_left, ir_right):
    x = np.ascontiguousarray(x, dtype=np.float64)
    l = np.ascontiguousarray(ir_left, dtype=np.float32)
    r = np.ascontiguousarray(ir_right, dtype=np.float32)
    out = np.zeros(2 * len(l), np.float64)
    lib().oracle_reference_live_conv(_dp(x), len(x), _fp(l), _fp(r), len(l), _dp(out))
    return out


def upola(x, h_left, h_right, block=512):
    x = np.ascontiguousarray(x, dtype=np.float32)
    nb = len(x) // block
    hl = np.ascontiguousarray(h_left, dtype=np.float32)
    hr = np.ascontiguousarray(h_right, dtype=np.float32)
    ol = np.zeros(nb * block, np.float32)
    or_ = np.zeros(nb * block, np.float32)
    secs = lib().oracle_upola(_fp(x), nb, block, _fp(hl), _fp(hr), len(hl), _fp(ol), _fp(or_))
    return ol, or_, secs
