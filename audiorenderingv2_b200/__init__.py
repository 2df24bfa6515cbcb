"""audiorenderingv2_b200 -- ctypes binding of libarv2.so (include/arv2.h).

The product is the C-ABI shared library (hand-written sm_100a CUDA + C++ host code in
``csrc/``); the C++ mirror of the reference's ``AudioRenderer`` class lives in
``csrc/host/audio_renderer.hpp``.  This module is the thin Python face of the same ABI
used by ``tests/`` and ``bench.py``; names follow the reference
(prebuild/obj_raytracer/AudioRenderer.h:16-152, OptixModel.h, Context.cpp).

There is no CPU fallback: the library must be built (``__graft_entry__.build()``) and
anything that touches the GPU raises ``Arv2Error`` when no B200 / driver is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARV2_LIB") or os.path.join(_HERE, "lib", "libarv2.so")   # ARV2_LIB: A/B builds while tuning
MAX_BANDS = 8
CONV_LINEAR, CONV_REFERENCE = 0, 1


class Arv2Error(RuntimeError):
    pass


class Material(C.Structure):
    _fields_ = [("name", C.c_char_p), ("mat_absorption", C.c_float * MAX_BANDS), ("scattering", C.c_float)]


class RendererDesc(C.Structure):
    _fields_ = [
        ("ir_length_in_seconds", C.c_uint32), ("sample_rate", C.c_int32),
        ("rays_x", C.c_int32), ("rays_y", C.c_int32), ("rays_z", C.c_int32),
        ("bands", C.c_int32), ("device", C.c_int32), ("record_rays", C.c_int32),
        ("path_cache", C.c_int32), ("bvh_builder", C.c_int32),
        ("materials", C.POINTER(Material)), ("n_materials", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("initial_volume", C.c_float), ("ir_length_in_seconds", C.c_uint32),
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("write_first_ir_to_file", C.c_int32), ("write_first_output_to_file", C.c_int32),
        ("re_render_distance_threshold", C.c_float), ("re_render_angle_threshold", C.c_float),
        ("mono", C.c_int32),
        ("scene_file_path", C.c_char * 512), ("audio_file_path", C.c_char * 512),
        ("materials_file_path", C.c_char * 512),
        ("initial_receiver_pos", C.c_float * 3), ("initial_emitter_pos", C.c_float * 3),
        ("base_power", C.c_float), ("rays", C.c_float * 3), ("ray_energy_threshold", C.c_float),
        ("ray_max_bounces", C.c_uint32), ("hrtf_absorption_rate", C.c_float),
        ("n_materials", C.c_int32), ("material_names", (C.c_char * 64) * 64),
        ("material_absorption", C.c_float * 64),
        ("seed", C.c_uint64), ("bands", C.c_int32),
    ]


_lib = None
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

# every symbol include/arv2.h declares: (restype, argtypes)
SYMBOLS = {
    "arv2_last_error": (C.c_char_p, []),
    "arv2_version": (C.c_char_p, []),
    "arv2_scene_load_obj": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "arv2_scene_from_triangles": (C.c_int, [_fp, _ip, C.c_int64, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(_vp)]),
    "arv2_scene_counts": (C.c_int, [_vp, C.POINTER(C.c_int64), _ip]),
    "arv2_scene_get_triangles": (C.c_int, [_vp, _fp, _ip]),
    "arv2_scene_mesh_material": (C.c_char_p, [_vp, C.c_int32]),
    "arv2_scene_bounds": (C.c_int, [_vp, _fp, _fp]),
    "arv2_scene_bvh_stats": (C.c_int, [_vp, _vp]),
    "arv2_scene_destroy": (None, [_vp]),
    "arv2_receiver_load": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_vp)]),
    "arv2_receiver_from_triangles": (C.c_int, [_fp, C.c_int64, _fp, C.c_int64, C.POINTER(_vp)]),
    "arv2_receiver_counts": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "arv2_receiver_place": (C.c_int, [_vp, _fp, C.c_float, _fp, _fp]),
    "arv2_receiver_destroy": (None, [_vp]),
    "arv2_material_absorption": (C.c_float, [C.c_char_p, C.POINTER(Material), C.c_int32]),
    "arv2_config_parse": (C.c_int, [C.c_char_p, C.POINTER(Config)]),
    "arv2_config_load": (C.c_int, [C.c_char_p, C.POINTER(Config)]),
    "arv2_create": (C.c_int, [_vp, _vp, C.POINTER(RendererDesc), C.POINTER(_vp)]),
    "arv2_destroy": (None, [_vp]),
    "arv2_set_emitter": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float]),
    "arv2_set_receiver": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float, C.c_float]),
    "arv2_set_thresholds": (C.c_int, [_vp, C.c_float, C.c_uint32]),
    "arv2_set_base_power": (C.c_int, [_vp, C.c_float]),
    "arv2_set_hrtf_absorption_rate": (C.c_int, [_vp, C.c_float]),
    "arv2_set_mono": (C.c_int, [_vp, C.c_int32]),
    "arv2_set_seed": (C.c_int, [_vp, C.c_uint64]),
    "arv2_set_coherent_order": (C.c_int, [_vp, C.c_int32]),
    "arv2_set_sweep_min_rays": (C.c_int, [_vp, C.c_int64]),
    "arv2_set_stream": (C.c_int, [_vp, _vp]),
    "arv2_render": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "arv2_render_range": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_double)]),
    "arv2_finalize": (C.c_int, [_vp]),
    "arv2_rerender": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "arv2_comm_unique_id": (C.c_int, [_vp]),
    "arv2_comm_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _vp, C.POINTER(_vp)]),
    "arv2_comm_info": (C.c_int, [_vp, _ip, _ip, _ip]),
    "arv2_comm_destroy": (None, [_vp]),
    "arv2_comm_reduce_f32": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int32, _vp]),
    "arv2_shard_range": (None, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "arv2_render_sharded": (C.c_int, [_vp, _vp, C.POINTER(C.c_double)]),
    "arv2_render_tiles": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "arv2_set_shard_mode": (C.c_int, [_vp, C.c_int32]),
    "arv2_multi_create": (C.c_int, [_vp, _vp, C.POINTER(RendererDesc), _ip, C.c_int32, C.POINTER(_vp)]),
    "arv2_multi_size": (C.c_int32, [_vp]),
    "arv2_multi_ctx": (_vp, [_vp, C.c_int32]),
    "arv2_multi_render": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "arv2_multi_destroy": (None, [_vp]),
    "arv2_ir_length": (C.c_int, [_vp, _ip, _ip]),
    "arv2_get_ir": (C.c_int, [_vp, _fp, _fp]),
    "arv2_set_ir": (C.c_int, [_vp, _fp, _fp]),
    "arv2_ir_device": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "arv2_hist_device": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "arv2_path_cache_info": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "arv2_last_counters": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.c_int32]),
    "arv2_last_segments": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "arv2_last_upload_bytes": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "arv2_get_records": (C.c_int, [_vp, C.c_int64, _ip, _ip, _fp, _ip, C.POINTER(C.c_int64)]),
    "arv2_write_ir_text": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "arv2_write_convolved_text": (C.c_int, [C.c_char_p, C.c_char_p, _fp, _fp, C.c_size_t]),
    "arv2_convolve_file": (C.c_int, [_vp, _fp, C.c_size_t, _fp, _fp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "arv2_stream_open": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "arv2_stream_set_ir": (C.c_int, [_vp, C.c_int32, _fp, _fp]),
    "arv2_stream_set_ir_device": (C.c_int, [_vp, C.c_int32, _vp, _vp]),
    "arv2_stream_process": (C.c_int, [_vp, _fp, _fp]),
    "arv2_stream_process_device": (C.c_int, [_vp, _vp, _vp, _vp]),
    "arv2_stream_process_device_blocks": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp]),
    "arv2_stream_process_blocks": (C.c_int, [_vp, _fp, _fp, _fp, C.c_int32]),
    "arv2_stream_mix_device": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp]),
    "arv2_stream_set_gains": (C.c_int, [_vp, _fp]),
    "arv2_stream_reset": (C.c_int, [_vp]),
    "arv2_stream_close": (None, [_vp]),
    "arv2_global_angle": (C.c_float, [C.c_float, C.c_float]),
    "arv2_policy_create": (C.c_int, [C.c_float, C.c_float, _fp, C.c_float, C.POINTER(_vp)]),
    "arv2_policy_update": (C.c_int, [_vp, _fp, C.c_float, C.c_double, C.c_int32]),
    "arv2_policy_destroy": (None, [_vp]),
    "arv2_playback_fill": (C.c_int64, [C.POINTER(C.c_double), C.c_uint32, C.c_double, C.c_int32, _fp, _fp, C.c_size_t, C.c_size_t, C.c_float]),
    "arv2_ring_create": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "arv2_ring_add": (C.c_int, [_vp, C.POINTER(C.c_double), C.c_size_t]),
    "arv2_ring_get_and_reset": (C.c_int, [_vp, C.POINTER(C.c_double), C.c_size_t]),
    "arv2_ring_destroy": (None, [_vp]),
    "arv2_live_callback": (C.c_int, [_vp, C.POINTER(C.c_double), C.c_size_t, _vp]),
    "arv2_wav_read": (C.c_int, [C.c_char_p, C.POINTER(_fp), C.POINTER(C.c_size_t), _ip, _ip]),
    "arv2_wav_write_stereo_normalized": (C.c_int, [C.c_char_p, _fp, _fp, C.c_size_t, C.c_int32]),
    "arv2_free": (None, [_vp]),
}


def lib():
    """Load libarv2.so (no fallback: a missing build is an error)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Arv2Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise Arv2Error(f"arv2 error {rc}: {lib().arv2_last_error().decode(errors='replace')}")


def _f(a):
    return a.ctypes.data_as(_fp)


def _i(a):
    return a.ctypes.data_as(_ip)


class BvhStats(C.Structure):
    """struct arv2_bvh_stats (include/arv2.h)."""
    _fields_ = [("n_tris", C.c_int64), ("n_nodes", C.c_int64), ("n_leaves", C.c_int64), ("max_leaf_tris", C.c_int32),
                ("depth", C.c_int32), ("valid", C.c_int32), ("sah_nodes", C.c_double), ("sah_tris", C.c_double)]


class Scene:
    """struct OptixModel (OR/OptixModel.h:21-32)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load_obj(cls, path):
        """loadOBJ (OR/OptixModel.cpp:75-151)."""
        h = _vp()
        _check(lib().arv2_scene_load_obj(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def from_triangles(cls, tri_verts, tri_mesh, material_names):
        tv = np.ascontiguousarray(tri_verts, dtype=np.float32).reshape(-1, 9)
        tm = np.ascontiguousarray(tri_mesh, dtype=np.int32)
        names = (C.c_char_p * max(1, len(material_names)))(*[n.encode() for n in material_names])
        h = _vp()
        _check(lib().arv2_scene_from_triangles(_f(tv), _i(tm), tv.shape[0], names, len(material_names), C.byref(h)))
        return cls(h)

    def counts(self):
        n = C.c_int64(); m = C.c_int32()
        _check(lib().arv2_scene_counts(self._h, C.byref(n), C.byref(m)))
        return n.value, m.value

    def triangles(self):
        n, _ = self.counts()
        tv = np.empty((n, 3, 3), np.float32)
        tm = np.empty(n, np.int32)
        _check(lib().arv2_scene_get_triangles(self._h, _f(tv), _i(tm)))
        return tv, tm

    def mesh_materials(self):
        _, m = self.counts()
        return [lib().arv2_scene_mesh_material(self._h, i).decode() for i in range(m)]

    def bounds(self):
        lo = np.empty(3, np.float32); hi = np.empty(3, np.float32)
        _check(lib().arv2_scene_bounds(self._h, _f(lo), _f(hi)))
        return lo, hi

    def bvh_stats(self):
        """Builds (on the host) the BVH the renderer would upload and checks it; dict of arv2_bvh_stats."""
        st = BvhStats()
        _check(lib().arv2_scene_bvh_stats(self._h, C.byref(st)))
        return {n: getattr(st, n) for n, _ in BvhStats._fields_}

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.arv2_scene_destroy(self._h)
            self._h = None


class Receiver:
    """class Sphere / HalfSphere (OR/Sphere.cpp, OR/HalfSphere.cpp)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load(cls, left_obj, right_obj):
        h = _vp()
        _check(lib().arv2_receiver_load(os.fsencode(left_obj), os.fsencode(right_obj), C.byref(h)))
        return cls(h)

    @classmethod
    def from_triangles(cls, left, right):
        l = np.ascontiguousarray(left, dtype=np.float32).reshape(-1, 9)
        r = np.ascontiguousarray(right, dtype=np.float32).reshape(-1, 9)
        h = _vp()
        _check(lib().arv2_receiver_from_triangles(_f(l), l.shape[0], _f(r), r.shape[0], C.byref(h)))
        return cls(h)

    def counts(self):
        a = C.c_int64(); b = C.c_int64()
        _check(lib().arv2_receiver_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def place(self, cam, rotation_deg):
        """placeReceiver (OR/OptixModel.cpp:153-257), host only."""
        nl, nr = self.counts()
        l = np.empty((nl, 3, 3), np.float32); r = np.empty((nr, 3, 3), np.float32)
        c = np.ascontiguousarray(cam, dtype=np.float32)
        _check(lib().arv2_receiver_place(self._h, _f(c), float(rotation_deg), _f(l), _f(r)))
        return l, r

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.arv2_receiver_destroy(self._h)
            self._h = None


def material_absorption(name, materials=()):
    """getMaterialAbsorption (OR/AudioRenderer.cpp:34-56); materials = [(name, absorption)]."""
    arr = (Material * max(1, len(materials)))()
    keep = []
    for i, (n, a) in enumerate(materials):
        b = n.encode(); keep.append(b)
        arr[i].name = b
        arr[i].mat_absorption[0] = a
    return lib().arv2_material_absorption(name.encode(), arr, len(materials))


def parse_config(text: str) -> Config:
    """Context::loadContext (OR/Context.cpp:15-165)."""
    cfg = Config()
    _check(lib().arv2_config_parse(text.encode(), C.byref(cfg)))
    return cfg


def load_config(path) -> Config:
    cfg = Config()
    _check(lib().arv2_config_load(os.fsencode(path), C.byref(cfg)))
    return cfg


def config_materials(cfg: Config):
    return [(cfg.material_names[i].value.decode(), cfg.material_absorption[i]) for i in range(cfg.n_materials)]


COMM_ID_BYTES = 128


def _make_desc(ir_length_in_seconds, sample_rate, materials, rays_per_dimension, bands, device, record_rays, path_cache, bvh_builder):
    """arv2_renderer_desc + the byte strings / array it points into (keep them alive while it is used)."""
    d = RendererDesc()
    d.ir_length_in_seconds = int(ir_length_in_seconds)
    d.sample_rate = int(sample_rate)
    d.rays_x, d.rays_y, d.rays_z = (int(v) for v in rays_per_dimension)
    d.bands = bands; d.device = device
    d.record_rays = 1 if record_rays else 0
    d.path_cache = 1 if path_cache else 0
    d.bvh_builder = bvh_builder
    arr = (Material * max(1, len(materials)))()
    keep = [arr]
    for i, m in enumerate(materials):
        name, a = m[0], m[1]
        b = name.encode(); keep.append(b)
        arr[i].name = b
        vals = list(a) if isinstance(a, (list, tuple, np.ndarray)) else [a] * MAX_BANDS
        for k in range(MAX_BANDS):
            arr[i].mat_absorption[k] = float(vals[min(k, len(vals) - 1)])
        arr[i].scattering = float(m[2]) if len(m) > 2 else 0.0
    d.materials = arr; d.n_materials = len(materials)
    return d, keep


def shard_range(n_rays, rank, n_ranks):
    """The contiguous slice (begin, count) of the seeded ray set rank `rank` of `n_ranks` traces."""
    b = C.c_int64(); c = C.c_int64()
    lib().arv2_shard_range(int(n_rays), int(rank), int(n_ranks), C.byref(b), C.byref(c))
    return b.value, c.value


def _prefer_bundled_nccl():
    """libarv2 binds NCCL with dlopen: ARV2_NCCL_LIB, else a libnccl.so.2 already in the process, else the system's.  In a
    Python process that imports torch LATER, the system's (older) library would then stand in for the one torch's
    libtorch_cuda.so was linked against (same soname) and break `import torch`; so when the pip-installed NCCL that torch
    ships with is present and nothing was chosen, it is named explicitly.  A C++ host (arv2_cli) is not affected."""
    if os.environ.get("ARV2_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            p = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(p):
                os.environ["ARV2_NCCL_LIB"] = p
                return
    except (ImportError, ValueError, AttributeError):
        pass


class Comm:
    """One NCCL rank inside libarv2 (arv2_comm_*).  `Comm.unique_id()` on one rank, the 128 bytes handed to the
    others by any transport (torch.distributed broadcast, a file, MPI ...), then `Comm(device, rank, n, id)` on each."""

    def __init__(self, device, rank, n_ranks, unique_id: bytes):
        assert len(unique_id) == COMM_ID_BYTES
        _prefer_bundled_nccl()
        self._h = _vp()
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES)
        _check(lib().arv2_comm_create(int(device), int(rank), int(n_ranks), C.cast(buf, _vp), C.byref(self._h)))
        self.rank, self.n_ranks = rank, n_ranks

    @staticmethod
    def unique_id() -> bytes:
        _prefer_bundled_nccl()
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().arv2_comm_unique_id(C.cast(buf, _vp)))
        return buf.raw

    def reduce_f32(self, d_buf, count, root=0, cuda_stream=None):
        """ncclReduce(sum) of `count` device floats onto `root`, in place."""
        _check(lib().arv2_comm_reduce_f32(self._h, d_buf, int(count), int(root), cuda_stream))

    def nccl_version(self):
        v = C.c_int32()
        _check(lib().arv2_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.arv2_comm_destroy(self._h)
            self._h = None

    __del__ = close


class AudioRenderer:
    """Mirror of class AudioRenderer (OR/AudioRenderer.h:16-152) over the C ABI."""

    def __init__(self, model: Scene, ir_length_in_seconds, sample_rate, materials, rays_per_dimension,
                 receiver: Receiver | None = None, bands=1, device=0, record_rays=False, path_cache=False,
                 bvh_builder=0, _borrowed=None):
        """materials: [(name, absorption | [absorption per band], scattering=0)]."""
        self.bands = bands
        self.n_rays = int(rays_per_dimension[0]) * int(rays_per_dimension[1]) * int(rays_per_dimension[2])
        self.sample_rate = int(sample_rate)
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._h = _vp(_borrowed)                  # a context owned by a MultiRenderer
        else:
            d, self._keep = _make_desc(ir_length_in_seconds, sample_rate, materials, rays_per_dimension, bands, device,
                                       record_rays, path_cache, bvh_builder)
            self._h = _vp()
            _check(lib().arv2_create(model._h, receiver._h if receiver is not None else None, C.byref(d), C.byref(self._h)))
        n = C.c_int32(); b = C.c_int32()
        _check(lib().arv2_ir_length(self._h, C.byref(n), C.byref(b)))
        self.ir_length = n.value

    # -- setters (reference names) ------------------------------------------------
    def setEmitterPosInOptix(self, pos):
        _check(lib().arv2_set_emitter(self._h, *[float(v) for v in pos]))

    def setSphereCenterInOptix(self, pos, yaw_deg=0.0):
        """placeReceiver + setSphereCenterInOptix."""
        _check(lib().arv2_set_receiver(self._h, float(pos[0]), float(pos[1]), float(pos[2]), float(yaw_deg)))

    def setThresholds(self, energy, max_bounces):
        _check(lib().arv2_set_thresholds(self._h, float(energy), int(max_bounces)))

    def setBasePower(self, p):
        _check(lib().arv2_set_base_power(self._h, float(p)))

    def set_hrtf_absorption_rate(self, r):
        _check(lib().arv2_set_hrtf_absorption_rate(self._h, float(r)))

    def setMonoOutput(self, v):
        _check(lib().arv2_set_mono(self._h, 1 if v else 0))

    def set_seed(self, seed):
        _check(lib().arv2_set_seed(self._h, int(seed)))

    def set_coherent_order(self, on):
        _check(lib().arv2_set_coherent_order(self._h, 1 if on else 0))

    def set_sweep_min_rays(self, n):
        """Launches of >= n rays use the bounce-synchronous sweep_kernel (0: never); scheduling only."""
        _check(lib().arv2_set_sweep_min_rays(self._h, int(n)))

    def set_stream(self, cuda_stream_ptr):
        _check(lib().arv2_set_stream(self._h, cuda_stream_ptr))

    # -- rendering ------------------------------------------------------------------
    def render(self):
        """AudioRenderer::render; returns device milliseconds of the trace."""
        ms = C.c_double()
        _check(lib().arv2_render(self._h, C.byref(ms)))
        return ms.value

    def render_range(self, ray_begin, n_rays, zero_first=True):
        ms = C.c_double()
        _check(lib().arv2_render_range(self._h, int(ray_begin), int(n_rays), 1 if zero_first else 0, C.byref(ms)))
        return ms.value

    def render_tiles(self, rank, n_ranks, zero_first=True):
        """One rank's direction tiles of the seeded set into the histogram (no exchange, no finalise)."""
        ms = C.c_double()
        _check(lib().arv2_render_tiles(self._h, int(rank), int(n_ranks), 1 if zero_first else 0, C.byref(ms)))
        return ms.value

    def set_shard_mode(self, mode):
        """arv2_render_sharded: 1 = direction tiles (default), 0 = contiguous slices of ray ids."""
        _check(lib().arv2_set_shard_mode(self._h, int(mode)))

    def finalize(self):
        _check(lib().arv2_finalize(self._h))

    def render_sharded(self, comm: "Comm"):
        """This rank's slice + NCCL all-reduce of the histogram + finalise, inside the library (collective)."""
        ms = C.c_double()
        _check(lib().arv2_render_sharded(self._h, comm._h, C.byref(ms)))
        return ms.value

    def rerender(self):
        ms = C.c_double()
        _check(lib().arv2_rerender(self._h, C.byref(ms)))
        return ms.value

    def get_ir(self):
        l = np.empty((self.bands, self.ir_length), np.float32)
        r = np.empty((self.bands, self.ir_length), np.float32)
        _check(lib().arv2_get_ir(self._h, _f(l), _f(r)))
        return l, r

    def set_ir(self, left, right):
        l = np.ascontiguousarray(left, dtype=np.float32); r = np.ascontiguousarray(right, dtype=np.float32)
        assert l.size == self.ir_length and r.size == self.ir_length
        _check(lib().arv2_set_ir(self._h, _f(l), _f(r)))

    def ir_device(self):
        a = _vp(); b = _vp()
        _check(lib().arv2_ir_device(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def hist_device(self):
        p = _vp(); n = C.c_int64()
        _check(lib().arv2_hist_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def last_segments(self):
        s = C.c_int64()
        _check(lib().arv2_last_segments(self._h, C.byref(s)))
        return s.value

    def last_counters(self, n=24):
        """Launch counters of the last render (arv2_last_counters; traversal tallies need the stats build)."""
        out = (C.c_uint64 * n)()
        _check(lib().arv2_last_counters(self._h, out, n))
        return [int(v) for v in out]

    def path_cache_info(self):
        """(cached segments, device bytes) of the receiver-independent path cache."""
        a = C.c_int64(); b = C.c_int64()
        _check(lib().arv2_path_cache_info(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_upload_bytes(self):
        b = C.c_int64()
        _check(lib().arv2_last_upload_bytes(self._h, C.byref(b)))
        return b.value

    def records(self, n_rays=None):
        """Per-ray records of the range last traced (sized by the library, not by the caller)."""
        last = C.c_int64()
        _check(lib().arv2_get_records(self._h, 0, None, None, None, None, C.byref(last)))
        n = last.value if n_rays is None else min(int(n_rays), last.value)
        b = np.empty(n, np.int32); e = np.empty(n, np.int32); s = np.empty(n, np.int32)
        en = np.empty((n, self.bands), np.float32)
        _check(lib().arv2_get_records(self._h, n, _i(b), _i(e), _f(en), _i(s), None))
        return dict(bin=b, ear=e, energy=en, nseg=s)

    def write_ir_text(self, left_path, right_path):
        _check(lib().arv2_write_ir_text(self._h, os.fsencode(left_path), os.fsencode(right_path)))

    # -- convolution ----------------------------------------------------------------
    def convoluteAudioFile(self, samples, mode=CONV_LINEAR):
        """AudioRenderer::convoluteAudioFile; returns (left, right, conv_ms, process_ms)."""
        x = np.ascontiguousarray(samples, dtype=np.float32)
        yl = np.empty_like(x); yr = np.empty_like(x)
        t = C.c_double(); tp = C.c_double()
        _check(lib().arv2_convolve_file(self._h, _f(x), x.size, _f(yl), _f(yr), mode, C.byref(t), C.byref(tp)))
        return yl, yr, t.value, tp.value

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            if self._owned:
                _lib.arv2_destroy(self._h)
            self._h = None

    __del__ = close


class MultiRenderer:
    """One process, several GPUs (arv2_multi_*): one context and one NCCL rank per device, a render runs
    arv2_render_sharded on one host thread per device.  `self.renderers[i]` is the context of device i
    (set parameters on all of them with `each`); every device ends up with the full IR."""

    def __init__(self, model: Scene, ir_length_in_seconds, sample_rate, materials, rays_per_dimension, devices,
                 receiver: Receiver | None = None, bands=1, record_rays=False, bvh_builder=0):
        d, self._keep = _make_desc(ir_length_in_seconds, sample_rate, materials, rays_per_dimension, bands, 0, record_rays, False, bvh_builder)
        dev = (C.c_int32 * len(devices))(*[int(x) for x in devices])
        _prefer_bundled_nccl()
        self._h = _vp()
        _check(lib().arv2_multi_create(model._h, receiver._h if receiver is not None else None, C.byref(d), dev, len(devices), C.byref(self._h)))
        self.renderers = [AudioRenderer(model, ir_length_in_seconds, sample_rate, materials, rays_per_dimension, bands=bands,
                                        _borrowed=lib().arv2_multi_ctx(self._h, i)) for i in range(lib().arv2_multi_size(self._h))]

    def each(self, fn):
        for r in self.renderers:
            fn(r)

    def render(self):
        ms = C.c_double()
        _check(lib().arv2_multi_render(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            for r in self.renderers:
                r._h = None
            _lib.arv2_multi_destroy(self._h)
            self._h = None

    __del__ = close


class ConvStream:
    """Streaming partitioned convolver (replaces AudioRenderer::convoluteLiveInput)."""

    def __init__(self, n_sources, block, ir_length, device=0):
        self._h = _vp()
        self.n_sources, self.block, self.ir_length = n_sources, block, ir_length
        _check(lib().arv2_stream_open(device, n_sources, block, ir_length, C.byref(self._h)))

    def set_ir(self, source, left, right):
        l = np.ascontiguousarray(left, dtype=np.float32); r = np.ascontiguousarray(right, dtype=np.float32)
        assert l.size == self.ir_length and r.size == self.ir_length
        _check(lib().arv2_stream_set_ir(self._h, source, _f(l), _f(r)))

    def set_ir_device(self, source, d_left, d_right):
        _check(lib().arv2_stream_set_ir_device(self._h, source, d_left, d_right))

    def process(self, block_in):
        x = np.ascontiguousarray(block_in, dtype=np.float32).reshape(self.n_sources, self.block)
        out = np.empty((self.n_sources, 2, self.block), np.float32)
        _check(lib().arv2_stream_process(self._h, _f(x), _f(out)))
        return out

    def process_device(self, d_in, d_out, cuda_stream=None):
        _check(lib().arv2_stream_process_device(self._h, d_in, d_out, cuda_stream))

    def process_device_blocks(self, d_in, d_out, n_blocks, cuda_stream=None):
        """n_blocks consecutive blocks: d_in [n_blocks][n_sources][block], d_out [n_blocks][n_sources][2][block]."""
        _check(lib().arv2_stream_process_device_blocks(self._h, d_in, d_out, n_blocks, cuda_stream))

    def process_blocks(self, blocks_in, want_out=True, want_mix=False):
        """Up to 16 blocks, host buffers: blocks_in [n_blocks][n_sources][block] -> (out [n_blocks][n_sources][2][block] | None,
        mix [n_blocks][2][block] | None)."""
        x = np.ascontiguousarray(blocks_in, dtype=np.float32).reshape(-1, self.n_sources, self.block)
        nb = x.shape[0]
        out = np.empty((nb, self.n_sources, 2, self.block), np.float32) if want_out else None
        mix = np.empty((nb, 2, self.block), np.float32) if want_mix else None
        _check(lib().arv2_stream_process_blocks(self._h, _f(x), _f(out) if want_out else None, _f(mix) if want_mix else None, nb))
        return out, mix

    def mix_device(self, d_out, d_mix, n_blocks, cuda_stream=None):
        _check(lib().arv2_stream_mix_device(self._h, d_out, d_mix, n_blocks, cuda_stream))

    def set_gains(self, gains):
        g = None if gains is None else np.ascontiguousarray(gains, dtype=np.float32)
        _check(lib().arv2_stream_set_gains(self._h, _f(g) if g is not None else None))

    def reset(self):
        _check(lib().arv2_stream_reset(self._h))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.arv2_stream_close(self._h)
            self._h = None

    __del__ = close


def global_angle(orientation_x, orientation_z):
    """Camera::calculate_global_angle (OR/Camera.cpp:31-41)."""
    return lib().arv2_global_angle(float(orientation_x), float(orientation_z))


class RerenderPolicy:
    """Re-render trigger of the GL loop (OR/main.cpp:470-498)."""

    def __init__(self, distance_threshold, angle_threshold_deg, start_pos, start_angle_deg):
        self._h = _vp()
        p = np.ascontiguousarray(start_pos, dtype=np.float32)
        _check(lib().arv2_policy_create(float(distance_threshold), float(angle_threshold_deg), _f(p), float(start_angle_deg), C.byref(self._h)))

    def update(self, pos, angle_deg, now_s, is_rendering=False):
        p = np.ascontiguousarray(pos, dtype=np.float32)
        return bool(lib().arv2_policy_update(self._h, _f(p), float(angle_deg), float(now_s), 1 if is_rendering else 0))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.arv2_policy_destroy(self._h)
            self._h = None


def playback_fill(n_frames, stream_time, sample_rate, left, right, output_buffer_len, volume=1.0):
    """audioHandler (OR/main.cpp:69-97): one interleaved RTAUDIO_FLOAT64 callback buffer."""
    l = np.ascontiguousarray(left, dtype=np.float32); r = np.ascontiguousarray(right, dtype=np.float32)
    out = np.zeros(2 * n_frames, np.float64)
    n = lib().arv2_playback_fill(out.ctypes.data_as(C.POINTER(C.c_double)), int(n_frames), float(stream_time), int(sample_rate),
                                 _f(l), _f(r), l.size, int(output_buffer_len), float(volume))
    return out, int(n)


class Ring:
    """CircularBuffer<double> (OR/CircularBuffer.h)."""

    def __init__(self, size):
        self._h = _vp()
        _check(lib().arv2_ring_create(int(size), C.byref(self._h)))

    def add(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(lib().arv2_ring_add(self._h, v.ctypes.data_as(C.POINTER(C.c_double)), v.size))

    def get_and_reset(self, n):
        out = np.empty(n, np.float64)
        _check(lib().arv2_ring_get_and_reset(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), int(n)))
        return out

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.arv2_ring_destroy(self._h)
            self._h = None


def live_callback(stream: "ConvStream", samples, ring: Ring):
    """audioHandlerWithMic + convoluteLiveInput for one block of mic samples."""
    x = np.ascontiguousarray(samples, dtype=np.float64)
    _check(lib().arv2_live_callback(stream._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.size, ring._h))


def write_convolved_text(left_path, right_path, left, right):
    """output_convolute_left.txt / output_convolute_right.txt of the reference (OR/AudioRenderer.cpp:720-744)."""
    l = np.ascontiguousarray(left, dtype=np.float32); r = np.ascontiguousarray(right, dtype=np.float32)
    _check(lib().arv2_write_convolved_text(os.fsencode(left_path), os.fsencode(right_path), _f(l), _f(r), l.size))


def wav_read(path):
    p = _fp(); n = C.c_size_t(); sr = C.c_int32(); ch = C.c_int32()
    _check(lib().arv2_wav_read(os.fsencode(path), C.byref(p), C.byref(n), C.byref(sr), C.byref(ch)))
    try:
        a = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
    finally:
        lib().arv2_free(p)
    return sr.value, ch.value, a


def wav_write_stereo_normalized(path, left, right, sample_rate):
    l = np.ascontiguousarray(left, dtype=np.float32); r = np.ascontiguousarray(right, dtype=np.float32)
    _check(lib().arv2_wav_write_stereo_normalized(os.fsencode(path), _f(l), _f(r), l.size, int(sample_rate)))
