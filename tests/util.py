"""Shared helpers for the parity tests: build the same scene for the CUDA path
(audiorenderingv2_b200, through the C ABI) and for the CPU oracle."""
import numpy as np

import audiorenderingv2_b200 as arv
import oracle
from oracle import scene as osc


def oracle_model(tri_verts, tri_mesh, names):
    return osc.Model(meshes=[osc.Mesh(str(names[i]), np.ascontiguousarray(tri_verts[tri_mesh == i])) for i in range(len(names))])


class Case:
    """One render configuration, runnable on both sides."""

    def __init__(self, tri_verts, tri_mesh, names, receiver, rays, emitter, center, yaw=0.0, materials=(),
                 base_power=3.62, energy_thres=0.0, max_bounces=100, hrtf=0.9, sample_rate=16000, ir_seconds=1,
                 mono=False, bands=1, seed=1):
        self.tv = np.ascontiguousarray(tri_verts, np.float32)
        self.tm = np.ascontiguousarray(tri_mesh, np.int32)
        self.names = [str(n) for n in names]
        self.receiver = receiver
        self.rays = rays
        self.emitter, self.center, self.yaw = emitter, center, yaw
        self.materials = list(materials)      # [(name, absorption|list, scattering)]
        self.base_power, self.energy_thres, self.max_bounces = base_power, energy_thres, max_bounces
        self.hrtf, self.sample_rate, self.ir_seconds = hrtf, sample_rate, ir_seconds
        self.mono, self.bands, self.seed = mono, bands, seed

    # ---- CUDA path through the C ABI
    def renderer(self, **kw):
        scene = arv.Scene.from_triangles(self.tv, self.tm, self.names)
        recv = arv.Receiver.from_triangles(*self.receiver) if self.receiver is not None else None
        r = arv.AudioRenderer(scene, self.ir_seconds, self.sample_rate, self.materials, self.rays, receiver=recv,
                              bands=self.bands, **kw)
        r.setMonoOutput(self.mono)
        r.setBasePower(self.base_power)
        r.setThresholds(self.energy_thres, self.max_bounces)
        r.set_hrtf_absorption_rate(self.hrtf)
        r.setEmitterPosInOptix(self.emitter)
        r.setSphereCenterInOptix(self.center, self.yaw)
        r.set_seed(self.seed)
        return r

    # ---- oracle
    def flat(self, center=None, yaw=None):
        model = oracle_model(self.tv, self.tm, self.names)
        rt = osc.ReceiverTemplate(*self.receiver) if self.receiver is not None else None
        fs = osc.flatten(model, rt, self.center if center is None else center, self.yaw if yaw is None else yaw, [], bands=self.bands)
        lut = {m[0]: m for m in self.materials}
        for i, n in enumerate(self.names):
            if n in lut:
                a = lut[n][1]
                vals = list(a) if isinstance(a, (list, tuple, np.ndarray)) else [a] * self.bands
                fs.absorption[i, :] = np.array(vals[: self.bands], np.float32)
                fs.scattering[i] = lut[n][2] if len(lut[n]) > 2 else 0.0
        return fs

    def params(self, center=None):
        c = self.center if center is None else center
        return oracle.make_params(rays=self.rays, emitter=self.emitter, sphere_center=c, base_power=self.base_power,
                                  energy_thres=self.energy_thres, max_bounces=self.max_bounces, hrtf=self.hrtf,
                                  sample_rate=self.sample_rate, mono=self.mono,
                                  ir_length=self.ir_seconds * self.sample_rate, bands=self.bands, seed=self.seed)

    def oracle_run(self, center=None, yaw=None, **kw):
        o = oracle.trace(self.params(center), self.flat(center, yaw), **kw)
        o["ir_left"], o["ir_right"] = oracle.finalize_ir(o["hist"], self.mono)
        return o


def check_parity(rec, l, r, segs, o, min_match=0.9999):
    """north_star tolerances: >= 99.99 % of rays in the same receiver-hit bin, per-bin IR
    energy within 1e-4 relative, total energy within 1e-5."""
    match = float(np.mean(rec["bin"] == o["bin"]))
    assert match >= min_match, f"bin parity {match}"
    same = rec["bin"] == o["bin"]
    assert np.array_equal(rec["ear"][same], o["ear"][same])
    ol, orr = o["ir_left"], o["ir_right"]
    tot, otot = float(l.astype(np.float64).sum() + r.astype(np.float64).sum()), float(ol.astype(np.float64).sum() + orr.astype(np.float64).sum())
    assert abs(tot - otot) <= 1e-5 * max(abs(otot), 1e-30), (tot, otot)
    if match == 1.0:
        # identical rays -> every bin must agree to fp32 accumulation accuracy
        for a, b in ((l, ol), (r, orr)):
            nz = b != 0
            assert np.array_equal(a != 0, nz)
            assert np.all(np.abs(a[nz] - b[nz]) <= 1e-4 * np.abs(b[nz]))
        assert segs == o["segments"]
        assert np.array_equal(rec["nseg"], o["nseg"])
        assert np.array_equal(rec["energy"], o["energy"])
    return match
