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


def deposit_hist(case, rec, sel, ir_len):
    """The fp64 histogram the rays `sel` of the per-ray records `rec` deposit: the deposit rule of
    __closesthit__radiance (OR/devicePrograms.cu:124-170) in numpy -- primary ear += e, other ear at
    bin + delay (or bin when that overflows) += e * (1 - hrtf) unless mono."""
    bands = rec["energy"].shape[1]
    h = np.zeros((2, bands, ir_len), np.float64)
    b, e, en = rec["bin"][sel], rec["ear"][sel], rec["energy"][sel]
    ok = (e > 0) & (b >= 0) & (b < ir_len)
    b, e, en = b[ok], e[ok], en[ok]
    primary = e - 1                                               # receiver_left (-1) -> ear 1 -> left
    delay = int(float(case.sample_rate) * 0.00044)                # :125
    ob = np.where(b + delay < ir_len, b + delay, b)
    cross = np.float32(1.0) - np.float32(case.hrtf)               # :139
    for k in range(bands):
        np.add.at(h[:, k, :], (primary, b), en[:, k].astype(np.float64))
        if not case.mono:
            np.add.at(h[:, k, :], (1 - primary, ob), (en[:, k] * cross).astype(np.float64))
    return h


def check_parity(rec, l, r, segs, o, min_match=0.9999, case=None):
    """north_star tolerances: >= 99.99 % of rays in the same receiver-hit bin, per-bin IR energy within 1e-4
    relative, total energy within 1e-5.  When some rays differ (float intersection order on large scenes) the per-bin
    check still runs: the flipped rays' deposits are taken out of the oracle's histogram and put back from the CUDA
    path's own per-ray records, so every bin is compared at 1e-4 for all the rays both sides agree on."""
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
    else:
        assert case is not None, "check_parity needs the Case to run the per-bin check when rays differ"
        ir_len = ol.shape[-1]
        differ = ~(same & (rec["ear"] == o["ear"]) & np.all(rec["energy"] == o["energy"], axis=1))
        assert differ.mean() <= 1.0 - min_match + 1e-12, f"{differ.sum()} rays differ"
        take_out, put_in = deposit_hist(case, o, differ, ir_len), deposit_hist(case, rec, differ, ir_len)
        exp_l, exp_r = oracle.finalize_ir(o["hist"] - take_out + put_in, case.mono)
        # bins the correction touched carry the cancellation error of the subtraction (1e-16 of what was there)
        touched = (np.abs(take_out) + np.abs(put_in)).astype(np.float64)
        if case.mono:
            touched = touched.sum(axis=0, keepdims=True).repeat(2, axis=0)
        for a, b, t in ((l, exp_l, touched[0]), (r, exp_r, touched[1])):
            err = np.abs(a.astype(np.float64) - b.astype(np.float64))
            assert np.all(err <= 1e-4 * np.abs(b) + 1e-9 * t), f"per-bin energy off by up to {float(np.max(err / np.maximum(np.abs(b), 1e-30))):.3g}"
            assert np.array_equal((a != 0)[t == 0], (b != 0)[t == 0])
    return match
