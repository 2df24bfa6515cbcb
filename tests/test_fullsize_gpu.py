"""BASELINE.json's configurations at their real sizes against the CPU oracle (tolerances of north_star: >= 99.99 % of rays
in the same receiver-hit bin, per-bin energy 1e-4, total energy 1e-5):
  C2  conference-scale room, the full 1M rays, 50 bounces, 2 s IR @48 kHz        (restates OR/devicePrograms.cu:62-254)
  C3  a 200k-ray shard deep inside the 100M-ray set (global ray ids beyond 2^26, ray_begin offsets)
  C4  1M-triangle hall, 8 frequency bands, seed 11, 300k rays + interactive receiver moves through the path cache
plus size-independent properties at the full size: the result must not depend on how the rays are scheduled
(breadth-first tasks vs depth-first lanes, direction-sorted vs id order, one launch vs ranges), and a re-render from
the path cache must equal a fresh trace.  Plus the long-path case of the per-depth queues."""
import numpy as np
import pytest

import oracle
from audiorenderingv2_b200 import scenes
from util import Case, check_parity

pytestmark = pytest.mark.gpu

N_FULL = 1_000_000


@pytest.fixture(scope="module")
def c2(golden_receiver):
    tv, tm, names = scenes.conference_room()
    return Case(tv, tm, names, golden_receiver, rays=(N_FULL, 1, 1), emitter=(2.0, 1.5, 2.0), center=(9.0, 1.4, 5.5),
                yaw=30.0, materials=scenes.materials(), base_power=100.0, max_bounces=50, sample_rate=48000, ir_seconds=2,
                hrtf=0.9, seed=7)


def _render(r):
    r.render()
    l, rr = r.get_ir()
    return l, rr, r.last_segments(), r.records()


def test_c2_fullsize_parity_with_oracle(c2):
    """BASELINE configs[1] at its full size: every one of the 1M rays against the oracle."""
    r = c2.renderer(record_rays=True)
    l, rr, segs, rec = _render(r)
    o = c2.oracle_run()
    match = check_parity(rec, l, rr, segs, o, case=c2)
    assert match >= 0.9999
    # the rays both sides agree on took the same number of segments and arrive with the same energy
    same = rec["bin"] == o["bin"]
    assert np.mean(rec["nseg"][same] == o["nseg"][same]) >= 0.9999
    assert abs(segs - o["segments"]) <= 1e-4 * o["segments"]
    assert (rec["ear"] > 0).mean() > 0.2                  # a fifth of the rays reach the receiver in this room
    # the same 1M rays through the bounce-synchronous tracer of large launches (sweep_kernel: every survivor re-binned
    # by origin cell x direction between sweeps): per-ray records EQUAL to the per-SM-queue tracer's
    rs = c2.renderer(record_rays=True)
    rs.set_sweep_min_rays(1)
    ls, rrs, segs_s, recs = _render(rs)
    assert rs.last_counters()[2] == 1 + (50 - 8) // 2 and r.last_counters()[2] == 0
    assert segs_s == segs
    for key in ("bin", "ear", "nseg", "energy"):
        assert np.array_equal(recs[key], rec[key]), key
    assert np.allclose(ls, l, rtol=1e-6, atol=0) and np.allclose(rrs, rr, rtol=1e-6, atol=0)


def test_c3_shard_deep_in_the_100m_ray_set(c2):
    """BASELINE configs[2]: 100M rays sharded over ranks.  One 200k-ray shard that starts at ray 87 500 000 (the slice
    of rank 7 of 8) against the oracle on the same global ray ids: directions are f(seed, global id), the energy per
    ray is base_power / (1e8 * 4pi/3), records are indexed by the position in the shard."""
    n_total, begin, n = 100_000_000, 87_500_000, 200_000
    case = Case(c2.tv, c2.tm, c2.names, c2.receiver, rays=(n_total, 1, 1), emitter=c2.emitter, center=c2.center, yaw=c2.yaw,
                materials=c2.materials, base_power=100.0, max_bounces=50, sample_rate=48000, ir_seconds=2, hrtf=0.9, seed=7)
    r = case.renderer(record_rays=True)
    r.render_range(begin, n, zero_first=True)
    r.finalize()
    l, rr = r.get_ir()
    rec = r.records()
    assert len(rec["bin"]) == n
    o = oracle.trace(case.params(), case.flat(), ray_begin=begin, n_rays=n)
    o["ir_left"], o["ir_right"] = oracle.finalize_ir(o["hist"])
    assert check_parity(rec, l, rr, r.last_segments(), o, case=case) >= 0.9999
    # and the shard is not the first 200k rays of the set
    o0 = oracle.trace(case.params(), case.flat(), ray_begin=0, n_rays=2000)
    assert not np.array_equal(o0["bin"], o["bin"][:2000])


@pytest.fixture(scope="module")
def c4(golden_receiver):
    """BASELINE configs[3]: synthetic 1M-triangle hall, 8 frequency bands (bench.py --workload c4 geometry)."""
    tv, tm, names = scenes.atrium()
    return Case(tv, tm, names, golden_receiver, rays=(300_000, 1, 1), emitter=(8.0, 1.6, 15.0), center=(30.0, 1.6, 15.0),
                yaw=30.0, materials=scenes.materials(bands=8), base_power=100.0, max_bounces=50, sample_rate=48000,
                ir_seconds=2, hrtf=0.9, bands=8, seed=11)


def test_c4_million_triangle_hall_parity_and_receiver_moves(c4):
    assert len(c4.tv) >= 1_000_000
    r = c4.renderer(record_rays=True, path_cache=True)
    r.render()                                    # builds the path cache, then deposits from it
    l, rr = r.get_ir()
    o = c4.oracle_run()
    assert check_parity(r.records(), l, rr, r.last_segments(), o, case=c4) >= 0.9999
    assert l.shape == (8, 96000) and (l.sum(axis=1) > 0).all()
    # a fresh full trace (no cache) gives the same IR
    f = c4.renderer(record_rays=True)
    f.render()
    lf, rf = f.get_ir()
    assert check_parity(f.records(), lf, rf, f.last_segments(), o, case=c4) >= 0.9999
    # ... and so does the tracer bench.py's C4 leg runs at 10M rays (sweep_kernel, 8 bands: 80 B path states)
    fs = c4.renderer(record_rays=True)
    fs.set_sweep_min_rays(1)
    fs.render()
    assert fs.last_counters()[2] > 1
    lfs, rfs = fs.get_ir()
    for key in ("bin", "ear", "nseg", "energy"):
        assert np.array_equal(fs.records()[key], f.records()[key]), key
    assert np.allclose(lfs, lf, rtol=1e-6, atol=0) and np.allclose(rfs, rf, rtol=1e-6, atol=0)
    # interactive receiver moves: re-deposit from the cached paths, against the oracle at the new position
    for k in (1, 2):
        pos, yaw = (30.0 - 1.5 * k, 1.6 + 0.2 * k, 15.0 + 2.0 * k), 30.0 + 40.0 * k
        r.setSphereCenterInOptix(pos, yaw)
        r.rerender()
        l, rr = r.get_ir()
        om = c4.oracle_run(center=pos, yaw=yaw)
        assert check_parity(r.records(), l, rr, r.last_segments(), om, case=c4) >= 0.9999


def test_c2_fullsize_result_is_scheduling_invariant(c2, monkeypatch):
    """wave_kernel (32-path tasks, per-depth queues, direction-sorted start order, 2048 paths alive per SM: the
    cap binds at this size) against the depth-first kernel in ray-id order, and against three ray ranges."""
    ra = c2.renderer(record_rays=True)
    la, rra, sa, reca = _render(ra)
    monkeypatch.setenv("ARV2_NO_WAVE", "1")
    rb = c2.renderer(record_rays=True)
    monkeypatch.delenv("ARV2_NO_WAVE")
    rb.set_coherent_order(False)
    lb, rrb, sb, recb = _render(rb)
    assert sa == sb and sa > 20 * N_FULL
    for key in ("bin", "ear", "nseg", "energy"):
        assert np.array_equal(reca[key], recb[key]), key
    assert np.allclose(la, lb, rtol=1e-6, atol=0) and np.allclose(rra, rrb, rtol=1e-6, atol=0)
    # conservation: what the rays deposit is what the bins hold (primary ear + hrtf-attenuated other ear)
    hit = reca["ear"] > 0
    inside = hit & (reca["bin"] >= 0) & (reca["bin"] < la.shape[-1])
    dep = reca["energy"].reshape(len(hit), -1)[:, 0].astype(np.float64)
    tot = float(la.astype(np.float64).sum() + rra.astype(np.float64).sum())
    assert abs(tot - dep[inside].sum() * (1.0 + (1.0 - c2.hrtf))) <= 1e-5 * tot
    # ray ranges (the multi-GPU sharding) reproduce the single launch
    cuts = [0, 333_333, 600_001, N_FULL]
    segs = 0
    for i in range(3):
        ra.render_range(cuts[i], cuts[i + 1] - cuts[i], zero_first=(i == 0))
        segs += ra.last_segments()
    ra.finalize()
    lc, rrc = ra.get_ir()
    assert segs == sa
    assert np.allclose(lc, la, rtol=1e-6, atol=0) and np.allclose(rrc, rra, rtol=1e-6, atol=0)


def test_c2_fullsize_rerender_equals_trace(c2):
    """IR re-render from the path cache after receiver moves = a fresh full trace at that position."""
    full = c2.renderer()
    cached = c2.renderer(path_cache=True)
    cached.render()
    for k in (1, 3):
        pos, yaw = (9.0 - 0.7 * k, 1.4, 5.5 - 0.4 * k), 30.0 + 20.0 * k
        full.setSphereCenterInOptix(pos, yaw); cached.setSphereCenterInOptix(pos, yaw)
        full.render(); ms = cached.rerender()
        la, ra = full.get_ir(); lb, rb = cached.get_ir()
        assert full.last_segments() == cached.last_segments()
        assert np.allclose(la, lb, rtol=1e-6, atol=0) and np.allclose(ra, rb, rtol=1e-6, atol=0)
        assert ms < 5.0


def test_long_paths_use_multi_segment_queues(golden_scenes, golden_receiver):
    """300 bounces in the closed box: more depths than queues, so a task advances 5 segments; parity with the
    oracle ray for ray."""
    case = Case(golden_scenes["caja_verts"], golden_scenes["caja_mesh"], golden_scenes["caja_names"], golden_receiver,
                rays=(32, 16, 8), emitter=(3, 1, -2), center=(-8, 4, 6), yaw=20.0, max_bounces=300, ir_seconds=60,
                sample_rate=8000, materials=[("Material.001", 0.01)])
    r = case.renderer(record_rays=True)
    r.render()
    l, rr = r.get_ir()
    o = case.oracle_run()
    assert check_parity(r.records(), l, rr, r.last_segments(), o) == 1.0
    assert r.records()["nseg"].max() > 150


def test_eight_bands_diffuse_scheduling_invariant(golden_receiver, monkeypatch):
    """8 frequency bands + Lambert bounces (the 80 B path state of the queues), 300k rays so that the per-SM cap
    binds: breadth-first tasks and depth-first lanes must produce the same rays and the same IR."""
    tv, tm, names = scenes.conference_room()
    mats = [(n, a, 0.3) for n, a, _ in scenes.materials(bands=8)]
    case = Case(tv, tm, names, golden_receiver, rays=(300_000, 1, 1), emitter=(2.0, 1.5, 2.0), center=(9.0, 1.4, 5.5),
                yaw=30.0, materials=mats, base_power=100.0, max_bounces=40, sample_rate=48000, ir_seconds=2, hrtf=0.8,
                bands=8, seed=11)
    ra = case.renderer(record_rays=True)
    la, rra, sa, reca = _render(ra)
    monkeypatch.setenv("ARV2_NO_WAVE", "1")
    rb = case.renderer(record_rays=True)
    monkeypatch.delenv("ARV2_NO_WAVE")
    lb, rrb, sb, recb = _render(rb)
    assert sa == sb
    for key in ("bin", "ear", "nseg", "energy"):
        assert np.array_equal(reca[key], recb[key]), key
    assert la.shape[0] == 8 and np.allclose(la, lb, rtol=1e-6, atol=0) and np.allclose(rra, rrb, rtol=1e-6, atol=0)
