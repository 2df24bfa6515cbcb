"""The bounce-synchronous tracer (sweep_kernel: launches of >= arv2_set_sweep_min_rays rays; all paths alive advance
together and the survivors are re-binned by origin cell x direction cell between two sweeps) against the CPU oracle and
against wave_kernel.  It only re-schedules the same per-ray arithmetic, so every per-ray record must be EQUAL to the
per-SM-queue tracer's, whatever the bin geometry or the segments per sweep (OR/devicePrograms.cu:233-252 is the loop
both replace)."""
import numpy as np
import pytest

import audiorenderingv2_b200 as arv
from audiorenderingv2_b200 import scenes
from util import Case, check_parity
from test_trace_gpu import c1

pytestmark = pytest.mark.gpu


def run(case, sweeps, **kw):
    r = case.renderer(record_rays=True, **kw)
    r.set_sweep_min_rays(1 if sweeps else 0)
    r.render()
    l, rr = r.get_ir()
    return r, r.records(), l, rr, r.last_segments()


def n_sweeps(r):
    return int(r.last_counters()[2])


def same(a, b):
    ra, reca, la, rra, sa = a
    rb, recb, lb, rrb, sb = b
    assert sa == sb
    for k in ("bin", "ear", "nseg", "energy"):
        assert np.array_equal(reca[k], recb[k]), k
    assert np.allclose(la, lb, rtol=1e-6, atol=0) and np.allclose(rra, rrb, rtol=1e-6, atol=0)


@pytest.mark.parametrize("env", [{}, {"ARV2_SWEEP_SEGMENTS": "1", "ARV2_SWEEP_FIRST": "1"},
                                 {"ARV2_SWEEP_SEGMENTS": "7", "ARV2_SWEEP_FIRST": "2", "ARV2_SWEEP_CELL_BITS": "5", "ARV2_SWEEP_DIR_BITS": "3"},
                                 {"ARV2_SWEEP_CELL_BITS": "0", "ARV2_SWEEP_DIR_BITS": "2", "ARV2_SWEEP_DIR_MAJOR": "1"}],
                         ids=["default", "one-segment-sweeps", "coarse-dirs", "dir-major-tiny"])
def test_sweeps_equal_oracle_and_wave_c1(golden_scenes, golden_receiver, monkeypatch, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    case = c1(golden_scenes, golden_receiver, hrtf=0.9)
    a = run(case, True)
    first = int(env.get("ARV2_SWEEP_FIRST", 8)); per = int(env.get("ARV2_SWEEP_SEGMENTS", 2))
    assert n_sweeps(a[0]) == 1 + max(0, -(-(case.max_bounces - first) // per))
    b = run(case, False)
    assert n_sweeps(b[0]) == 0
    same(a, b)
    assert check_parity(a[1], a[2], a[3], a[4], case.oracle_run()) == 1.0


def test_sweeps_closed_box_300_bounces(golden_scenes, golden_receiver):
    case = Case(golden_scenes["caja_verts"], golden_scenes["caja_mesh"], golden_scenes["caja_names"], golden_receiver,
                rays=(32, 16, 8), emitter=(3, 1, -2), center=(-8, 4, 6), yaw=20.0, max_bounces=300, ir_seconds=60,
                sample_rate=8000, materials=[("Material.001", 0.01)])
    a = run(case, True)
    assert n_sweeps(a[0]) == 1 + -(-(300 - 8) // 2)
    assert check_parity(a[1], a[2], a[3], a[4], case.oracle_run()) == 1.0
    assert a[1]["nseg"].max() > 150


def test_sweeps_eight_bands_diffuse_and_ranges(golden_scenes, golden_receiver):
    bands = 8
    names = [str(n) for n in golden_scenes["test_names"]]
    rng = np.random.default_rng(5)
    mats = [(n, [float(v) for v in rng.uniform(0.05, 0.6, bands)], s) for n, s in zip(sorted(set(names)), (0.0, 0.5, 1.0))]
    case = c1(golden_scenes, golden_receiver, rays=(100, 100, 3), materials=mats, bands=bands, hrtf=0.7)
    a = run(case, True)
    same(a, run(case, False))
    assert check_parity(a[1], a[2], a[3], a[4], case.oracle_run(), case=case) >= 0.9999
    # ray ranges of the seeded set (the multi-GPU path), accumulated into one histogram
    n = 30000
    r2 = case.renderer(record_rays=True)
    r2.set_sweep_min_rays(1)
    cuts = [0, 9999, 20011, n]
    tot, bins = 0, []
    for i in range(3):
        r2.render_range(cuts[i], cuts[i + 1] - cuts[i], zero_first=(i == 0))
        assert n_sweeps(r2) > 1
        tot += r2.last_segments()
        bins.append(r2.records(cuts[i + 1] - cuts[i])["bin"])
    r2.finalize()
    l2, rr2 = r2.get_ir()
    assert tot == a[4] and np.array_equal(np.concatenate(bins), a[1]["bin"])
    assert np.allclose(l2, a[2], rtol=1e-6, atol=0) and np.allclose(rr2, a[3], rtol=1e-6, atol=0)


def test_sweeps_conference_room_and_path_cache(golden_receiver):
    """The 331k-triangle room of configs[1]: sweeps == wave == oracle, and a path cache FILLED by sweeps (mode 1 of the
    kernel) re-renders receiver moves like a fresh trace."""
    tv, tm, names = scenes.conference_room()
    case = Case(tv, tm, names, golden_receiver, rays=(100, 100, 2), emitter=(2.0, 1.5, 2.0), center=(9.0, 1.4, 5.5),
                yaw=30.0, materials=scenes.materials(), max_bounces=50, sample_rate=48000, ir_seconds=2, hrtf=0.9)
    a = run(case, True)
    same(a, run(case, False))
    assert check_parity(a[1], a[2], a[3], a[4], case.oracle_run(), case=case) >= 0.9999
    rc = case.renderer(record_rays=True, path_cache=True)
    rc.set_sweep_min_rays(1)
    rc.render()
    same((rc, rc.records(), *rc.get_ir(), rc.last_segments()), a)
    rc.setSphereCenterInOptix((6.0, 1.2, 3.0), 75.0)
    rc.rerender()
    moved = (rc, rc.records(), *rc.get_ir(), rc.last_segments())
    fresh = case.renderer(record_rays=True)
    fresh.setSphereCenterInOptix((6.0, 1.2, 3.0), 75.0)
    fresh.render()
    same(moved, (fresh, fresh.records(), *fresh.get_ir(), fresh.last_segments()))


def test_sweeps_degenerate_launches(golden_scenes, golden_receiver):
    """max_bounces 1 (one sweep, nothing handed over), fewer rays than one CTA, no receiver."""
    case = c1(golden_scenes, golden_receiver, rays=(7, 3, 1), max_bounces=1)
    a = run(case, True)
    assert n_sweeps(a[0]) == 1
    assert check_parity(a[1], a[2], a[3], a[4], case.oracle_run()) == 1.0
    case = Case(golden_scenes["caja_verts"], golden_scenes["caja_mesh"], golden_scenes["caja_names"], None,
                rays=(32, 32, 1), emitter=(0, 0, 0), center=(1, 1, 1), max_bounces=11)
    a = run(case, True)
    assert n_sweeps(a[0]) == 1 + -(-(11 - 8) // 2) and not a[2].any() and a[4] > 32 * 32 * 10
    same(a, run(case, False))
