"""Multi-GPU path inside the library (arv2_comm_*, arv2_render_sharded, arv2_multi_*): rays shard over the ranks, the
fp64 histograms are summed by NCCL on the context's stream, every rank ends up with the IR of the whole ray set.
On a one-GPU box the one-rank forms run (no exchange, same code path); with >= 2 GPUs visible the sharded render must
equal the single-GPU one (`gpurun --gpus 2`, and bench.py checks the same at every N it runs)."""
import numpy as np
import pytest

import audiorenderingv2_b200 as arv
from util import Case

pytestmark = pytest.mark.gpu


def case(golden_scenes, golden_receiver):
    return Case(golden_scenes["caja_verts"], golden_scenes["caja_mesh"], golden_scenes["caja_names"], golden_receiver,
                rays=(100, 100, 3), emitter=(3, 1, -2), center=(-8, 4, 6), yaw=20.0, max_bounces=30, ir_seconds=2,
                materials=[("Material.001", 0.2)], seed=9)


def configure(c, r):
    r.setMonoOutput(c.mono); r.setBasePower(c.base_power); r.setThresholds(c.energy_thres, c.max_bounces)
    r.set_hrtf_absorption_rate(c.hrtf); r.setEmitterPosInOptix(c.emitter); r.setSphereCenterInOptix(c.center, c.yaw); r.set_seed(c.seed)


def test_render_sharded_with_one_rank_is_render(golden_scenes, golden_receiver):
    c = case(golden_scenes, golden_receiver)
    a = c.renderer(); a.render(); la, ra = a.get_ir()
    b = c.renderer()
    comm = arv.Comm(0, 0, 1, arv.Comm.unique_id())
    assert comm.nccl_version() >= 21800
    b.render_sharded(comm)
    lb, rb = b.get_ir()
    assert a.last_segments() == b.last_segments()
    assert np.array_equal(la, lb) and np.array_equal(ra, rb)
    comm.close()


def test_multi_renderer_equals_single(golden_scenes, golden_receiver):
    import torch
    c = case(golden_scenes, golden_receiver)
    a = c.renderer(); a.render(); la, ra = a.get_ir()
    scene = arv.Scene.from_triangles(c.tv, c.tm, c.names)
    recv = arv.Receiver.from_triangles(*c.receiver)
    n_dev = torch.cuda.device_count()
    for devices in ([0], list(range(n_dev))) if n_dev > 1 else ([0],):
        m = arv.MultiRenderer(scene, c.ir_seconds, c.sample_rate, c.materials, c.rays, devices, receiver=recv)
        m.each(lambda r: configure(c, r))
        m.render()
        assert sum(r.last_segments() for r in m.renderers) == a.last_segments()
        for r in m.renderers:                      # every device holds the full IR
            l, rr = r.get_ir()
            assert np.allclose(l, la, rtol=1e-6, atol=0) and np.allclose(rr, ra, rtol=1e-6, atol=0)
        # a receiver move on every device, again
        m.each(lambda r: r.setSphereCenterInOptix((-6, 3, 5), 75.0))
        m.render()
        a.setSphereCenterInOptix((-6, 3, 5), 75.0); a.render(); l2, r2 = a.get_ir()
        l, rr = m.renderers[-1].get_ir()
        assert np.allclose(l, l2, rtol=1e-6, atol=0) and np.allclose(rr, r2, rtol=1e-6, atol=0)
        m.close()
        a.setSphereCenterInOptix(c.center, c.yaw); a.render()


def test_sources_sharded_over_two_gpus_mix_to_one_stereo_buffer():
    """BASELINE configs[4]'s sharding: source s lives on GPU s mod R; every rank convolves its sources, mixes them to
    stereo on the device and the per-rank mixes are summed onto rank 0 by ncclReduce inside libarv2.  Rank 0's buffer
    must be the fp64 direct convolution of every source with its IR, summed."""
    import threading
    import torch
    import oracle
    from audiorenderingv2_b200 import sharding
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n_total, block, ir_len, nb = 2, 6, 512, 12000, 40
    rng = np.random.default_rng(8)
    t = np.arange(ir_len) / 48000.0
    irs = [((rng.standard_normal(ir_len) * np.exp(-6.9 * t / 0.15)).astype(np.float32),
            (rng.standard_normal(ir_len) * np.exp(-6.9 * t / 0.1)).astype(np.float32)) for _ in range(n_total)]
    x = (0.1 * rng.standard_normal((n_total, nb * block))).astype(np.float32)
    uid = arv.Comm.unique_id()
    result, errors = {}, []

    def rank_main(rank):
        try:
            torch.cuda.set_device(rank)
            dev = torch.device("cuda", rank)
            mine = sharding.sources_of(rank, world, n_total)
            comm = arv.Comm(rank, rank, world, uid)
            st = arv.ConvStream(len(mine), block, ir_len, device=rank)
            for j, s in enumerate(mine):
                st.set_ir(j, *irs[s])
            xb = torch.from_numpy(np.ascontiguousarray(x[mine].reshape(len(mine), nb, block).transpose(1, 0, 2))).to(dev)
            yb = torch.empty(nb, len(mine), 2, block, device=dev)
            mix = torch.empty(nb, 2, block, device=dev)
            s_ = torch.cuda.Stream(device=dev)
            st.process_device_blocks(xb.data_ptr(), yb.data_ptr(), nb, s_.cuda_stream)
            st.mix_device(yb.data_ptr(), mix.data_ptr(), nb, s_.cuda_stream)
            comm.reduce_f32(mix.data_ptr(), mix.numel(), 0, s_.cuda_stream)
            s_.synchronize()
            result[rank] = mix.cpu().numpy()
            st.close(); comm.close()
        except Exception as e:      # noqa: BLE001
            errors.append((rank, repr(e)))

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join(timeout=120)
    assert not errors, errors
    got = result[0]
    for ear in (0, 1):
        want = sum(oracle.direct_conv(x[s], irs[s][ear])[: nb * block] for s in range(n_total))
        g = got[:, ear, :].reshape(-1)
        assert np.linalg.norm(g - want) <= 1e-5 * np.linalg.norm(want)
