"""world_size-2 gloo test of the multi-GPU path's host logic: each rank traces its ray
range of the same seeded set into a private fp64 histogram, the histograms are summed
with torch.distributed all_reduce, and the result equals the single-rank render.
(The per-rank tracer here is the oracle -- on a GPU box it is arv2_render_range.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import scene as osc
    from audiorenderingv2_b200 import sharding
    g = os.path.join(ROOT, "tests", "golden")
    sc = np.load(os.path.join(g, "scenes.npz")); rc = np.load(os.path.join(g, "receiver.npz"))
    model = osc.Model(meshes=[osc.Mesh("m", sc["caja_verts"])])
    flat = osc.flatten(model, osc.ReceiverTemplate(rc["left"], rc["right"]), (-8, 4, 6), 0.0, [("m", 0.3)])
    n = 30_001
    p = oracle.make_params(rays=(n, 1, 1), emitter=(3, 1, -2), sphere_center=(-8, 4, 6), max_bounces=12, hrtf=0.9,
                           sample_rate=8000, ir_length=8000, seed=4)
    begin, count = sharding.ray_range(rank, world, n)
    o = oracle.trace(p, flat, ray_begin=begin, n_rays=count, n_threads=2)
    hist = torch.from_numpy(o["hist"].reshape(-1).copy())
    segs = torch.tensor([o["segments"]], dtype=torch.int64)
    sharding.all_reduce_hist(hist, world)
    dist.all_reduce(segs)
    if rank == 0:
        full = oracle.trace(p, flat, n_threads=2)
        np.savez(out_path, ok_hist=np.allclose(hist.numpy(), full["hist"].reshape(-1), rtol=1e-12, atol=0),
                 ok_segs=int(segs.item()) == full["segments"], nonzero=int((hist != 0).sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single(tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    assert bool(r["ok_hist"]) and bool(r["ok_segs"]) and int(r["nonzero"]) > 10
