"""world_size-2 gloo test of the multi-GPU path's host logic.  The product has no CPU tracer (no fallback), so what runs
here of the PRODUCT is the host side of the sharded render: the library's own shard rule (arv2_shard_range), the NCCL
unique id made by rank 0 inside libarv2 (arv2_comm_unique_id) and handed to the other rank over the process group -- the
bootstrap bench.py uses -- and the refusal to make a communicator without a CUDA device.  The per-rank trace of the
rule's slice is done by the oracle; the summed histogram must equal the single-rank render.  (On GPUs:
tests/test_multigpu_gpu.py and bench.py's sharded-equals-single check.)"""
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
    import audiorenderingv2_b200 as arv
    begin, count = arv.shard_range(n, rank, world)                      # the library's rule
    assert (begin, count) == sharding.ray_range(rank, world, n)
    uid = torch.zeros(arv.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(arv.Comm.unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)                                              # how bench.py bootstraps arv2_comm_create
    uid_ok = int(uid.sum().item()) > 0
    if not torch.cuda.is_available():
        try:
            arv.Comm(0, rank, world, bytes(uid.numpy().tobytes()))
            uid_ok = False
        except arv.Arv2Error as e:
            uid_ok = uid_ok and "error -3" in str(e)
    o = oracle.trace(p, flat, ray_begin=begin, n_rays=count, n_threads=2)
    hist = torch.from_numpy(o["hist"].reshape(-1).copy())
    segs = torch.tensor([o["segments"]], dtype=torch.int64)
    sharding.all_reduce_hist(hist, world)
    dist.all_reduce(segs)
    if rank == 0:
        full = oracle.trace(p, flat, n_threads=2)
        np.savez(out_path, ok_hist=np.allclose(hist.numpy(), full["hist"].reshape(-1), rtol=1e-12, atol=0),
                 ok_segs=int(segs.item()) == full["segments"], nonzero=int((hist != 0).sum()), uid_ok=uid_ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single(tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    assert bool(r["ok_hist"]) and bool(r["ok_segs"]) and int(r["nonzero"]) > 10 and bool(r["uid_ok"])
