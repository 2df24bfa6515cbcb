#!/usr/bin/env python
"""bench.py -- the hot path of AudioRenderingV2 on B200: IR trace, IR re-render, streaming
convolution (BASELINE.json metric: "Grays/s IR trace at 1/2/4/8 B200; IR re-render ms;
conv us per 512-sample block").

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU port of the same tracer on host cores

A step = one IR render of this rank's slice of the seeded ray set (arv2_render_range:
zero the fp64 histogram, trace, then all-reduce the histogram over NCCL and finalise).
Workload (BASELINE.json configs[1]): procedural conference-scale room (331k triangles --
conference.obj is a missing blob in the reference checkout), 1M rays per GPU, 50 bounces,
2 s IR @48 kHz, 1 source / 1 receiver.  1 ray = 1 traced segment = 1 closest-hit query.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS = (100, 100, 100)          # per GPU
MAX_BOUNCES = 50
FS = 48000
IR_SECONDS = 2
EMITTER = (2.0, 1.5, 2.0)
RECEIVER = (9.0, 1.4, 5.5)
YAW = 30.0
SEED = 7
BANDS = 1
BYTES_PER_SEGMENT = 17 * 64 + 4 * 36 + 64      # 1296 B for T ~ 332k
WORKLOAD = "c2"


def select_workload(name):
    """c2 = BASELINE configs[1] (the default and the only driver-facing line);
    c4 = configs[3]: synthetic 1M-triangle hall, 8 frequency bands, 10M rays per GPU."""
    global RAYS, EMITTER, RECEIVER, SEED, BANDS, BYTES_PER_SEGMENT, WORKLOAD
    WORKLOAD = name
    if name == "c4":
        RAYS = (1000, 100, 100)
        EMITTER = (8.0, 1.6, 15.0)
        RECEIVER = (30.0, 1.6, 15.0)
        SEED = 11
        BANDS = 8
        BYTES_PER_SEGMENT = 19 * 64 + 4 * 36 + 64 + 56      # 1480 B


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_receiver():
    d = np.load(os.path.join(ROOT, "tests", "golden", "receiver.npz"))
    return d["left"], d["right"]


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled
    through NVML every few ms from a thread (nvidia-smi -lms cannot start fast enough for a
    50 ms region); falls back to one nvidia-smi query when NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm, self.bits = [], 0
        self.stop_flag = False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception:
            self.h = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.bits |= int(fn(self.h))
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        if self.h is not None:
            self.stop_flag = True
            self.th.join(timeout=1.0)
            if self.sm:
                return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max,
                        "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b), "samples": len(self.sm)}
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]),
                    "reasons": [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")],
                    "samples": 1, "note": "single nvidia-smi query right after the timed region"}
        except Exception as e:       # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"clock query unavailable: {e}"], "samples": 0}


def scene_case():
    from audiorenderingv2_b200 import scenes
    if WORKLOAD == "c4":
        tv, tm, names = scenes.atrium()
        return tv, tm, names, scenes.materials(bands=8)
    tv, tm, names = scenes.conference_room()
    return tv, tm, names, scenes.materials()


def oracle_flat(tv, tm, names, mats, recv):
    import oracle  # noqa: F401  (cpu_baseline / --impl reference only)
    from oracle import scene as osc
    model = osc.Model(meshes=[osc.Mesh(names[i], np.ascontiguousarray(tv[tm == i])) for i in range(len(names))])
    flat = osc.flatten(model, osc.ReceiverTemplate(*recv), RECEIVER, YAW, [], bands=BANDS)
    lut = {m[0]: m[1] for m in mats}
    for i, n in enumerate(names):
        a = lut[n]
        flat.absorption[i, :] = np.array(a if isinstance(a, (list, tuple)) else [a] * BANDS, np.float32)
    return flat


def cpu_port_rate(n_total_rays, budget_s, chunk):
    """Times the oracle (CPU port of the identical tracer, built -O3 -march=native on this machine) on all host
    cores on a bounded sample of the same workload.  Returns (Grays/s, cores, sample description, segments/ray)."""
    import oracle
    oracle.use_native()                                       # BASELINE.md section 5; bit-identical to the portable build
    tv, tm, names, mats = scene_case()
    flat = oracle_flat(tv, tm, names, mats, load_receiver())
    prep = oracle.PreparedScene(flat, bands=BANDS)            # BVH build untimed (as on the GPU side)
    p = oracle.make_params(rays=(n_total_rays, 1, 1), emitter=EMITTER, sphere_center=RECEIVER, base_power=100.0,
                           max_bounces=MAX_BOUNCES, hrtf=0.9, sample_rate=FS, ir_length=IR_SECONDS * FS, bands=BANDS, seed=SEED)
    cores = os.cpu_count() or 1
    segs, rays, t = 0, 0, 0.0
    while t < budget_s and rays + chunk <= n_total_rays:
        t0 = time.perf_counter()
        s, _ = prep.trace(p, rays, chunk, n_threads=cores)
        t += time.perf_counter() - t0
        segs += s; rays += chunk
    return segs / t / 1e9, cores, f"first {rays} rays of the same seeded set ({segs} segments, {t:.1f} s), g++ -O3 -march=native", segs / max(rays, 1)


REFERENCE_CHUNK = 100_000       # rays the reference arm traces per step


def run_reference(args):
    """--impl reference: the reference is CUDA/OptiX on Windows and has no CPU path; its
    arm is the CPU port of the identical tracer (oracle/) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total = RAYS[0] * RAYS[1] * RAYS[2] * args.gpus
    chunk = REFERENCE_CHUNK
    import oracle
    oracle.use_native()
    tv, tm, names, mats = scene_case()
    flat = oracle_flat(tv, tm, names, mats, load_receiver())
    prep = oracle.PreparedScene(flat, bands=BANDS)
    p = oracle.make_params(rays=(n_total, 1, 1), emitter=EMITTER, sphere_center=RECEIVER, base_power=100.0,
                           max_bounces=MAX_BOUNCES, hrtf=0.9, sample_rate=FS, ir_length=IR_SECONDS * FS, bands=BANDS, seed=SEED)
    cores = os.cpu_count() or 1
    for i in range(args.warmup):
        prep.trace(p, 0, 20_000, n_threads=cores)
    t, segs = 0.0, 0
    for i in range(args.steps):
        t0 = time.perf_counter()
        s, _ = prep.trace(p, (i * chunk) % max(1, n_total - chunk), chunk, n_threads=cores)
        t += time.perf_counter() - t0
        segs += s
    val = segs / t / 1e9
    cfg = workload_config(args.gpus)
    cfg["reference_rays_per_step"] = chunk
    cfg["reference_sample"] = (f"each step traces {chunk} rays of the {n_total}-ray seeded set (a rate metric: Grays/s does not "
                               "depend on the sample size); scene, bounces, IR length as in `workload`")
    line = {
        "impl": "reference", "metric": "Grays/s IR trace", "value": val, "unit": "Grays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "Grays/s", "cores": cores, "kind": "port",
                         "sample": f"{chunk} rays per step of the same seeded set, oracle built g++ -O3 -march=native on this box "
                                   "(OptiX reference not buildable offline)"},
        "e2e": {"value": val, "unit": "Grays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    desc = ("configs[1]: conference-scale room (procedural stand-in, 331004 triangles + 1020 receiver), "
            "1M rays per GPU, 50 bounces, 2 s IR @48 kHz")
    if WORKLOAD == "c4":
        desc = ("configs[3]: synthetic 1M-triangle hall (1048584 triangles + 1020 receiver), 8 frequency bands, "
                "10M rays per GPU, 50 bounces, 2 s IR @48 kHz")
    return {"workload": desc,
            "rays_per_gpu": RAYS[0] * RAYS[1] * RAYS[2], "total_rays": RAYS[0] * RAYS[1] * RAYS[2] * n_gpus,
            "max_bounces": MAX_BOUNCES, "sample_rate": FS, "ir_seconds": IR_SECONDS, "bands": BANDS,
            "parallelism": (f"ray sharding x{n_gpus}: rank r traces the direction tiles t = r (mod {n_gpus}) of the one seeded ray set "
                            "(arv2_set_shard_mode(0) / ARV2_SHARD_CONTIGUOUS=1: contiguous slices of ray ids), ncclAllReduce of the fp64 IR "
                            "histogram inside libarv2 (arv2_render_sharded)") if n_gpus > 1 else
                           "1 GPU (arv2_render_sharded with a 1-rank communicator: no exchange)",
            "l2": "flushed between timed steps (512 MiB write)"}


class Job:
    """One rank of the bench: device, process group, the library's own NCCL communicator."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import audiorenderingv2_b200 as arv
        self.torch, self.dist, self.arv = torch, dist, arv
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} needs {args.gpus} ranks (launch with torch.distributed.run); WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        # libarv2's communicator: rank 0 makes the NCCL id inside the library, torch.distributed only carries the 128 bytes
        uid = torch.zeros(arv.COMM_ID_BYTES, dtype=torch.uint8, device=self.dev)
        if self.rank == 0:
            uid = torch.frombuffer(bytearray(arv.Comm.unique_id()), dtype=torch.uint8).to(self.dev)
        if self.world > 1:
            dist.broadcast(uid, 0)
        self.comm = arv.Comm(self.local, self.rank, self.world, bytes(uid.cpu().numpy().tobytes()))
        self.stream = torch.cuda.Stream(device=self.dev)
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_max_sum(self, per_step_ms, total):
        """(sum over steps of the max over ranks of the step time, sum over ranks of `total`)"""
        t = self.torch.tensor(per_step_ms, dtype=self.torch.float64, device=self.dev)
        s = self.torch.tensor([float(total)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(s)
        return float(t.sum().item()), float(s.item())

    def renderer(self, scene, receiver, mats, total_rays, **kw):
        r = self.arv.AudioRenderer(scene, IR_SECONDS, FS, mats, (total_rays, 1, 1), receiver=receiver, device=self.local, bands=BANDS, **kw)
        r.setBasePower(100.0); r.setThresholds(0.0, MAX_BOUNCES); r.set_hrtf_absorption_rate(0.9)
        r.setEmitterPosInOptix(EMITTER); r.setSphereCenterInOptix(RECEIVER, YAW); r.set_seed(SEED)
        r.set_stream(self.stream.cuda_stream)
        return r

    def timed_renders(self, r, steps, warmup, sample_clocks=False):
        """`steps` sharded renders (this rank's slice, all-reduce, finalise: arv2_render_sharded), each bracketed by a
        barrier + synchronize and timed with CUDA events on the launching stream; L2 flushed before each."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                r.render_sharded(self.comm)
            self.barrier()
            clocks = ClockSampler(self.local) if sample_clocks else None
            step_ms, kern_ms, segs = [], [], 0
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            for _ in range(steps):
                self.flush.fill_(1)                     # evict the BVH from L2 between timed steps
                self.barrier()
                ev0.record(self.stream)
                kern_ms.append(r.render_sharded(self.comm))
                ev1.record(self.stream)
                torch.cuda.synchronize()
                step_ms.append(ev0.elapsed_time(ev1))
                segs += r.last_segments()
            self.barrier()
            clk = clocks.stop() if clocks else None
        total_ms, total_segs = self.reduce_max_sum(step_ms, segs)
        # per rank: mean trace-kernel time and mean step time (what the slowest rank and the exchange cost)
        mine = torch.tensor([float(np.mean(kern_ms)), float(np.mean(step_ms))], dtype=torch.float64, device=self.dev)
        allr = [torch.zeros_like(mine) for _ in range(self.world)]
        if self.world > 1:
            self.dist.all_gather(allr, mine)
        else:
            allr = [mine]
        return {"value": total_segs / (total_ms * 1e-3) / 1e9, "total_ms": total_ms, "total_segs": total_segs,
                "kernel_ms": float(np.mean(kern_ms)), "local_segs_per_step": segs / steps, "clocks": clk,
                "rank_kernel_ms": [round(float(a[0].item()), 4) for a in allr], "rank_step_ms": [round(float(a[1].item()), 4) for a in allr],
                "tracer": tracer_label(r)}


def tracer_label(r):
    """Which kernel traced the last launch: launches of >= 3M rays go through the bounce-synchronous sweep_kernel
    (arv2_set_sweep_min_rays), smaller ones through the per-SM queues of wave_kernel."""
    n = int(r.last_counters()[2])
    return f"sweep_kernel ({n} sweeps, survivors re-binned by origin cell x direction between sweeps)" if n else "wave_kernel"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="arv2", choices=["arv2", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="skip the re-render / convolution / C3 / C4 sub-metrics")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"], help="c2 = BASELINE configs[1] (default), c4 = configs[3]")
    ap.add_argument("--stats-pass", action="store_true", help="internal: print the traversal tallies of the stats build and exit")
    args = ap.parse_args()
    select_workload(args.workload)
    if os.environ.get("ARV2_BENCH_RAYS"):           # tuning aid (tail-effect experiments); not a driver-facing line
        global RAYS
        RAYS = (int(os.environ["ARV2_BENCH_RAYS"]), 1, 1)
    if os.environ.get("ARV2_BENCH_BOUNCES"):
        global MAX_BOUNCES
        MAX_BOUNCES = int(os.environ["ARV2_BENCH_BOUNCES"])
    args.warmup = max(args.warmup, 3) if args.impl == "arv2" else args.warmup

    if args.impl == "reference":
        return run_reference(args)
    if args.stats_pass:
        return stats_pass()

    job = Job(args)
    torch, dist, arv, world, rank, local, dev = job.torch, job.dist, job.arv, job.world, job.rank, job.local, job.dev

    tv, tm, names, mats = scene_case()
    recv = load_receiver()
    per_gpu = RAYS[0] * RAYS[1] * RAYS[2]
    total_rays = per_gpu * world
    scene = arv.Scene.from_triangles(tv, tm, names)
    receiver = arv.Receiver.from_triangles(*recv)
    t_build = time.perf_counter()
    # the seeded ray set has total_rays rays; rank r traces its share of them (its direction tiles: ~per_gpu rays)
    r = job.renderer(scene, receiver, mats, total_rays)
    t_build = time.perf_counter() - t_build

    # first render of a context and of a new seed: direction keys + radix sort of the ray order + the trace (ADVICE:
    # the steady-state figure below re-uses the order of its seed; the reference draws fresh directions every render)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); r.render_sharded(job.comm); torch.cuda.synchronize(); first_ms = 1e3 * (time.perf_counter() - t0)
    r.set_seed(SEED + 1)
    t0 = time.perf_counter(); r.render_sharded(job.comm); torch.cuda.synchronize(); new_seed_ms = 1e3 * (time.perf_counter() - t0)
    r.set_seed(SEED)

    main_run = job.timed_renders(r, args.steps, args.warmup, sample_clocks=True)
    value, total_ms, total_segs, clk = main_run["value"], main_run["total_ms"], main_run["total_segs"], main_run["clocks"]

    # ---- e2e: the call a user of the reference makes (full_render_cycle minus the convolution): move the receiver
    # (host -> device upload of its sub-tree), render, read both IRs back to host memory.
    ir_bytes = 2 * BANDS * r.ir_length * 4
    e2e_ms, e2e_segs = [], 0
    with torch.cuda.stream(job.stream):
        for k in range(args.steps):
            job.flush.fill_(1)
            job.barrier()
            t0 = time.perf_counter()
            r.setSphereCenterInOptix((RECEIVER[0] + 0.01 * (k + 1), RECEIVER[1], RECEIVER[2]), YAW)
            r.render_sharded(job.comm)
            r.get_ir()
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
            e2e_segs += r.last_segments()
    e2e_total_ms, e2e_total_segs = job.reduce_max_sum(e2e_ms, e2e_segs)
    e2e_value = e2e_total_segs / (e2e_total_ms * 1e-3) / 1e9
    upload_bytes = r.last_upload_bytes()
    r.setSphereCenterInOptix(RECEIVER, YAW)

    # ---- the sharded IR is the single-GPU IR (every N the driver runs checks it on hardware)
    sharded_check = None
    if world > 1:
        with torch.cuda.stream(job.stream):
            r.render_sharded(job.comm)
            ls, rs = r.get_ir()
            sharded_segs = job.reduce_max_sum([0.0], r.last_segments())[1]
            job.barrier()
            if rank == 0:
                r.render()                                  # all total_rays rays on this one GPU
                l1, r1 = r.get_ir()
                den = np.maximum(np.abs(l1), 1e-30)
                sharded_check = {"ir_max_rel_diff": float(np.max(np.abs(ls - l1)[l1 != 0] / den[l1 != 0])) if (l1 != 0).any() else 0.0,
                                 "nonzero_bins_equal": bool(np.array_equal(ls != 0, l1 != 0) and np.array_equal(rs != 0, r1 != 0)),
                                 "segments_equal": int(sharded_segs) == r.last_segments(), "rays": total_rays}
                sharded_check["ok"] = bool(sharded_check["ir_max_rel_diff"] <= 1e-6 and sharded_check["nonzero_bins_equal"] and sharded_check["segments_equal"])
            job.barrier()

    kernel_ms = main_run["kernel_ms"]
    segs_per_launch = main_run["local_segs_per_step"]
    achieved = segs_per_launch * BYTES_PER_SEGMENT / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()

    traffic = None
    try:      # per-launch DRAM bytes of the trace kernel from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["trace"]["dram_bytes_per_launch"] if WORKLOAD == "c2" else None
    except (OSError, KeyError, ValueError):
        pass

    extras = {"first_render_ms": first_ms, "new_seed_render_ms": new_seed_ms,
              "render_note": "value / e2e re-use the direction-sorted ray order of their seed (cached per seed and ray range); "
                             "first_render_ms = first render of a context (sort + queue allocation + trace), new_seed_render_ms = a render "
                             "right after arv2_set_seed (sort + trace); the scene BVH is built once per context (bvh_build_s), the "
                             "reference rebuilds its GAS on every move"}
    r.close()
    if not args.skip_extras and WORKLOAD == "c2":
        extras.update(bench_rerender(arv, torch, dev, local, scene, receiver, mats, args))
        extras.update(bench_conv(arv, torch, dist, dev, local, rank, world, args, comm=job.comm))
        extras.update(bench_lbvh(arv, scene, receiver, mats, local))
        extras.update(bench_c3_strong(job, scene, receiver, mats))
        del scene
        extras.update(bench_c4(job, args))
        if rank == 0:
            extras["traversal"] = traversal_figures(local)
        job.barrier()
    if not args.skip_extras and WORKLOAD == "c4":
        # configs[3] "interactive receiver moves": 10M rays x 50 segments x (32 B + 8 bands x 4 B) = 32 GB of cached paths
        job.flush = None
        torch.cuda.empty_cache()
        extras.update(bench_rerender(arv, torch, dev, local, scene, receiver, mats, args))

    line = {
        "metric": "Grays/s IR trace", "value": value, "unit": "Grays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "segments_per_step": total_segs / args.steps, "paths_per_s": total_rays * args.steps / (total_ms * 1e-3),
        "bvh_build_s": t_build,
        "e2e": {"value": e2e_value, "unit": "Grays/s", "h2d_bytes_per_step": upload_bytes,
                "d2h_bytes_per_step": ir_bytes + 8 * 24, "ms_per_step": e2e_total_ms / args.steps},
        "gpu_launches": 2 * args.steps,
        "rank_kernel_ms": main_run["rank_kernel_ms"], "rank_step_ms": main_run["rank_step_ms"],
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": f"wave_kernel<{BANDS},0>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": segs_per_launch * BYTES_PER_SEGMENT,
                     "bytes_per_segment": BYTES_PER_SEGMENT, "kernel_ms": kernel_ms,
                     "note": "nominal: the scene is L2-resident and DRAM idles; what binds is in `traversal`"},
    }
    if sharded_check is not None:
        line["sharded_equals_single"] = sharded_check
    line.update(extras)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_port_rate(per_gpu, 12.0, 50_000)
        line["cpu_baseline"] = {"value": v, "unit": "Grays/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    job.comm.close()
    if world > 1:
        dist.destroy_process_group()


def bench_c3_strong(job, scene, receiver, mats):
    """BASELINE configs[2]: the same room, 100M rays in total, strong-sharded over the N ranks (each traces 100M / N),
    ncclAllReduce of the histogram inside the library."""
    n_total = 100_000_000
    r = job.renderer(scene, receiver, mats, n_total)
    run = job.timed_renders(r, steps=2, warmup=1)
    r.close()
    job.torch.cuda.empty_cache()
    return {"c3_strong_grays_per_s": run["value"], "c3_strong_ms_per_render": run["total_ms"] / 2, "c3_total_rays": n_total,
            "c3_rays_per_gpu": n_total // job.world, "c3_segments_per_render": run["total_segs"] / 2, "c3_tracer": run["tracer"]}


def bench_c4(job, args):
    """BASELINE configs[3] (the target's scene): synthetic 1M-triangle hall, 8 frequency bands, 10M rays per GPU (weak),
    plus the interactive receiver move (re-render from the path cache) on this rank's 10M rays."""
    global RAYS, EMITTER, RECEIVER, SEED, BANDS, BYTES_PER_SEGMENT, WORKLOAD
    saved = (RAYS, EMITTER, RECEIVER, SEED, BANDS, BYTES_PER_SEGMENT, WORKLOAD)
    select_workload("c4")
    try:
        arv, torch = job.arv, job.torch
        tv, tm, names, mats = scene_case()
        scene = arv.Scene.from_triangles(tv, tm, names)
        receiver = arv.Receiver.from_triangles(*load_receiver())
        per_gpu = RAYS[0] * RAYS[1] * RAYS[2]
        r = job.renderer(scene, receiver, mats, per_gpu * job.world)
        run = job.timed_renders(r, steps=3, warmup=2)
        r.close()
        peak = measured_hbm_peak()[0]
        achieved = run["local_segs_per_step"] * BYTES_PER_SEGMENT / (run["kernel_ms"] * 1e-3) / 1e9
        out = {"c4_grays_per_s": run["value"], "c4_ms_per_render": run["total_ms"] / 3, "c4_rays_per_gpu": per_gpu,
               "c4_roofline_frac": achieved / peak, "c4_bytes_per_segment": BYTES_PER_SEGMENT, "c4_kernel_ms": run["kernel_ms"],
               "c4_tracer": run["tracer"]}
        job.flush = None
        torch.cuda.empty_cache()
        rr = bench_rerender(arv, torch, job.dev, job.local, scene, receiver, mats, args)
        t = torch.tensor([rr["rerender_ms"]], dtype=torch.float64, device=job.dev)
        if job.world > 1:
            job.dist.all_reduce(t, op=job.dist.ReduceOp.MAX)
        out.update({"c4_rerender_ms": float(t.item()), "c4_rerender_cached_segments": rr["rerender_cached_segments"],
                    "c4_rerender_roofline_frac": rr["rerender_roofline"]["frac"], "c4_path_cache_build_ms": rr["path_cache_build_ms"]})
        job.flush = torch.empty(512 << 20, dtype=torch.uint8, device=job.dev)
        return out
    finally:
        RAYS, EMITTER, RECEIVER, SEED, BANDS, BYTES_PER_SEGMENT, WORKLOAD = saved


NCU_FIGURES = os.path.join(ROOT, "profiles", "ncu_figures.json")


def stats_pass():
    """Runs in a child process with ARV2_LIB = lib/libarv2_stats.so (trace kernels compiled with -DARV2_TRACE_STATS):
    one render of the workload, prints the traversal tallies."""
    import audiorenderingv2_b200 as arv
    tv, tm, names, mats = scene_case()
    scene = arv.Scene.from_triangles(tv, tm, names)
    receiver = arv.Receiver.from_triangles(*load_receiver())
    n = min(RAYS[0] * RAYS[1] * RAYS[2], int(os.environ.get("ARV2_STATS_RAYS", 10_000_000)))      # the launch size the bench times: the tracer depends on it
    r = arv.AudioRenderer(scene, IR_SECONDS, FS, mats, (n, 1, 1), receiver=receiver, bands=BANDS)
    r.setBasePower(100.0); r.setThresholds(0.0, MAX_BOUNCES); r.set_hrtf_absorption_rate(0.9)
    r.setEmitterPosInOptix(EMITTER); r.setSphereCenterInOptix(RECEIVER, YAW); r.set_seed(SEED)
    r.render()
    ms = r.render()
    c = r.last_counters()
    print(json.dumps({"rays": n, "segments": c[1], "node_visits": c[16], "warp_node_steps": c[17], "leaf_visits": c[18],
                      "tri_tests": c[19], "warp_leaf_steps": c[20], "kernel_ms_with_tallies": ms, "tracer": tracer_label(r)}))


def traversal_figures(local):
    """What north_star asks the bench to evidence for the traversal: node visits per segment and the lanes of a warp
    that are active per traversal step (live, from the stats build of the same kernels), the bytes the traversal asks
    of L1/L2 per second, and ncu's warp-execution efficiency / L2 throughput of the committed capture (labelled)."""
    out = {}
    stats_lib = os.path.join(ROOT, "audiorenderingv2_b200", "lib", "libarv2_stats.so")
    for wl in ("c2", "c4"):
        try:
            env = dict(os.environ, ARV2_LIB=stats_lib, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
                env.pop(k, None)
            if not env["CUDA_VISIBLE_DEVICES"]:
                env["CUDA_VISIBLE_DEVICES"] = str(local)
            else:
                env["CUDA_VISIBLE_DEVICES"] = env["CUDA_VISIBLE_DEVICES"].split(",")[local]
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--stats-pass", "--workload", wl], env=env,
                                 capture_output=True, text=True, timeout=300)
            d = json.loads(res.stdout.strip().splitlines()[-1])
            segs = max(d["segments"], 1)
            out[wl] = {"node_visits_per_segment": d["node_visits"] / segs, "leaf_visits_per_segment": d["leaf_visits"] / segs,
                       "tri_tests_per_segment": d["tri_tests"] / segs,
                       "lanes_per_node_step": d["node_visits"] / max(d["warp_node_steps"], 1),
                       "lanes_per_leaf_step": d["leaf_visits"] / max(d["warp_leaf_steps"], 1),
                       "requested_bytes_per_segment": (64 * d["node_visits"] + 48 * d["tri_tests"]) / segs,
                       "tracer": d.get("tracer"),
                       "source": f"live: lib/libarv2_stats.so (-DARV2_TRACE_STATS), {d['rays']} rays of the workload"}
        except Exception as e:       # noqa: BLE001
            out[wl] = {"error": repr(e)[:200]}
    try:
        out["ncu"] = json.load(open(NCU_FIGURES))
    except (OSError, ValueError):
        out["ncu"] = None
    return out


def measured_hbm_peak():
    """(GB/s, source): MEASURED_PEAKS.json (driver-written) or the profiling recipe's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback"


def bench_rerender(arv, torch, dev, local, scene, receiver, mats, args):
    """IR re-render ms: receiver moves re-deposit from the cached receiver-independent
    paths (BASELINE target: < 1 ms for the conference scene at 1M rays)."""
    per_gpu = RAYS[0] * RAYS[1] * RAYS[2]
    r = arv.AudioRenderer(scene, IR_SECONDS, FS, mats, (per_gpu, 1, 1), receiver=receiver, device=local, path_cache=True, bands=BANDS)
    r.setBasePower(100.0); r.setThresholds(0.0, MAX_BOUNCES); r.set_hrtf_absorption_rate(0.9)
    r.setEmitterPosInOptix(EMITTER); r.setSphereCenterInOptix(RECEIVER, YAW); r.set_seed(SEED)
    build_ms = r.render()
    ms, wall = [], []
    for k in range(max(args.steps, 10) + 3):
        r.setSphereCenterInOptix((RECEIVER[0] - 0.1 * k, RECEIVER[1], RECEIVER[2] - 0.05 * k), YAW + 3.0 * k)
        t0 = time.perf_counter()
        m = r.rerender()
        wall.append(1e3 * (time.perf_counter() - t0))
        ms.append(m)
    segs = r.last_segments()
    cached, cache_bytes = r.path_cache_info()
    r.close()
    med = float(np.median(ms[3:]))
    peak = measured_hbm_peak()[0]
    # 32 B per cached segment is the algorithmic traffic of a re-render; the scan itself streams 8 B per
    # segment (quantised path vertices), so this fraction can exceed what a 32 B scan could reach
    return {"rerender_ms": med, "rerender_wall_ms": float(np.median(wall[3:])),
            "rerender_cached_segments": cached, "rerender_segments_before_first_hit": segs, "path_cache_bytes": cache_bytes,
            "path_cache_build_ms": build_ms,
            "rerender_roofline": {"bound": "hbm", "kernel": "rr_mask_kernel + rr_walk_kernel", "bytes_per_cached_segment": 32,
                                  "achieved": cached * 32 / (med * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": cached * 32 / (med * 1e-3) / 1e9 / peak}}


def bench_lbvh(arv, scene, receiver, mats, local):
    """K1: GPU LBVH build (Morton + radix sort + Karras + refit + collapse) of the same scene,
    and the trace rate on that tree (the host SAH tree is the default for static scenes)."""
    per_gpu = RAYS[0] * RAYS[1] * RAYS[2]
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        r = arv.AudioRenderer(scene, IR_SECONDS, FS, mats, (per_gpu, 1, 1), receiver=receiver, device=local, bvh_builder=1)
        dt = 1e3 * (time.perf_counter() - t0)
        best = dt if best is None else min(best, dt)
        if _ < 2:
            r.close()
    r.setBasePower(100.0); r.setThresholds(0.0, MAX_BOUNCES); r.set_hrtf_absorption_rate(0.9)
    r.setEmitterPosInOptix(EMITTER); r.setSphereCenterInOptix(RECEIVER, YAW); r.set_seed(SEED)
    r.render()
    ms = r.render()
    segs = r.last_segments()
    r.close()
    return {"lbvh_create_ms": best, "lbvh_trace_grays_per_s": segs / (ms * 1e-3) / 1e9}


def bench_conv(arv, torch, dist, dev, local, rank, world, args, comm=None):
    """conv us per 512-sample block: 16 sources x 2 s IR @48 kHz (96000 taps -> 188 partitions), sources sharded over
    the GPUs (BASELINE config 5), each rank's sources mixed to stereo on the device and the per-rank mixes summed onto
    rank 0 with ncclReduce inside libarv2 -- the one stereo buffer playback consumes."""
    n_src_total, block, ir_len = 16, 512, IR_SECONDS * FS
    n_src = max(1, n_src_total // world)
    st = arv.ConvStream(n_src, block, ir_len, device=local)
    rng = np.random.default_rng(200 + rank)
    t = np.arange(ir_len) / FS
    for s in range(n_src):
        env = np.exp(-6.9 * t / 1.2)
        st.set_ir(s, (rng.standard_normal(ir_len) * env).astype(np.float32), (rng.standard_normal(ir_len) * env).astype(np.float32))
    n_blocks = 256
    x = (0.1 * torch.randn(n_blocks, n_src, block, device=dev)).contiguous()
    y = torch.empty(n_src, 2, block, device=dev)
    yb = torch.empty(n_blocks, n_src, 2, block, device=dev)
    mix = torch.empty(n_blocks, 2, block, device=dev)
    s = torch.cuda.Stream(device=dev)

    def run(with_mix):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(s)
        st.process_device_blocks(x.data_ptr(), yb.data_ptr(), n_blocks, s.cuda_stream)    # one call, steps overlap on the device
        if with_mix:
            st.mix_device(yb.data_ptr(), mix.data_ptr(), n_blocks, s.cuda_stream)
            if comm is not None and world > 1:
                comm.reduce_f32(mix.data_ptr(), mix.numel(), 0, s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / n_blocks

    with torch.cuda.stream(s):
        for k in range(32):
            st.process_device(x[k].data_ptr(), y.data_ptr(), s.cuda_stream)
        torch.cuda.synchronize()
        run(True)
        dev_us = run(False)
        mix_us = run(True)
    # host-buffer path (mapped pinned staging, completion word): the live-callback call, one block and 8 blocks
    xin = (0.1 * rng.standard_normal((8, n_src, block))).astype(np.float32)
    for _ in range(8):
        st.process(xin[0])
    t0 = time.perf_counter()
    for _ in range(64):
        st.process(xin[0])
    host_us = 1e6 * (time.perf_counter() - t0) / 64
    t0 = time.perf_counter()
    for _ in range(32):
        st.process_blocks(xin, want_out=False, want_mix=True)
    host8_us = 1e6 * (time.perf_counter() - t0) / (32 * 8)
    tt = torch.tensor([dev_us, host_us, mix_us, host8_us], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    st.close()
    bytes_per_block_src = 3 * 188 * 512 * 8 + 4096 + 6144
    return {"conv_us_per_block": float(tt[0].item()), "conv_us_per_block_host_buffers": float(tt[1].item()),
            "conv_us_per_block_with_stereo_mix": float(tt[2].item()),
            "conv_us_per_block_host_buffers_8_blocks_mix": float(tt[3].item()),
            "conv_sources_per_gpu": n_src, "conv_deadline_us": 1e6 * block / FS,
            "conv_mix": "per-rank device mix of its sources + ncclReduce(sum) onto rank 0 (arv2_stream_mix_device, arv2_comm_reduce_f32)",
            "conv_achieved_gbs": n_src * bytes_per_block_src / (float(tt[0].item()) * 1e-6) / 1e9}


if __name__ == "__main__":
    main()
