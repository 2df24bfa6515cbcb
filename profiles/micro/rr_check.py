"""re-render vs full trace at scale (C2, 1M rays): identical segment count and IR after receiver moves."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import audiorenderingv2_b200 as arv
torch.cuda.set_device(0)
tv, tm, names, mats = bench.scene_case()
recv = bench.load_receiver()
scene = arv.Scene.from_triangles(tv, tm, names)
receiver = arv.Receiver.from_triangles(*recv)
N = int(os.environ.get("RR_RAYS", "1000000"))
def mk(cache):
    r = arv.AudioRenderer(scene, bench.IR_SECONDS, bench.FS, mats, (N, 1, 1), receiver=receiver, device=0, path_cache=cache)
    r.setBasePower(100.0); r.setThresholds(0.0, bench.MAX_BOUNCES); r.set_hrtf_absorption_rate(0.9)
    r.setEmitterPosInOptix(bench.EMITTER); r.setSphereCenterInOptix(bench.RECEIVER, bench.YAW); r.set_seed(bench.SEED)
    return r
a, b = mk(False), mk(True)
b.render()
ok = True
for k in range(4):
    pos = (bench.RECEIVER[0] - 0.7 * k, bench.RECEIVER[1], bench.RECEIVER[2] - 0.4 * k); yaw = bench.YAW + 20.0 * k
    a.setSphereCenterInOptix(pos, yaw); b.setSphereCenterInOptix(pos, yaw)
    a.render(); ms = b.rerender()
    la, ra = a.get_ir(); lb, rb = b.get_ir()
    sa, sb = a.last_segments(), b.last_segments()
    d = max(float(np.abs(la - lb).max()), float(np.abs(ra - rb).max())); m = float(max(la.max(), ra.max()))
    good = sa == sb and d <= 1e-6 * m
    ok &= good
    print(os.environ.get("LABEL", ""), "move", k, "segments", sa, sb, "max |dIR|", d, "of", m, "rerender ms", round(ms, 4), "OK" if good else "MISMATCH")
print(os.environ.get("LABEL", ""), "ALL OK" if ok else "FAILED")
