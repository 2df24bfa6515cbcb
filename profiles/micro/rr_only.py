"""re-render only: median device ms over receiver moves on the C2 scene (env switches apply)."""
import os, sys, argparse, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import audiorenderingv2_b200 as arv
torch.cuda.set_device(0)
bench.select_workload(os.environ.get("RR_WORKLOAD", "c2"))
tv, tm, names, mats = bench.scene_case()
recv = bench.load_receiver()
scene = arv.Scene.from_triangles(tv, tm, names)
receiver = arv.Receiver.from_triangles(*recv)
try:
    r = bench.bench_rerender(arv, torch, torch.device("cuda", 0), 0, scene, receiver, mats, argparse.Namespace(steps=int(os.environ.get("RR_STEPS", "20"))))
    print(os.environ.get("LABEL", ""), json.dumps(r))
except Exception as ex:
    print(os.environ.get("LABEL", ""), "ERROR", repr(ex)[:300])
