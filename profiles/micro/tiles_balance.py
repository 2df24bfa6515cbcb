"""work balance of the direction-tile shards: the 8 shards of an 8M-ray (C2 room) launch traced one after the other on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import audiorenderingv2_b200 as arv
bench.select_workload(os.environ.get("WL", "c2"))
tv, tm, names, mats = bench.scene_case()
scene = arv.Scene.from_triangles(tv, tm, names)
receiver = arv.Receiver.from_triangles(*bench.load_receiver())
R = 8
per = bench.RAYS[0] * bench.RAYS[1] * bench.RAYS[2]
r = arv.AudioRenderer(scene, bench.IR_SECONDS, bench.FS, mats, (per * R, 1, 1), receiver=receiver, bands=bench.BANDS)
r.setBasePower(100.0); r.setThresholds(0.0, bench.MAX_BOUNCES); r.set_hrtf_absorption_rate(0.9)
r.setEmitterPosInOptix(bench.EMITTER); r.setSphereCenterInOptix(bench.RECEIVER, bench.YAW); r.set_seed(bench.SEED)
for k in range(R):
    r.render_tiles(k, R, zero_first=(k == 0))
ms, segs = [], []
for k in range(R):
    ms.append(r.render_tiles(k, R, zero_first=(k == 0))); segs.append(r.last_segments())
ms, segs = np.array(ms), np.array(segs, float)
print(os.environ.get("LABEL", ""), "tile bits", os.environ.get("ARV2_TILE_BITS", "12"), "| ms per shard", np.round(ms, 3), "| max/mean time %.4f" % (ms.max() / ms.mean()),
      "| max/mean segments %.4f" % (segs.max() / segs.mean()), "| Grays/s at the slowest rank x 8: %.2f" % (segs.sum() / ms.max() / 1e6))
