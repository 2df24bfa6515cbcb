"""conv-only A/B: us per 512-sample block (C5: 16 sources x 96000 taps) for each library in VARIANTS."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, ROOT)
    import argparse, torch, bench
    import audiorenderingv2_b200 as arv
    torch.cuda.set_device(0)
    r = bench.bench_conv(arv, torch, None, torch.device("cuda", 0), 0, 0, 1, argparse.Namespace())
    print(json.dumps(r))
else:
    for d in os.environ.get("VARIANTS", "lib").split():
        env = dict(os.environ, ARV2_LIB=os.path.join(ROOT, "audiorenderingv2_b200", d, "libarv2.so"))
        out = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
        try:
            r = json.loads(out.stdout.strip().splitlines()[-1])
            print(d, round(r["conv_us_per_block"], 2), "us/block device,", round(r["conv_us_per_block_host_buffers"], 1), "us host buffers")
        except Exception:
            print(d, "FAILED", out.stderr[-400:])
