"""host-buffer convolver call: us per call for 1 / 8 blocks, 1 and 16 sources (env switches apply)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import audiorenderingv2_b200 as arv
for n_src in (1, 16):
    st = arv.ConvStream(n_src, 512, 96000)
    rng = np.random.default_rng(1)
    for s in range(n_src):
        st.set_ir(s, rng.standard_normal(96000).astype(np.float32) * 1e-3, rng.standard_normal(96000).astype(np.float32) * 1e-3)
    x = (0.1 * rng.standard_normal((8, n_src, 512))).astype(np.float32)
    for _ in range(20): st.process(x[0])
    t0 = time.perf_counter()
    for _ in range(200): st.process(x[0])
    one = 1e6 * (time.perf_counter() - t0) / 200
    for _ in range(5): st.process_blocks(x, want_out=True, want_mix=True)
    t0 = time.perf_counter()
    for _ in range(100): st.process_blocks(x, want_out=True, want_mix=True)
    eight = 1e6 * (time.perf_counter() - t0) / 100
    t0 = time.perf_counter()
    for _ in range(100): st.process_blocks(x, want_out=False, want_mix=True)
    eight_mix = 1e6 * (time.perf_counter() - t0) / 100
    print(os.environ.get("LABEL", ""), f"sources {n_src}: 1 block {one:.1f} us/call, 8 blocks out+mix {eight:.1f} us/call ({eight / 8:.1f}/block), 8 blocks mix only {eight_mix:.1f} us/call")
    st.close()
