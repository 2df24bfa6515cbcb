"""host enqueue time against device time per block of arv2_stream_process_device_blocks (one call of 256 blocks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import audiorenderingv2_b200 as arv
dev = torch.device("cuda", 0)
for n_src in (1, 2, 4, 16):
    st = arv.ConvStream(n_src, 512, 96000)
    rng = np.random.default_rng(1)
    for s in range(n_src):
        st.set_ir(s, rng.standard_normal(96000).astype(np.float32) * 1e-3, rng.standard_normal(96000).astype(np.float32) * 1e-3)
    nb = 256
    x = (0.1 * torch.randn(nb, n_src, 512, device=dev)).contiguous(); y = torch.empty(nb, n_src, 2, 512, device=dev)
    s_ = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s_):
        st.process_device_blocks(x.data_ptr(), y.data_ptr(), 32, s_.cuda_stream); torch.cuda.synchronize()
        for rep in range(2):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            t0 = time.perf_counter()
            st.process_device_blocks(x.data_ptr(), y.data_ptr(), nb, s_.cuda_stream)
            t1 = time.perf_counter()
            e1.record(s_); torch.cuda.synchronize()
        print(os.environ.get("LABEL", ""), f"sources {n_src}: device {1e3 * e0.elapsed_time(e1) / nb:.2f} us/block, host enqueue {1e6 * (t1 - t0) / nb:.2f} us/block")
    st.close()
