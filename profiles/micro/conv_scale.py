"""conv scaling probe: us per 512-sample block for (sources, IR length) combinations -- separates per-CTA latency
from chip-wide bandwidth.  ARV2_CONV_NO_PDL=1 serialises the steps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import audiorenderingv2_b200 as arv
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
block = 512
for n_src, ir_len in [(16, 96000), (8, 96000), (4, 96000), (1, 96000), (16, 48000), (16, 24000), (16, 6000), (32, 96000)]:
    st = arv.ConvStream(n_src, block, ir_len, device=0)
    rng = np.random.default_rng(1)
    for s in range(n_src):
        st.set_ir(s, rng.standard_normal(ir_len).astype(np.float32), rng.standard_normal(ir_len).astype(np.float32))
    nb = 256
    x = (0.1 * torch.randn(nb, n_src, block, device=dev)).contiguous()
    y = torch.empty(n_src, 2, block, device=dev)
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        for k in range(32):
            st.process_device(x[k].data_ptr(), y.data_ptr(), s.cuda_stream)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for k in range(nb):
            st.process_device(x[k].data_ptr(), y.data_ptr(), s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / nb
    P = (ir_len + block - 1) // block
    mb = n_src * 3 * P * block * 8 / 1e6
    print(f"n_src {n_src:3d} ir_len {ir_len:6d} P {P:3d}: {us:7.2f} us/block  {mb:6.1f} MB/step  {mb / us * 1e-3:6.2f} TB/s", flush=True)
    st.close()
