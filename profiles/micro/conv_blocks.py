"""device-resident blocks: us per block for n sources, one call of 256 blocks (env switches apply)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import audiorenderingv2_b200 as arv
dev = torch.device("cuda", 0)
for n_src in [int(v) for v in os.environ.get("NSRC_LIST", "1,2,4,16").split(",")]:
    st = arv.ConvStream(n_src, 512, 96000)
    rng = np.random.default_rng(1)
    for s in range(n_src):
        st.set_ir(s, rng.standard_normal(96000).astype(np.float32) * 1e-3, rng.standard_normal(96000).astype(np.float32) * 1e-3)
    nb = 256
    x = (0.1 * torch.randn(nb, n_src, 512, device=dev)).contiguous(); y = torch.empty(nb, n_src, 2, 512, device=dev)
    s_ = torch.cuda.Stream(device=dev)
    res = []
    with torch.cuda.stream(s_):
        for call_blocks in (256, 8):
            st.process_device_blocks(x.data_ptr(), y.data_ptr(), 32, s_.cuda_stream); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            for k in range(0, nb, call_blocks):
                st.process_device_blocks(x[k].data_ptr(), y[k].data_ptr(), call_blocks, s_.cuda_stream)
            e1.record(s_); torch.cuda.synchronize()
            res.append(1e3 * e0.elapsed_time(e1) / nb)
    print(os.environ.get("LABEL", ""), f"sources {n_src}: {res[0]:.2f} us/block (one call of 256), {res[1]:.2f} us/block (calls of 8)")
    st.close()
