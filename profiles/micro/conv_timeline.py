"""timeline of consecutive streaming steps (library built with -DARV2_CONV_TIMING): 2 sources, one call of 256 blocks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import audiorenderingv2_b200 as arv
dev = torch.device("cuda", 0)
n_src = 2
st = arv.ConvStream(n_src, 512, 96000)
rng = np.random.default_rng(1)
for s in range(n_src):
    st.set_ir(s, rng.standard_normal(96000).astype(np.float32) * 1e-3, rng.standard_normal(96000).astype(np.float32) * 1e-3)
nb = 120
x = (0.1 * torch.randn(nb, n_src, 512, device=dev)).contiguous(); y = torch.empty(nb, n_src, 2, 512, device=dev)
s_ = torch.cuda.Stream(device=dev)
with torch.cuda.stream(s_):
    st.process_device_blocks(x.data_ptr(), y.data_ptr(), nb, s_.cuda_stream); torch.cuda.synchronize()
st.close()
