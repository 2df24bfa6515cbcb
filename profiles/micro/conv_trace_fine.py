"""where the producer lane's time goes around the issue stall (library built with -DARV2_CONV_TIMING -DARV2_CONV_TRACE -DARV2_CONV_TRACE_FINE)."""
import os, sys, ctypes as C
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
nb = 150
x = (0.1 * torch.randn(nb, n_src, 512, device=dev)).contiguous(); y = torch.empty(nb, n_src, 2, 512, device=dev)
s_ = torch.cuda.Stream(device=dev)
with torch.cuda.stream(s_):
    st.process_device_blocks(x.data_ptr(), y.data_ptr(), nb, s_.cuda_stream); torch.cuda.synchronize()
buf = np.zeros((256, 16, 96), np.uint64)
lib = arv.lib()
lib.arv2_debug_conv_trace.restype = C.c_int
lib.arv2_debug_conv_trace.argtypes = [C.c_void_p, C.c_size_t]
assert lib.arv2_debug_conv_trace(buf.ctypes.data, buf.nbytes) == 0
for slot in (100, 101):
    for cta in (0, 1, 7):
        c = buf[slot, cta]; t0 = int(c[0])
        rel = lambda v: (int(v) - t0) / 1e3 if v else float("nan")
        print(f"slot {slot} cta {cta}: wait-done {rel(c[37]):.2f} end {rel(c[38]):.2f}")
        for j in range(8):
            print(f"   partition {10 + j}: stage free {rel(c[64 + 10 + j]):.2f}  expect_tx done {rel(c[40 + 3 * j]):.2f}  X copy issued {rel(c[41 + 3 * j]):.2f}  H copy issued {rel(c[42 + 3 * j]):.2f}")
st.close()
