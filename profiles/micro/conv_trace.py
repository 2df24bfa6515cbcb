"""printf-free timeline of consecutive streaming steps (library built with -DARV2_CONV_TIMING -DARV2_CONV_TRACE)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import audiorenderingv2_b200 as arv
dev = torch.device("cuda", 0)
n_src = int(os.environ.get("NSRC", 2))
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
first = 100
t0 = int(buf[first, 0, 0])
rel = lambda v: (int(v) - t0) / 1e3 if v else float("nan")
for slot in range(first, first + 4):
    for cta in (0, 1, 4, 7):
        c = buf[slot, cta]
        parts = [rel(c[2 + i]) for i in range(34) if c[2 + i]]
        issue = [rel(c[40 + i]) for i in range(24) if c[40 + i]]
        print(f"slot {slot} cta {cta}: start {rel(c[0]):7.2f} ring-ready {rel(c[1]):7.2f} mac-done {rel(c[36]):7.2f} wait-done {rel(c[37]):7.2f} end {rel(c[38]):7.2f} | n={len(parts)}")
        print("     partition done:", " ".join(f"{p:.2f}" for p in parts))
        print("     copies issued: ", " ".join(f"{p:.2f}" for p in issue))
        print("     stage free at: ", " ".join(f"{rel(c[64 + i]):.2f}" for i in range(24) if c[64 + i]))
st.close()
