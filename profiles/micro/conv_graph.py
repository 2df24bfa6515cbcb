import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import audiorenderingv2_b200 as arv
FS=48000; ir_len=2*FS; n_src=16; block=512
torch.cuda.set_device(0); dev=torch.device("cuda",0)
st = arv.ConvStream(n_src, block, ir_len, device=0)
rng = np.random.default_rng(200)
t = np.arange(ir_len)/FS
for s in range(n_src):
    env = np.exp(-6.9*t/1.2)
    st.set_ir(s, (rng.standard_normal(ir_len)*env).astype(np.float32), (rng.standard_normal(ir_len)*env).astype(np.float32))
n_blocks=256
x = (0.1*torch.randn(n_blocks, n_src, block, device=dev)).contiguous()
y = torch.empty(n_src, 2, block, device=dev)
s = torch.cuda.Stream(device=dev)
with torch.cuda.stream(s):
    for k in range(32): st.process_device(x[k].data_ptr(), y.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter()
    e0.record(s)
    for k in range(n_blocks): st.process_device(x[k].data_ptr(), y.data_ptr(), s.cuda_stream)
    e1.record(s)
    t1=time.perf_counter()
    torch.cuda.synchronize()
    print("stream loop: %.2f us/block device, cpu enqueue %.2f us/block" % (1e3*e0.elapsed_time(e1)/n_blocks, 1e6*(t1-t0)/n_blocks))
    # graph
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(n_blocks): st.process_device(x[k].data_ptr(), y.data_ptr(), s.cuda_stream)
        g.replay(); torch.cuda.synchronize()
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
        print("graph replay: %.2f us/block" % (1e3*e0.elapsed_time(e1)/n_blocks))
    except Exception as ex:
        print("graph failed:", repr(ex)[:300])
