"""Parity of the CUDA convolver (through the C ABI) against the oracle: fp64 direct
linear convolution (north_star: <= 1e-5 relative L2) and the restated reference file /
live semantics."""
import os

import numpy as np
import pytest

import audiorenderingv2_b200 as arv
import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL = 1e-5


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def renderer(sample_rate, ir_seconds):
    scene = arv.Scene.from_triangles(np.zeros((0, 3, 3), np.float32), np.zeros(0, np.int32), [])
    return arv.AudioRenderer(scene, ir_seconds, sample_rate, [], (1, 1, 1))


def decaying_ir(n, seed, rt60_s, fs):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    return (rng.standard_normal(n) * np.exp(-6.9 * t / rt60_s)).astype(np.float32)


def test_file_linear_vs_direct_fp64():
    fs = 16000
    r = renderer(fs, 1)
    hl, hr = decaying_ir(fs, 1, 0.4, fs), decaying_ir(fs, 2, 0.3, fs)
    r.set_ir(hl, hr)
    x = np.load(os.path.join(GOLDEN, "guitar_2s.npy"))
    yl, yr, _, _ = r.convoluteAudioFile(x, arv.CONV_LINEAR)
    assert rel_l2(yl, oracle.direct_conv(x, hl)[: len(x)]) <= TOL
    assert rel_l2(yr, oracle.direct_conv(x, hr)[: len(x)]) <= TOL


def test_file_impulse_and_ragged_lengths():
    fs = 8000
    r = renderer(fs, 1)
    rng = np.random.default_rng(3)
    for n in (1, 511, 512, 513, 5000, 8000 + 77):
        x = rng.standard_normal(n).astype(np.float32)
        d = np.zeros(fs, np.float32); d[0] = 1.0
        s = np.zeros(fs, np.float32); s[700] = 0.5
        r.set_ir(d, s)
        yl, yr, _, _ = r.convoluteAudioFile(x)
        assert np.allclose(yl, x, rtol=0, atol=2e-6 * max(1.0, np.abs(x).max()))
        exp = np.zeros(n, np.float32); exp[700:] = 0.5 * x[: max(0, n - 700)]
        assert np.allclose(yr, exp, rtol=0, atol=2e-6 * max(1.0, np.abs(x).max()))
    # linearity
    a, b = rng.standard_normal(4000).astype(np.float32), rng.standard_normal(4000).astype(np.float32)
    r.set_ir(decaying_ir(fs, 4, 0.2, fs), decaying_ir(fs, 5, 0.2, fs))
    ya = r.convoluteAudioFile(a)[0]; yb = r.convoluteAudioFile(b)[0]; yab = r.convoluteAudioFile(a + b)[0]
    assert rel_l2(yab, (ya.astype(np.float64) + yb)) <= 5e-6


def test_file_reference_mode_c1():
    """BASELINE config 1 convolution semantics: ir_len == fs == 16000 => the reference's
    FFT of size ir_len is a pure circular convolution per second, gain 2, whole seconds
    only, output truncated to the input length."""
    fs = 16000
    r = renderer(fs, 1)
    hl, hr = decaying_ir(fs, 11, 0.5, fs), decaying_ir(fs, 12, 0.25, fs)
    r.set_ir(hl, hr)
    x = np.load(os.path.join(GOLDEN, "guitar_2s.npy"))
    x = np.concatenate([x, x[:5000]])          # 2.3125 s: last partial second is dropped
    yl, yr, _, _ = r.convoluteAudioFile(x, arv.CONV_REFERENCE)
    assert rel_l2(yl, oracle.reference_file_conv(x, hl, fs)) <= TOL
    assert rel_l2(yr, oracle.reference_file_conv(x, hr, fs)) <= TOL


def test_file_reference_mode_two_second_ir():
    fs = 8000
    r = renderer(fs, 2)                        # ir_len = 2 fs: linear when support + fs <= ir_len
    hl = np.zeros(2 * fs, np.float32); hl[: fs // 2] = decaying_ir(fs // 2, 21, 0.1, fs)
    hr = decaying_ir(2 * fs, 22, 0.6, fs)      # wraps
    r.set_ir(hl, hr)
    x = np.random.default_rng(23).standard_normal(3 * fs + 100).astype(np.float32)
    yl, yr, _, _ = r.convoluteAudioFile(x, arv.CONV_REFERENCE)
    assert rel_l2(yl, oracle.reference_file_conv(x, hl, fs)) <= TOL
    assert rel_l2(yr, oracle.reference_file_conv(x, hr, fs)) <= TOL
    # left ear never wraps: equals 2 x the true linear convolution of the whole seconds
    lin = 2.0 * oracle.direct_conv(x[: 3 * fs], hl)[: len(x)]
    assert rel_l2(yl, lin) <= TOL


@pytest.mark.parametrize("block,n_src,ir_len", [(512, 3, 4800), (128, 2, 1000), (256, 1, 256), (1024, 2, 5000), (64, 1, 130), (512, 1, 1000),
                                                (1024, 1, 40000)])
def test_stream_matches_direct(block, n_src, ir_len):
    rng = np.random.default_rng(block + n_src)
    st = arv.ConvStream(n_src, block, ir_len)
    irs = [(decaying_ir(ir_len, 30 + i, 0.05, 48000), decaying_ir(ir_len, 60 + i, 0.03, 48000)) for i in range(n_src)]
    for i, (a, b) in enumerate(irs):
        st.set_ir(i, a, b)
    nb = 3 * ((ir_len + block - 1) // block) + 5       # several trips around the delay line
    x = (0.1 * rng.standard_normal((n_src, nb * block))).astype(np.float32)
    out = np.concatenate([st.process(x[:, k * block:(k + 1) * block]) for k in range(nb)], axis=2)
    for i, (a, b) in enumerate(irs):
        assert rel_l2(out[i, 0], oracle.direct_conv(x[i], a)[: nb * block]) <= TOL
        assert rel_l2(out[i, 1], oracle.direct_conv(x[i], b)[: nb * block]) <= TOL
    # the oracle's own scalar UPOLA port agrees too (it is the CPU baseline of the bench)
    ol, orr, _ = oracle.upola(x[0], irs[0][0], irs[0][1], block)
    assert rel_l2(ol, oracle.direct_conv(x[0], irs[0][0])[: nb * block]) <= 1e-4


@pytest.mark.parametrize("persistent", [False, True], ids=["launch-per-block", "one-launch-cluster-loop"])
def test_stream_back_to_back_steps_overlap_safely(persistent, monkeypatch):
    """Device-resident steps launched back to back overlap (programmatic dependent launch: step k+1 accumulates its
    old partitions while step k finishes).  The result must equal, bit for bit, the same steps run one at a time
    through the host-buffer call (a copy between steps serialises them), and match the fp64 direct convolution."""
    import torch
    if persistent:
        monkeypatch.setenv("ARV2_CONV_PERSISTENT", "1")      # the "blocks" mode below then runs stream_blocks_kernel
    n_src, block, ir_len, nb = 4, 512, 20000, 120
    rng = np.random.default_rng(5)
    irs = [(decaying_ir(ir_len, 130 + i, 0.2, 48000), decaying_ir(ir_len, 160 + i, 0.15, 48000)) for i in range(n_src)]
    x = (0.1 * rng.standard_normal((nb, n_src, block))).astype(np.float32)
    outs = []
    for mode in ("device", "blocks", "host"):
        st = arv.ConvStream(n_src, block, ir_len)
        for i, (a, b) in enumerate(irs):
            st.set_ir(i, a, b)
        if mode == "blocks":                              # the same through the multi-block entry point
            dev = torch.device("cuda", 0)
            dx = torch.from_numpy(x).to(dev)
            dy = torch.empty(nb, n_src, 2, block, device=dev)
            s = torch.cuda.Stream(device=dev)
            torch.cuda.synchronize()
            st.process_device_blocks(dx.data_ptr(), dy.data_ptr(), nb, s.cuda_stream)
            torch.cuda.synchronize()
            outs.append(dy.cpu().numpy())
        elif mode == "device":
            dev = torch.device("cuda", 0)
            dx = torch.from_numpy(x).to(dev)
            dy = torch.empty(nb, n_src, 2, block, device=dev)
            s = torch.cuda.Stream(device=dev)
            torch.cuda.synchronize()
            with torch.cuda.stream(s):
                for k in range(nb):
                    st.process_device(dx[k].data_ptr(), dy[k].data_ptr(), s.cuda_stream)
            torch.cuda.synchronize()
            outs.append(dy.cpu().numpy())
        else:
            outs.append(np.stack([st.process(x[k]) for k in range(nb)]))
        st.close()
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[2])
    y = np.concatenate(list(outs[0]), axis=2)          # [n_src][2][nb*block]
    xs = np.concatenate(list(x), axis=1)
    for i, (a, b) in enumerate(irs):
        assert rel_l2(y[i, 0], oracle.direct_conv(xs[i], a)[: nb * block]) <= TOL
        assert rel_l2(y[i, 1], oracle.direct_conv(xs[i], b)[: nb * block]) <= TOL


def test_stream_ir_swap_and_reset():
    block, ir_len = 256, 2000
    st = arv.ConvStream(1, block, ir_len)
    a = decaying_ir(ir_len, 1, 0.02, 48000); b = decaying_ir(ir_len, 2, 0.02, 48000)
    x = (0.1 * np.random.default_rng(9).standard_normal(40 * block)).astype(np.float32)
    st.set_ir(0, a, a)
    y1 = np.concatenate([st.process(x[k * block:(k + 1) * block])[0, 0] for k in range(20)])
    st.set_ir(0, b, b)                                   # takes effect at the block boundary
    y2 = np.concatenate([st.process(x[k * block:(k + 1) * block])[0, 0] for k in range(20, 40)])
    assert rel_l2(y1, oracle.direct_conv(x[: 20 * block], a)[: 20 * block]) <= TOL
    # after the swap every partition uses the new IR on the full input history
    assert rel_l2(y2[-8 * block:], oracle.direct_conv(x, b)[32 * block: 40 * block]) <= TOL
    st.reset()
    y3 = np.concatenate([st.process(x[k * block:(k + 1) * block])[0, 0] for k in range(5)])
    assert rel_l2(y3, oracle.direct_conv(x[: 5 * block], b)[: 5 * block]) <= TOL


def test_stream_ir_swap_between_overlapped_blocks():
    """An IR swap enqueued -- without any host synchronisation -- between two runs of overlapped device-resident blocks
    on the CALLER'S stream lands exactly at the block boundary: the swap (spectra kernel + pointer flip on the
    convolver's own stream) waits on an event for the steps already enqueued, the steps enqueued after it wait on an
    event for the swap.  Bit-identical to the serialised host-buffer path."""
    import torch
    block, ir_len, n_src, half = 512, 6000, 2, 24
    rng = np.random.default_rng(21)
    irs_a = [(decaying_ir(ir_len, 300 + i, 0.05, 48000), decaying_ir(ir_len, 310 + i, 0.04, 48000)) for i in range(n_src)]
    irs_b = [(decaying_ir(ir_len, 320 + i, 0.05, 48000), decaying_ir(ir_len, 330 + i, 0.04, 48000)) for i in range(n_src)]
    x = (0.1 * rng.standard_normal((2 * half, n_src, block))).astype(np.float32)
    dev = torch.device("cuda", 0)

    def run(device_path):
        st = arv.ConvStream(n_src, block, ir_len)
        for i, (a, b) in enumerate(irs_a):
            st.set_ir(i, a, b)
        if device_path:
            dx = torch.from_numpy(x).to(dev)
            dy = torch.empty(2 * half, n_src, 2, block, device=dev)
            s = torch.cuda.Stream(device=dev)
            torch.cuda.synchronize()
            st.process_device_blocks(dx.data_ptr(), dy.data_ptr(), half, s.cuda_stream)
            # no synchronisation here: set_ir works on the convolver's own stream and orders itself against the steps
            # in flight on `s` with events (the swap waits for them, the next steps wait for the swap)
            for i, (a, b) in enumerate(irs_b):
                st.set_ir(i, a, b)
            st.process_device_blocks(dx[half].data_ptr(), dy[half].data_ptr(), half, s.cuda_stream)
            torch.cuda.synchronize()
            out = dy.cpu().numpy()
        else:
            out = np.empty((2 * half, n_src, 2, block), np.float32)
            for k in range(2 * half):
                if k == half:
                    for i, (a, b) in enumerate(irs_b):
                        st.set_ir(i, a, b)
                out[k] = st.process(x[k])
        st.close()
        return out

    assert np.array_equal(run(True), run(False))


def test_live_reference_semantics():
    """convoluteLiveInput (one 4096-sample callback, OR/AudioRenderer.cpp:593-661): the
    first ir_len output samples of the stream convolver x2 equal the reference's circular
    result whenever it does not wrap (4096 + support <= ir_len)."""
    fs, ir_len = 44100, 44100
    hl = np.zeros(ir_len, np.float32); hl[:20000] = decaying_ir(20000, 5, 0.1, fs)
    hr = np.zeros(ir_len, np.float32); hr[:30000] = decaying_ir(30000, 6, 0.2, fs)
    x = np.random.default_rng(7).standard_normal(4096)
    ref = oracle.reference_live_conv(x, hl, hr)         # interleaved LRLR, 2*ir_len doubles
    st = arv.ConvStream(1, 512, ir_len)
    st.set_ir(0, hl, hr)
    xin = np.zeros(ir_len + 512, np.float32); xin[:4096] = x
    nb = ir_len // 512
    out = np.concatenate([st.process(xin[k * 512:(k + 1) * 512]) for k in range(nb)], axis=2)
    assert rel_l2(2.0 * out[0, 0], ref[0:2 * nb * 512:2]) <= TOL
    assert rel_l2(2.0 * out[0, 1], ref[1:2 * nb * 512:2]) <= TOL


def test_live_callback_fills_ring_like_the_reference():
    """audioHandlerWithMic -> convoluteLiveInput -> CircularBuffer (OR/main.cpp:99-135): after one
    4096-sample callback the ring holds the interleaved, 2x-gain convolution of the block; popping
    nFrames*2 values gives what the reference would hand to RtAudio (no wrap: 4096 + support <= ir_len)."""
    fs, ir_len = 44100, 44100
    hl = np.zeros(ir_len, np.float32); hl[:3000] = decaying_ir(3000, 5, 0.02, fs)
    hr = np.zeros(ir_len, np.float32); hr[:2000] = decaying_ir(2000, 6, 0.02, fs)
    x = np.random.default_rng(8).standard_normal(4096)
    st = arv.ConvStream(1, 512, ir_len)
    st.set_ir(0, hl, hr)
    ring = arv.Ring(2 * ir_len)
    arv.live_callback(st, x, ring)
    got = ring.get_and_reset(2 * 4096)
    ref = oracle.reference_live_conv(x, hl, hr)[: 2 * 4096]
    assert rel_l2(got, ref) <= TOL


def test_c5_sixteen_sources_two_second_ir():
    """BASELINE configs[4] at its full size: 16 sources x 96 000-tap stereo IR (2 s @48 kHz, 188 partitions of 512),
    400 blocks through arv2_stream_process_device_blocks, against the fp64 direct convolution: source 0 in full (both
    ears, 204 800 outputs x 96 000 taps), every other source on windows that cover the first blocks, the first trip
    around the frequency-domain delay line (blocks 186..190) and the last blocks."""
    import torch
    n_src, block, ir_len, nb, fs = 16, 512, 96000, 400, 48000
    st = arv.ConvStream(n_src, block, ir_len)
    irs = [(decaying_ir(ir_len, 200 + i, 1.2, fs), decaying_ir(ir_len, 300 + i, 1.2, fs)) for i in range(n_src)]
    for i, (a, b) in enumerate(irs):
        st.set_ir(i, a, b)
    rng = np.random.default_rng(100)
    x = (0.1 * rng.standard_normal((n_src, nb * block))).astype(np.float32)
    xb = torch.from_numpy(np.ascontiguousarray(x.reshape(n_src, nb, block).transpose(1, 0, 2))).cuda()     # [nb][src][block]
    yb = torch.empty(nb, n_src, 2, block, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        st.process_device_blocks(xb.data_ptr(), yb.data_ptr(), nb, s.cuda_stream)
    torch.cuda.synchronize()
    out = yb.cpu().numpy().transpose(1, 2, 0, 3).reshape(n_src, 2, nb * block)
    for ear in (0, 1):
        assert rel_l2(out[0, ear], oracle.direct_conv(x[0], irs[0][ear])[: nb * block]) <= TOL
    windows = [(0, 2 * block), (186 * block, 5 * block), ((nb - 2) * block, 2 * block)]
    for i in range(1, n_src):
        for ear in (0, 1):
            got = np.concatenate([out[i, ear, b:b + n] for b, n in windows])
            ref = np.concatenate([oracle.direct_conv_window(x[i], irs[i][ear], b, n) for b, n in windows])
            assert rel_l2(got, ref) <= TOL, (i, ear)
    st.close()


def test_stream_two_quick_swaps_do_not_race():
    """Two swaps of the same source back to back while steps are in flight on another stream: the second one rewrites
    the buffer that was active before the first -- which a running step may still read unless the swap waits for it."""
    import torch
    block, ir_len, n_src, nb = 512, 20000, 3, 40
    rng = np.random.default_rng(31)
    irs = [[(decaying_ir(ir_len, 400 + 10 * v + i, 0.1, 48000), decaying_ir(ir_len, 450 + 10 * v + i, 0.08, 48000)) for i in range(n_src)] for v in range(3)]
    x = (0.1 * rng.standard_normal((2 * nb, n_src, block))).astype(np.float32)
    dev = torch.device("cuda", 0)
    st = arv.ConvStream(n_src, block, ir_len)
    for i, (a, b) in enumerate(irs[0]):
        st.set_ir(i, a, b)
    dx = torch.from_numpy(x).to(dev); dy = torch.empty(2 * nb, n_src, 2, block, device=dev)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    st.process_device_blocks(dx.data_ptr(), dy.data_ptr(), nb, s.cuda_stream)
    for v in (1, 2):
        for i, (a, b) in enumerate(irs[v]):
            st.set_ir(i, a, b)
    st.process_device_blocks(dx[nb].data_ptr(), dy[nb].data_ptr(), nb, s.cuda_stream)
    torch.cuda.synchronize()
    got = dy.cpu().numpy()
    ref = arv.ConvStream(n_src, block, ir_len)
    for i, (a, b) in enumerate(irs[0]):
        ref.set_ir(i, a, b)
    want = np.empty_like(got)
    for k in range(2 * nb):
        if k == nb:
            for i, (a, b) in enumerate(irs[2]):
                ref.set_ir(i, a, b)
        want[k] = ref.process(x[k])
    assert np.array_equal(got, want)
    st.close(); ref.close()


def test_stream_host_blocks_and_stereo_mix():
    """arv2_stream_process_blocks: up to 16 blocks per call through mapped pinned buffers = the same blocks one call at
    a time, bit for bit; the stereo mix (the buffer playback consumes) = the sources summed in order with their gains,
    and matches the fp64 direct convolutions summed."""
    n_src, block, ir_len, nb = 5, 512, 9000, 32
    rng = np.random.default_rng(77)
    irs = [(decaying_ir(ir_len, 500 + i, 0.1, 48000), decaying_ir(ir_len, 520 + i, 0.07, 48000)) for i in range(n_src)]
    x = (0.1 * rng.standard_normal((nb, n_src, block))).astype(np.float32)
    gains = np.array([1.0, 0.5, 2.0, 0.25, 1.5], np.float32)
    a = arv.ConvStream(n_src, block, ir_len); b = arv.ConvStream(n_src, block, ir_len)
    for st in (a, b):
        for i, (l, r) in enumerate(irs):
            st.set_ir(i, l, r)
    a.set_gains(gains)
    one = np.stack([b.process(x[k]) for k in range(nb)])
    outs, mixes = [], []
    for k0 in range(0, nb, 16):
        o, m = a.process_blocks(x[k0:k0 + 16], want_out=True, want_mix=True)
        outs.append(o); mixes.append(m)
    out, mix = np.concatenate(outs), np.concatenate(mixes)
    assert np.array_equal(out, one)
    acc = np.zeros((nb, 2, block), np.float32)
    for s_ in range(n_src):
        acc = (np.float32(gains[s_]) * out[:, s_].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)      # fmaf(g, x, acc)
    assert np.allclose(mix, acc, rtol=2e-7, atol=1e-12)          # one fused multiply-add per source, in source order
    xs = x.transpose(1, 0, 2).reshape(n_src, nb * block)
    for ear in (0, 1):
        want = sum(float(gains[i]) * oracle.direct_conv(xs[i], irs[i][ear])[: nb * block] for i in range(n_src))
        got = mix[:, ear, :].reshape(-1)
        assert rel_l2(got, want) <= TOL
    _, m2 = a.process_blocks(x[:3], want_out=False, want_mix=True)      # mix only: per-source outputs stay on the device
    assert m2.shape == (3, 2, block) and np.isfinite(m2).all()
    with pytest.raises(arv.Arv2Error):
        a.process_blocks(np.zeros((17, n_src, block), np.float32))
    a.close(); b.close()
