"""Soak: thousands of back-to-back overlapped launches must equal the serialised run bit for bit, and mixed block kinds
on their own streams must run without tripping a watchdog."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "orion-sdr_b200", "python"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import orion_b200 as ob
n = 24_000_000
taps = ob.fir_lowpass_design(2.4e6, 100e3, 38400.0)
def mk():
    return ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=8, demod=ob.DEMOD_FM, fs_demod=3e5, p0=25e3, audio_bw_hz=15e3, translate_hz=100e3)
xs = [torch.randn(2 * n, device="cuda") * 0.3 for _ in range(3)]
def run(no_overlap, iters):
    if no_overlap: os.environ["ORION_B200_NO_OVERLAP"] = "1"
    else: os.environ.pop("ORION_B200_NO_OVERLAP", None)
    c = mk()
    y = torch.empty(n // 8, device="cuda")
    t0 = time.perf_counter()
    for i in range(iters):
        c.process_dev(xs[i % 3].data_ptr(), n, y.data_ptr(), n // 8)
    c.synchronize()
    return y.cpu().numpy(), time.perf_counter() - t0
a, ta = run(False, 3000)
b, tb = run(True, 3000)
print("overlapped %.1f us/launch, serialised %.1f us/launch, bit-equal after 3000 launches: %s" % (ta / 3000 * 1e6, tb / 3000 * 1e6, np.array_equal(a.view(np.uint32), b.view(np.uint32))))
# mixed kinds interleaved on their own streams
blocks = [mk(), ob.FirDecimator(2.4e6, 8, 100e3, 38400.0), ob.LpCascade(48e3, 4.5e3), ob.Rotator(1e5, 2.4e6)]
outs = [torch.empty(n // 8, device="cuda"), torch.empty(2 * (n // 8), device="cuda"), torch.empty(2 * n, device="cuda"), torch.empty(2 * n, device="cuda")]
for i in range(400):
    blocks[0].process_dev(xs[i % 3].data_ptr(), n, outs[0].data_ptr(), n // 8)
    blocks[1].process_dev(xs[(i + 1) % 3].data_ptr(), n, outs[1].data_ptr(), n // 8)
    blocks[2].process_dev(xs[(i + 2) % 3].data_ptr(), 2 * n, outs[2].data_ptr(), 2 * n)
    blocks[3].process_dev(xs[i % 3].data_ptr(), n, outs[3].data_ptr(), n)
for b_ in blocks: b_.synchronize()
print("4 block kinds x 400 interleaved launches on 4 streams: ok")
