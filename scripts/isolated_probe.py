"""What an isolated C1 launch costs: CUDA events around ONE launch (a) on an idle stream -- the launch has to travel from the
host while the start event is already recorded -- and (b) queued behind a 300 us delay kernel, so that the launch is
already in the queue when the start event executes (device-side duration only).  Median of 15 launches each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "orion-sdr_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch, time
import orion_b200 as ob
n = 24_000_000
taps = ob.fir_lowpass_design(2.4e6, 100e3, 38400.0)
blk = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=8, demod=ob.DEMOD_FM, fs_demod=3e5, p0=25e3, audio_bw_hz=15e3, translate_hz=100e3)
xs = [torch.randn(2 * n, device="cuda") for _ in range(3)]
y = torch.empty(n // 8, device="cuda")
st = torch.cuda.Stream(); blk.set_stream(st.cuda_stream); blk.set_option(ob.OPT_OVERLAP_LAUNCHES, 1)
for i in range(30): blk.process_dev(xs[i % 3].data_ptr(), n, y.data_ptr(), n // 8)
blk.synchronize()
def one(delay):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        if delay: torch.cuda._sleep(600_000)            # ~300 us at 1.9 GHz
    e0.record(st)
    h0 = time.perf_counter()
    blk.process_dev(xs[one.i % 3].data_ptr(), n, y.data_ptr(), n // 8); one.i += 1
    h1 = time.perf_counter()
    e1.record(st)
    blk.synchronize()
    return e0.elapsed_time(e1) * 1e3, (h1 - h0) * 1e6
one.i = 0
for delay in (False, True):
    r = np.array([one(delay) for _ in range(15)])
    print(f"{'queued behind a delay kernel' if delay else 'idle stream               '}: events around one launch {np.median(r[:,0]):6.1f} us (min {r[:,0].min():.1f}), host enqueue {np.median(r[:,1]):5.1f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(20): blk.process_dev(xs[i % 3].data_ptr(), n, y.data_ptr(), n // 8)
e1.record(st); blk.synchronize()
print(f"20 launches back to back: {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us each")
