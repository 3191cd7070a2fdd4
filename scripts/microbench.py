"""Device-resident throughput of a few block shapes (CUDA events on a real stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "orion-sdr_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import orion_b200 as ob

def bench(name, blk, n_in, in_dtype, out_items, out_dtype, bytes_per_in, reps=30, nbuf=3, exact=False):
    xs = [torch.randn(n_in * (2 if in_dtype == torch.complex64 else 1), device="cuda").view(-1) for _ in range(nbuf)]
    y = torch.empty(out_items * (2 if out_dtype == torch.complex64 else 1), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    blk.set_stream(st.cuda_stream)
    blk.set_option(ob.OPT_OVERLAP_LAUNCHES, int(os.environ.get("OVERLAP", "1")))
    if "USE_TMA" in os.environ:
        blk.set_option(ob.OPT_USE_TMA, int(os.environ["USE_TMA"]))
    exact = bool(exact)
    if exact:
        # blocks whose absolute oscillator phase reaches the output replay the reference's f32 recurrence: a sequential host
        # walk (~3 ns per item) that does not depend on the data.  A streaming caller has it walked ahead of the samples
        # (orion_b200_block_prepare_oscillator); here: enough for the warm-up and the timed calls, reported on its own.
        nwarm = 6
        h0 = blk.exact_host_ms
        blk.prepare_oscillator(n_in, nwarm + reps + 1)
        walk_ms = blk.exact_host_ms - h0
        for i in range(nwarm):
            blk.process_dev(xs[i % nbuf].data_ptr(), n_in, y.data_ptr(), out_items)
        blk.synchronize()
    else:
        t_end = time.perf_counter() + float(os.environ.get("WARM_S", "1.5"))      # sustained load: let the SM clock ramp up
        while time.perf_counter() < t_end:
            for i in range(30):
                blk.process_dev(xs[i % nbuf].data_ptr(), n_in, y.data_ptr(), out_items)
            blk.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(reps):
        blk.process_dev(xs[i % nbuf].data_ptr(), n_in, y.data_ptr(), out_items)
    e1.record(st)
    blk.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = ""
    if exact:
        steps = n_in * (nwarm + reps + 1)
        extra = f"   [oscillator walked ahead on the host: {walk_ms:.0f} ms for {steps/1e6:.0f} M steps = {walk_ms*1e6/steps:.2f} ns/step, outside the timed region]"
    print(f"{name:34s} {ms*1e3:8.1f} us  {n_in/ms/1e6:8.1f} GS/s  {bytes_per_in*n_in/ms/1e6:8.1f} GB/s{extra}", flush=True)

n = int(os.environ.get("N_SAMPLES", "24000000"))
taps = ob.fir_lowpass_design(2.4e6, 100e3, 38400.0)
which = sys.argv[1:] or ["dec", "chain", "fm", "rot", "lp"]
if "dec" in which:
    bench("FirDecimator 63/8 (C32->C32)", ob.FirDecimator(2.4e6, 8, 100e3, 38400.0), n, torch.complex64, n // 8, torch.complex64, 9.0)
if "chain" in which or "chainfm" in which:
    bench("chain C1 FIR/8+NCO+FM+LR4", ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=8, demod=ob.DEMOD_FM, fs_demod=3e5, p0=25e3, audio_bw_hz=15e3, translate_hz=100e3), n, torch.complex64, n // 8, torch.float32, 8.5)
if "chain" in which:
    bench("chain FIR/8+AM (3 sections)", ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=8, demod=ob.DEMOD_AM, fs_demod=3e5, audio_bw_hz=15e3), n, torch.complex64, n // 8, torch.float32, 8.5)
if "fm" in which:
    bench("FmQuadratureDemod rate-1", ob.FmQuadratureDemod(3e5, 25e3, 15e3).with_translate(100e3), n, torch.complex64, n, torch.float32, 12.0)
if "rot" in which:
    bench("Rotator rate-1 (C32->C32)", ob.Rotator(1e5, 2.4e6), n, torch.complex64, n, torch.complex64, 16.0, reps=10, exact=True)
if "lp" in which:
    bench("LpCascade rate-1 (f32->f32)", ob.LpCascade(48e3, 4.5e3), n, torch.float32, n, torch.float32, 8.0)

# ---- the other BASELINE configs at their full sizes (parity for them: tests/test_gpu_parity.py) -------------
if "configs" in which: which = list(which) + ["c2", "c3", "c4"]
if "c2" in which:
    t2 = ob.kaiser_lowpass_taps(201, 0.01, 60.0)
    bench("C2 rot+FIRiq201/25+SSB, 12 M @1.2 MS/s", ob.Chain(mix=ob.MIX_ROTATE, mix_freq_hz=-250e3, mix_fs=1.2e6, fir=ob.FIR_IQ, taps=t2, decim=25,
          demod=ob.DEMOD_SSB, fs_demod=48e3, p0=0.0, audio_bw_hz=2800.0), 12_000_000, torch.complex64, 480_000, torch.float32, 8.16, reps=10, exact=True)
if "c3" in which:
    t3 = ob.fir_lowpass_design(384e3, 10e3, 6144.0)
    extra = np.stack([ob.lp_biquad_design(48e3, 3e3)] * 2)
    bench("C3 FIR63/8+AM+DC+LR4, 38.4 M @384 kS/s", ob.Chain(fir=ob.FIR_DECIM, taps=t3, decim=8, demod=ob.DEMOD_AM, fs_demod=48e3, audio_bw_hz=5e3,
          post_sos=extra), 38_400_000, torch.complex64, 4_800_000, torch.float32, 8.5)
if "c4" in which:
    bench("C4 FirDecimator 1023/32, 100 M @100 MS/s", ob.FirDecimator(100e6, 32, 450e3, 97800.0), 100_000_000, torch.complex64, 3_125_000,
          torch.complex64, 8.25, reps=10, nbuf=2)
