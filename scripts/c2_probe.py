"""Decomposition of the C2 step (rot + FIRiq201/25 + SSB) into its parts: device-resident, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "orion-sdr_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import orion_b200 as ob

n, m = 12_000_000, 25
t2 = ob.kaiser_lowpass_taps(201, 0.01, 60.0)

def run(name, blk, n_in, cplx_out, out_items, exact, reps=10):
    x = [torch.randn(2 * n_in, device="cuda") for _ in range(2)]
    y = torch.empty(out_items * (2 if cplx_out else 1), device="cuda")
    st = torch.cuda.Stream(); blk.set_stream(st.cuda_stream)
    blk.set_option(ob.OPT_OVERLAP_LAUNCHES, 1)
    if exact: blk.prepare_oscillator(n_in, reps + 6)
    for i in range(5): blk.process_dev(x[i % 2].data_ptr(), n_in, y.data_ptr(), out_items)
    blk.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(reps): blk.process_dev(x[i % 2].data_ptr(), n_in, y.data_ptr(), out_items)
    e1.record(st); blk.synchronize()
    print(f"{name:58s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us", flush=True)

which = sys.argv[1:] or ["fir", "firssb0", "firssb", "c2closed", "c2"]
if "fir" in which:
    run("FirLowpassIq 201 /25 alone (C32->C32)", ob.Chain(fir=ob.FIR_IQ, taps=t2, decim=m), n, True, n // m, False)
if "firssb0" in which:
    b = ob.Chain(fir=ob.FIR_IQ, taps=t2, decim=m, demod=ob.DEMOD_SSB, fs_demod=48e3, p0=0.0, audio_bw_hz=2800.0)
    b.set_option(ob.OPT_EXACT_NCO, 0)
    run("FIR + SSB (closed-form BFO)", b, n, False, n // m, False)
if "firssb" in which:
    run("FIR + SSB (exact BFO)", ob.Chain(fir=ob.FIR_IQ, taps=t2, decim=m, demod=ob.DEMOD_SSB, fs_demod=48e3, p0=0.0, audio_bw_hz=2800.0), n, False, n // m, True)
if "c2closed" in which:
    b = ob.Chain(mix=ob.MIX_ROTATE, mix_freq_hz=-250e3, mix_fs=1.2e6, fir=ob.FIR_IQ, taps=t2, decim=m, demod=ob.DEMOD_SSB, fs_demod=48e3, p0=0.0, audio_bw_hz=2800.0)
    b.set_option(ob.OPT_EXACT_NCO, 0)
    run("C2 closed-form oscillators", b, n, False, n // m, False)
if "c2" in which:
    run("C2 exact oscillators", ob.Chain(mix=ob.MIX_ROTATE, mix_freq_hz=-250e3, mix_fs=1.2e6, fir=ob.FIR_IQ, taps=t2, decim=m, demod=ob.DEMOD_SSB, fs_demod=48e3, p0=0.0, audio_bw_hz=2800.0), n, False, n // m, True)
if "am25" in which:
    run("FIR201/25 + AM (3 sections)", ob.Chain(fir=ob.FIR_IQ, taps=t2, decim=m, demod=ob.DEMOD_AM, fs_demod=48e3, audio_bw_hz=2800.0), n, False, n // m, False)
if "fm25" in which:
    run("FIR201/25 + FM + LR4", ob.Chain(fir=ob.FIR_IQ, taps=t2, decim=m, demod=ob.DEMOD_FM, fs_demod=48e3, p0=5e3, audio_bw_hz=2800.0), n, False, n // m, False)
