import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "orion-sdr_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import orion_b200 as ob
taps = ob.fir_lowpass_design(2.4e6, 100e3, 38400.0)
for n in [1_000_000, 2_000_000, 4_000_000, 8_000_000, 16_000_000, 24_000_000]:
    x = (torch.randn(2 * n, device="cuda") * 0.1)
    x[0::2] += 0.8                       # carrier: keeps the LR4 of |z|^2 positive
    y = torch.empty(n // 8, device="cuda")
    blk = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=8, demod=ob.DEMOD_AM, fs_demod=3e5, audio_bw_hz=15e3)
    st = torch.cuda.Stream(); blk.set_stream(st.cuda_stream)
    try:
        t0 = time.perf_counter()
        blk.process_dev(x.data_ptr(), n, y.data_ptr(), n // 8); blk.synchronize()
        t1 = time.perf_counter()
        blk.process_dev(x.data_ptr(), n, y.data_ptr(), n // 8); blk.synchronize()
        t2 = time.perf_counter()
        print(n, "ok first %.1f ms second %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), "nan" if torch.isnan(y).any().item() else "finite", flush=True)
    except Exception as e:
        print(n, "FAIL", str(e)[:120], flush=True)
        break
