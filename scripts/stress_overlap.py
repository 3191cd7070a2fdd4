"""Stress for the overlapped-launch hand-over at the threshold (every warp claims its tile at kernel start): repeats the
five-call run of tests/test_gpu_parity.py::test_overlapped_calls_near_the_threshold_with_small_grids many times, with and
without a competing kernel stream, and reports where a mismatch against the serialised run sits."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "orion-sdr_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import orion_b200 as ob
from signals import fm_iq

fs, m, calls = 2.4e6, 8, 5
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
noise_stream = torch.cuda.Stream()
A = torch.randn(4096, 4096, device="cuda")
for ntiles, grid in [(1030, 148), (1024, 64), (1500, 100), (2500, 40), (2368, 148), (1100, 148)]:
    n_call = ntiles * 256 * m - 8 * 77
    x = fm_iq(calls * n_call, fs)
    xd = torch.from_numpy(x).cuda()
    n_out = -(-n_call // m)
    taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
    def run(no_overlap, noise):
        os.environ["ORION_B200_GRID"] = str(grid)
        if no_overlap: os.environ["ORION_B200_NO_OVERLAP"] = "1"
        try:
            ch = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3, audio_bw_hz=15e3, translate_hz=100e3)
            yd = torch.zeros(calls * n_out, dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            if noise:
                with torch.cuda.stream(noise_stream):
                    for _ in range(3): (A @ A)
            for c in range(calls):
                ch.process_dev(xd.data_ptr() + c * n_call * 8, n_call, yd.data_ptr() + c * n_out * 4, n_out)
            ch.synchronize()
            torch.cuda.synchronize()
            return yd.cpu().numpy()
        finally:
            os.environ.pop("ORION_B200_GRID", None); os.environ.pop("ORION_B200_NO_OVERLAP", None)
    plain = run(True, False)
    for it in range(iters):
        try:
            over = run(False, it % 2 == 1)
        except Exception as e:
            print(f"ntiles {ntiles} grid {grid} iter {it}: EXCEPTION {e}"); bad += 1; continue
        if not np.array_equal(over.view(np.uint32), plain.view(np.uint32)):
            d = np.flatnonzero(over.view(np.uint32) != plain.view(np.uint32))
            print(f"ntiles {ntiles} grid {grid} iter {it} noise {it % 2}: {d.size} outputs differ, first {d[0]} (call {d[0] // n_out}, tile {(d[0] % n_out) // 256}), last {d[-1]} (call {d[-1] // n_out}, tile {(d[-1] % n_out) // 256}), max abs diff {np.abs(over[d] - plain[d]).max():.3e}")
            bad += 1
    print(f"ntiles {ntiles} grid {grid}: done", flush=True)
print("mismatching runs:", bad)
