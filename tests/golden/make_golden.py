#!/usr/bin/env python
"""Generates tests/golden/oracle_vectors.npz: outputs of the CPU oracle on small seeded inputs.

The reference (Rust) cannot be built or imported in this environment and holds no golden vectors of its
own for this path, so these fixtures freeze the ORACLE (which is pinned to the reference by the transcribed
behavioural tests and the independent numpy restatement).  They catch drift of the oracle itself -- a
different gcc / glibc libm on another box -- and give the GPU parity tests a target that does not depend on
executing the oracle at all.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
from signals import am_iq, cw_iq, fm_iq, noise_c64, noise_f32, pm_iq, ssb_iq  # noqa: E402


def build():
    g = {}
    # design-time
    g["fir_taps_c1"] = oracle.fir_lowpass_taps(2.4e6, 100e3, 38400.0)
    g["fir_taps_unit"] = oracle.fir_lowpass_taps(96e3, 10800.0, 2400.0)
    g["kaiser_81"] = oracle.kaiser_lowpass_taps(81, 0.2, 60.0)
    g["kaiser_201"] = oracle.kaiser_lowpass_taps(201, 0.01, 60.0)
    g["lr4_300k"] = oracle.lp_biquad_coeffs(300e3, 13500.0)
    g["lr4_48k"] = oracle.lp_biquad_coeffs(48e3, 4500.0)
    g["scalars"] = np.array([oracle.dc_pole(48e3, 2.0), oracle.cw_alpha(48e3, 300.0),
                             oracle.kaiser_transition_norm(101, 60.0), oracle.kaiser_num_taps(0.02, 60.0)], np.float64)
    # C1 chain on 48 000 samples: FirDecimator -> FM (translate) -- one-shot and in ragged chunks
    fs, m, n = 2.4e6, 8, 48_000
    x = fm_iq(n, fs)
    dec = oracle.FirDecimator(fs, m, 100e3, 38400.0)
    mid = dec.run(x)
    g["c1_in"], g["c1_decim"] = x, mid
    g["c1_fm"] = oracle.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3).run(mid)
    chunks = [1, 7, 8, 9, 1000, 4097, 13, 20_000, 0, 5, 22_860]
    dec2, fm2, outs, pos = oracle.FirDecimator(fs, m, 100e3, 38400.0), oracle.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3), [], 0
    for c in chunks:
        outs.append(fm2.run(dec2.run(x[pos:pos + c])))
        pos += c
    g["c1_chunks"], g["c1_fm_chunked"] = np.array(chunks), np.concatenate(outs)
    # stand-alone blocks
    xc = noise_c64(8192, seed=11)
    xf = noise_f32(8192, seed=12)
    g["noise_c64"], g["noise_f32"] = xc, xf
    g["rot_100k"] = oracle.Rotator(100e3, 2.4e6).rotate_block(xc)
    g["usb_1500"] = oracle.Rotator(1.5e3, 48e3).mix_usb_block(xc)
    g["nco_m250k"] = oracle.Nco(-250e3, 1.2e6).mix(xc)
    g["fir_iq_81"] = oracle.FirLowpassIq(81, 0.1, 60.0).run(xc)
    g["lp_cascade"] = oracle.LpCascade(48e3, 4.5e3).run(xf)
    g["lp_dc"] = oracle.LpDcCascade(48e3, 2520.0, 2.0).run(xf)
    g["dc_blocker"] = oracle.DcBlocker(48e3, 2.0).run(xf)
    n2 = 16_384
    g["pm_in"] = pm_iq(n2, 48e3)
    g["pm_out"] = oracle.PmQuadratureDemod(48e3, 0.8, 5e3).run(g["pm_in"])
    g["am_in"] = am_iq(n2, 48e3)
    g["am_out"] = oracle.AmEnvelopeDemod(48e3, 5e3).run(g["am_in"])
    g["am_abs_out"] = oracle.AmEnvelopeDemod(48e3, 5e3, abs_approx=True).run(g["am_in"])
    g["ssb_in"] = ssb_iq(n2, 48e3, f_bfo=1.5e3)
    g["ssb_out"] = oracle.SsbProductDemod(48e3, 1.5e3, 2800.0).run(g["ssb_in"])
    g["cw_in"] = cw_iq(n2, 48e3)
    g["cw_out"] = oracle.CwEnvelopeDemod(48e3, 700.0, 300.0).run(g["cw_in"])
    # atan2_approx on a fixed grid incl. axes, signed zeros and the wrap
    yy, xx = np.meshgrid(np.linspace(-1, 1, 33, dtype=np.float32), np.linspace(-1, 1, 33, dtype=np.float32))
    g["atan2_y"], g["atan2_x"] = yy.ravel(), xx.ravel()
    g["atan2_out"] = oracle.atan2_approx(g["atan2_y"], g["atan2_x"])
    return g


if __name__ == "__main__":
    g = build()
    path = os.path.join(HERE, "oracle_vectors.npz")
    np.savez_compressed(path, **g)
    print(path, os.path.getsize(path) // 1024, "KiB,", len(g), "arrays")
