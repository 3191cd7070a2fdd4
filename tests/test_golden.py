"""Committed fixtures (tests/golden/oracle_vectors.npz, made by tests/golden/make_golden.py).

CPU half: the oracle built on THIS machine must reproduce the fixtures bit for bit (catches compiler /
libm drift between the container that generated them and the box that runs the GPU tests).
GPU half: the CUDA path through the C ABI against the fixtures alone -- no oracle execution involved --
at the north_star tolerance (max abs error <= 1e-4 of full scale, SNR >= 90 dB; counts bit-exact)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from signals import assert_parity, bit_equal  # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))


def test_oracle_reproduces_the_fixtures_bit_for_bit():
    import make_golden
    fresh = make_golden.build()
    assert set(fresh) == set(G.files)
    for k in G.files:
        a, b = np.asarray(fresh[k]), G[k]
        if a.dtype.kind in "fc" and a.dtype != np.float64:
            assert bit_equal(a, b), f"fixture {k} drifted"
        else:
            np.testing.assert_array_equal(a, b, err_msg=k)


def test_host_design_helpers_reproduce_the_fixtures():
    import orion_b200 as ob
    assert bit_equal(ob.fir_lowpass_design(2.4e6, 100e3, 38400.0), G["fir_taps_c1"])
    assert bit_equal(ob.fir_lowpass_design(96e3, 10800.0, 2400.0), G["fir_taps_unit"])
    assert bit_equal(ob.kaiser_lowpass_taps(81, 0.2, 60.0), G["kaiser_81"])
    assert bit_equal(ob.kaiser_lowpass_taps(201, 0.01, 60.0), G["kaiser_201"])
    assert bit_equal(ob.lp_biquad_design(300e3, 13500.0), G["lr4_300k"])
    assert bit_equal(ob.lp_biquad_design(48e3, 4500.0), G["lr4_48k"])
    s = G["scalars"]
    assert np.float32(ob.dc_pole(48e3, 2.0)) == np.float32(s[0]) and np.float32(ob.cw_alpha(48e3, 300.0)) == np.float32(s[1])
    assert np.float32(ob.kaiser_transition_norm(101, 60.0)) == np.float32(s[2]) and ob.kaiser_num_taps(0.02, 60.0) == int(s[3])


@pytest.mark.gpu
def test_gpu_c1_chain_against_fixture():
    import orion_b200 as ob
    x = G["c1_in"]
    chain = ob.Chain(fir=ob.FIR_DECIM, taps=G["fir_taps_c1"], decim=8, demod=ob.DEMOD_FM, fs_demod=3e5, p0=25e3,
                     audio_bw_hz=15e3, translate_hz=100e3)
    out = chain.run(x)
    assert out.size == G["c1_fm"].size == 6000
    assert_parity(out, G["c1_fm"], what="C1 chain vs fixture")
    # the same stream in the fixture's ragged chunks (decimation phase restarts per call)
    c2 = ob.Chain(fir=ob.FIR_DECIM, taps=G["fir_taps_c1"], decim=8, demod=ob.DEMOD_FM, fs_demod=3e5, p0=25e3,
                  audio_bw_hz=15e3, translate_hz=100e3)
    outs, pos = [], 0
    for c in G["c1_chunks"]:
        outs.append(c2.run(x[pos:pos + int(c)]))
        pos += int(c)
    got = np.concatenate(outs)
    assert got.size == G["c1_fm_chunked"].size
    assert_parity(got, G["c1_fm_chunked"], what="C1 chain, ragged chunks vs fixture")
    # decimator alone, reference accumulation order: bit-exact
    d = ob.FirDecimator(2.4e6, 8, 100e3, 38400.0)
    d.set_option(ob.OPT_FIR_GLOBAL, 1)
    assert bit_equal(d.run(x), G["c1_decim"])


@pytest.mark.gpu
def test_gpu_blocks_against_fixtures():
    import orion_b200 as ob
    xc, xf = G["noise_c64"], G["noise_f32"]
    assert_parity(ob.Rotator(100e3, 2.4e6).run(xc), G["rot_100k"], what="rotator")
    assert_parity(ob.RotatorUsb(1.5e3, 48e3).run(xc), G["usb_1500"], what="usb")
    assert_parity(ob.NcoMixer(-250e3, 1.2e6).run(xc), G["nco_m250k"], what="nco")
    assert_parity(ob.FirLowpassIq(81, 0.1, 60.0).run(xc), G["fir_iq_81"], what="fir iq")
    assert_parity(ob.LpCascade(48e3, 4.5e3).run(xf), G["lp_cascade"], what="lp cascade")
    assert_parity(ob.LpDcCascade(48e3, 2520.0, 2.0).run(xf), G["lp_dc"], what="lp dc")
    assert_parity(ob.DcBlocker(48e3, 2.0).run(xf), G["dc_blocker"], what="dc blocker")
    assert_parity(ob.PmQuadratureDemod(48e3, 0.8, 5e3).run(G["pm_in"]), G["pm_out"], what="pm")
    assert_parity(ob.AmEnvelopeDemod(48e3, 5e3).run(G["am_in"]), G["am_out"], what="am")
    assert_parity(ob.AmEnvelopeDemod(48e3, 5e3).with_abs_approx(0.9482, 0.3920).run(G["am_in"]), G["am_abs_out"], what="am abs")
    assert_parity(ob.SsbProductDemod(48e3, 1.5e3, 2800.0).run(G["ssb_in"]), G["ssb_out"], what="ssb")
    assert_parity(ob.CwEnvelopeDemod(48e3, 700.0, 300.0).run(G["cw_in"]), G["cw_out"], what="cw")
