"""Pin the C oracle: it must agree BIT-FOR-BIT with the independent numpy-f32
restatement (oracle/np_oracle.py) on every hot-path function."""
import numpy as np
import pytest

import oracle as O
from oracle import np_oracle as N


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.complex64:
        a = a.view(np.float32)
    return a.view(np.uint32)


def same(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    fa = a.view(np.float32) if a.dtype == np.complex64 else a
    fb = b.view(np.float32) if b.dtype == np.complex64 else b
    nan = np.isnan(fa)                       # NaN payloads may differ: compare positions
    assert np.array_equal(nan, np.isnan(fb))
    assert np.array_equal(fa.view(np.uint32)[~nan], fb.view(np.uint32)[~nan])


def iq(n, seed, scale=0.5):
    r = np.random.default_rng(seed)
    return (scale * (r.standard_normal(n) + 1j * r.standard_normal(n))).astype(np.complex64)


@pytest.mark.parametrize("args", [(2.4e6, 100e3, 38400.0), (96e3, 10800.0, 2400.0),
                                  (48e3, 3e3, 100.0), (100e6, 450e3, 97800.0), (8.192e6, 3.5e3, 16e3)])
def test_fir_lowpass_design(args):
    same(O.fir_lowpass_taps(*args), N.fir_lowpass_taps(*args))


@pytest.mark.parametrize("args", [(3, 0.2, 60.0), (16, 0.2, 60.0), (101, 0.2, 60.0),
                                  (81, 0.2, 60.0), (201, 0.01, 60.0), (31, 0.1, 30.0), (33, 0.3, 10.0)])
def test_kaiser_design(args):
    same(O.kaiser_lowpass_taps(*args), N.kaiser_lowpass_taps(*args))


@pytest.mark.parametrize("fs,fc", [(48e3, 4500.0), (300e3, 13500.0), (64e3, 2700.0), (48e3, 2520.0)])
def test_biquad_design(fs, fc):
    same(O.lp_biquad_coeffs(fs, fc), N.lp_biquad_coeffs(fs, fc))
    assert np.float32(O.dc_pole(fs, 2.0)) == N.dc_pole(fs, 2.0)
    assert np.float32(O.cw_alpha(fs, 200.0)) == N.cw_alpha(fs, 200.0)


def test_atan2_approx():
    r = np.random.default_rng(7)
    y = r.standard_normal(4096).astype(np.float32)
    x = r.standard_normal(4096).astype(np.float32)
    y[:8] = [0, 0, 1, -1, 0, -0.0, 1e-30, -1e-30]
    x[:8] = [0, 1, 0, 0, -1, -1, -1e-30, 1e-30]
    same(O.atan2_approx(y, x), N.atan2_approx(y, x))
    # Parity trap: the reference's doc comment claims ~5e-4 rad, but the polynomial as
    # written (util.rs:313) is off by up to 0.178 rad at 45 degrees.  Parity means THIS
    # function, so a kernel using atan2f would fail the 1e-4 tolerance.
    th = np.linspace(-3.0, 3.0, 2001)
    ys, xs = np.sin(th).astype(np.float32), np.cos(th).astype(np.float32)
    err = np.abs(O.atan2_approx(ys, xs).astype(np.float64) - np.arctan2(ys.astype(np.float64), xs.astype(np.float64)))
    assert 0.17 < err.max() < 0.18


@pytest.mark.parametrize("f,fs", [(100e3, 2.4e6), (-250e3, 1.2e6), (1500.0, 48e3), (0.0, 48e3)])
def test_rotator(f, fs):
    n = 2500          # crosses two renormalisation points
    p = O.Rotator(f, fs).phasors(n)
    same(p, N.rotator_phasors(f, fs, n))
    x = iq(n, 3)
    same(O.Rotator(f, fs).rotate_block(x), N.rotate_block(x, p))
    same(O.Rotator(f, fs).mix_usb_block(x), N.mix_usb(x, p))
    same(O.Nco(f, fs).mix(x), N.nco_mix(x, p))


@pytest.mark.parametrize("m", [1, 4, 8])
def test_fir_decimator(m):
    taps = O.fir_lowpass_taps(96e3, 10800.0, 2400.0)
    x = iq(1003, 11)
    same(O.FirDecimator(taps=taps, m=m).run(x), N.fir_decimator_process(taps, m, x))
    # non-zero end taps exercise the "newest sample pairs with taps[L-1]" quirk
    t2 = np.random.default_rng(5).standard_normal(17).astype(np.float32)
    same(O.FirDecimator(taps=t2, m=m).run(x), N.fir_decimator_process(t2, m, x))
    same(O.FirLowpass(taps=t2).run(x.real.copy()), N.fir_lowpass_process(t2, x.real))


def test_fir_iq():
    taps = O.kaiser_lowpass_taps(31, 0.2, 60.0)
    x = iq(700, 12)
    same(O.FirLowpassIq(taps=taps).run(x), N.fir_iq_process(taps, x))


def test_iir_blocks():
    x = np.random.default_rng(9).standard_normal(3000).astype(np.float32)
    c = O.lp_biquad_coeffs(48e3, 4500.0)
    same(O.Biquad(*c).run(x), N.biquad_run(c, x))
    same(O.LpCascade(48e3, 4500.0).run(x), N.lp_cascade_run(48e3, 4500.0, x))
    same(O.LpDcCascade(48e3, 4500.0, 2.0).run(x), N.lp_dc_run(48e3, 4500.0, 2.0, x))
    xp = np.abs(x) + np.float32(0.1)
    same(O.LpDcCascade(48e3, 4500.0, 2.0, map_sqrt=True).run(xp), N.lp_dc_run(48e3, 4500.0, 2.0, xp, True))
    same(O.DcBlocker(48e3, 2.0).run(x), N.dc_run(N.dc_pole(48e3, 2.0), x))


def test_demods():
    n = 2600
    x = iq(n, 21)
    same(O.FmQuadratureDemod(48e3, 2500.0, 5e3).run(x), N.fm_demod(48e3, 2500.0, 5e3, x))
    same(O.FmQuadratureDemod(300e3, 25e3, 15e3).with_translate(100e3).run(x),
         N.fm_demod(300e3, 25e3, 15e3, x, translate_hz=100e3))
    same(O.PmQuadratureDemod(48e3, 1.0, 5e3).run(x), N.pm_demod(48e3, 1.0, 5e3, x))
    same(O.AmEnvelopeDemod(48e3, 5e3).run(x), N.am_demod(48e3, 5e3, x))
    same(O.AmEnvelopeDemod(48e3, 5e3, abs_approx=True).run(x),
         N.am_demod(48e3, 5e3, x, abs_approx=(0.9482, 0.3920)))
    same(O.SsbProductDemod(48e3, 1500.0, 2800.0).run(x), N.ssb_demod(48e3, 1500.0, 2800.0, x))
    cw = O.CwEnvelopeDemod(48e3, 700.0, 200.0); cw.set_gain(2.5)
    same(cw.run(x), N.cw_demod(48e3, 200.0, x, gain=2.5))


def test_streaming_state_is_chunk_invariant():
    """State persists across process() calls (python/tests/test_unit.py:294-303); the only
    chunk-dependent block is the decimator, whose phase restarts per call (decim.rs:65-71)."""
    x = iq(4096, 33)
    for mk in (lambda: O.FmQuadratureDemod(48e3, 2500.0, 5e3).with_translate(1e3),
               lambda: O.AmEnvelopeDemod(48e3, 5e3), lambda: O.SsbProductDemod(48e3, 1500.0, 2800.0),
               lambda: O.CwEnvelopeDemod(48e3, 700.0, 200.0), lambda: O.PmQuadratureDemod(48e3, 1.0, 5e3),
               lambda: O.FirLowpassIq(31, 0.2, 60.0), lambda: O.Rotator(1234.0, 48e3)):
        one = mk().run(x)
        b = mk()
        parts = [b.run(x[:1000]), b.run(x[1000:1001]), b.run(x[1001:3000]), b.run(x[3000:])]
        same(one, np.concatenate(parts))
    d1 = O.FirDecimator(96e3, 4, 10800.0, 2400.0).run(x)
    d = O.FirDecimator(96e3, 4, 10800.0, 2400.0)
    same(d1, np.concatenate([d.run(x[:1000]), d.run(x[1000:])]))       # multiples of m: invariant
    d = O.FirDecimator(96e3, 4, 10800.0, 2400.0)
    a, b2 = d.run(x[:1001]), d.run(x[1001:])
    assert a.size == 251 and b2.size == 774                              # ceil(n/m) each call
    dfull = O.FirDecimator(96e3, 1, 10800.0, 2400.0).run(x)              # same taps, m=1
    same(b2, dfull[1001::4])                                             # phase restarted at sample 1001


@pytest.mark.parametrize("L,m,n", [(63, 8, 20_001), (17, 1, 300), (201, 25, 30_000), (1023, 32, 40_000), (513, 128, 70_001)])
def test_kept_output_accelerators_equal_the_loop_for_loop_blocks(L, m, n):
    """oo_fir_decim_kept / oo_fir_iq_kept (what the full-size GPU parity tests use as the reference FIR) are
    bit-identical to FirDecimator::process / FirLowpassIq::push + stride pick run loop for loop."""
    r = np.random.default_rng(L * 7 + m)
    taps = r.standard_normal(L).astype(np.float32) / L         # non-zero end taps: both pairing rules exercised
    x = iq(n, 40 + m)
    same(O.fir_decim_kept(taps, m, x, threads=3), O.FirDecimator(taps=taps, m=m).run(x))
    same(O.fir_iq_kept(taps, m, x, threads=3), O.FirLowpassIq(taps=taps).run(x)[::m])
    # negative zero and denormal inputs go through the same adds
    x2 = x.copy(); x2[::7] = np.complex64(complex(-0.0, 1e-42))
    same(O.fir_decim_kept(taps, m, x2), O.FirDecimator(taps=taps, m=m).run(x2))


@pytest.mark.parametrize("iq", [False, True])
def test_agc_c_and_numpy_restatements_agree_bit_for_bit(iq):   # src/dsp/agc.rs:47-74, :121-149
    r = np.random.default_rng(0xA6C)
    n = 6000
    lvl = np.where((np.arange(n) // 1500) % 2 == 0, 0.05, 0.9)
    if iq:
        x = (lvl * (r.standard_normal(n) + 1j * r.standard_normal(n))).astype(np.complex64)
        blk = O.AgcRmsIq(48e3, 2.0, 40.0, 0.25)
    else:
        x = (lvl * r.standard_normal(n)).astype(np.float32)
        blk = O.AgcRms(48e3, 2.0, 40.0, 0.25)
    a = np.concatenate([blk.run(x[:2500]), blk.run(x[2500:])])            # env persists across calls
    b, env = N.agc_rms(48e3, 2.0, 40.0, 0.25, x, iq=iq)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.float32(blk.env) == env
