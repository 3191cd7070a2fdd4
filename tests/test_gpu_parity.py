"""GPU parity tests: every block of the C ABI (through the ctypes binding, i.e. through
liborion_b200.so) against the CPU oracle on identical seeded input.

Bar (BASELINE.json north_star): WorkReport, output count and decimation phase bit-exact;
values within max|gpu-ref| <= 1e-4 * max|ref| and SNR >= 90 dB.  With
ORION_B200_OPT_FIR_GLOBAL the FIR follows the reference's accumulation order and the result
must be bit-identical."""
import numpy as np
import pytest

import oracle
import orion_b200 as ob
from signals import (am_iq, assert_parity, bit_equal, cw_iq, fm_iq, noise_c64, noise_f32, parity, pm_iq, ssb_iq)

pytestmark = pytest.mark.gpu


def run_pair(gpu, ref, x, out_dtype, cap=None, ratio=1):
    """One process() call on both sides with equally sized outputs."""
    n_cap = cap if cap is not None else -(-x.size // ratio)
    og = np.zeros(n_cap, out_dtype)
    orf = np.zeros(n_cap, out_dtype)
    wg = gpu.process(x, og)
    wr = ref.process(x, orf)
    assert tuple(wg) == tuple(wr), (wg, wr)
    return og[:wg.out_written], orf[:wr.out_written]


def stream_pair(gpu, ref, x, out_dtype, chunks, ratio=1):
    """Feed the same ragged chunk sequence to both sides (decimation phase restarts per call)."""
    og, orf, pos = [], [], 0
    for c in chunks:
        a, b = run_pair(gpu, ref, x[pos:pos + c], out_dtype, ratio=ratio)
        og.append(a)
        orf.append(b)
        pos += c
    return np.concatenate(og), np.concatenate(orf)


# ---- FIR decimator ---------------------------------------------------------------------------------
@pytest.mark.parametrize("fs,m,cut,trans,n", [(96e3, 4, 10e3, 2400.0, 4096), (2.4e6, 8, 100e3, 38400.0, 200_003),
                                              (48e3, 2, 5e3, 1500.0, 30_001), (1e6, 16, 20e3, 30e3, 70_000),
                                              (1e6, 5, 50e3, 25e3, 41_234), (100e6, 32, 450e3, 97800.0, 150_000)])
def test_fir_decimator(fs, m, cut, trans, n):
    x = noise_c64(n, seed=n)
    g, r = ob.FirDecimator(fs, m, cut, trans), oracle.FirDecimator(fs, m, cut, trans)
    a, b = run_pair(g, r, x, np.complex64, ratio=m)
    assert a.size == -(-n // m)
    assert_parity(a, b, what="staged")
    g2 = ob.FirDecimator(fs, m, cut, trans)
    g2.set_option(ob.OPT_FIR_GLOBAL, 1)
    a2, _ = run_pair(g2, oracle.FirDecimator(fs, m, cut, trans), x, np.complex64, ratio=m)
    assert bit_equal(a2, b), "reference-order FIR must be bit-identical"


def test_fir_decimator_ragged_chunks_and_short_output():
    fs, m = 2.4e6, 8
    x = noise_c64(60_000, seed=7)
    chunks = [1, 7, 8, 9, 1000, 4097, 13, 20_000, 0, 5, 34_860]
    assert sum(chunks) == x.size
    g, r = ob.FirDecimator(fs, m, 100e3, 38400.0), oracle.FirDecimator(fs, m, 100e3, 38400.0)
    a, b = stream_pair(g, r, x, np.complex64, chunks, ratio=m)
    assert_parity(a, b, what="chunked")
    # output shorter than ceil(n/m): everything is consumed, the output is truncated (decim.rs:66-75)
    g, r = ob.FirDecimator(fs, m, 100e3, 38400.0), oracle.FirDecimator(fs, m, 100e3, 38400.0)
    a, b = run_pair(g, r, x[:10_000], np.complex64, cap=100)
    assert a.size == 100
    assert_parity(a, b)
    a, b = run_pair(g, r, x[10_000:30_000], np.complex64, ratio=m)   # state must have advanced over all input
    assert_parity(a, b, what="after truncated call")


def test_fir_decimator_tma_and_serial_agree():
    fs, m, n = 2.4e6, 8, 300_000
    x = noise_c64(n, seed=11)
    outs = []
    for tma, serial in [(1, 0), (0, 0), (1, 1)]:
        g = ob.FirDecimator(fs, m, 100e3, 38400.0)
        g.set_option(ob.OPT_USE_TMA, tma)
        g.set_option(ob.OPT_SERIAL_TILES, serial)
        outs.append(g.run(x))
    assert bit_equal(outs[0], outs[1]) and bit_equal(outs[0], outs[2])


# ---- FirLowpassIq ------------------------------------------------------------------------------------
@pytest.mark.parametrize("ntaps,fc,a,n", [(81, 0.1, 60.0, 50_001), (201, 0.01, 60.0, 120_000), (5, 0.3, 30.0, 999)])
def test_fir_lowpass_iq(ntaps, fc, a, n):
    x = noise_c64(n, seed=ntaps)
    g, r = ob.FirLowpassIq(ntaps, fc, a), oracle.FirLowpassIq(ntaps, fc, a)
    ya, yb = stream_pair(g, r, x, np.complex64, [n // 3, 1, n - n // 3 - 1])
    assert_parity(ya, yb)
    g2 = ob.FirLowpassIq(ntaps, fc, a)
    g2.set_option(ob.OPT_FIR_GLOBAL, 1)
    assert bit_equal(g2.run(x), oracle.FirLowpassIq(ntaps, fc, a).run(x))
    # filter_aligned == streamed output advanced by the group delay (fir.rs:260-276)
    io_g, io_r = x[:20_000].copy(), x[:20_000].copy()
    ob.FirLowpassIq(ntaps, fc, a).filter_aligned(io_g)
    oracle.FirLowpassIq(ntaps, fc, a).filter_aligned(io_r)
    assert_parity(io_g, io_r, what="filter_aligned")


def test_fir_identity_and_from_taps():
    x = noise_c64(5000, seed=3)
    g = ob.FirLowpassIq.from_taps([])                                   # fir.rs:193-196
    assert bit_equal(g.run(x), x)
    taps = np.random.default_rng(5).standard_normal(47).astype(np.float32)
    a = ob.FirDecimator.from_taps(taps, 6).run(x)
    b = oracle.FirDecimator(taps=taps, m=6).run(x)
    assert_parity(a, b)


# ---- oscillators -------------------------------------------------------------------------------------
@pytest.mark.parametrize("f,fs,n", [(100e3, 2.4e6, 16_384), (1.5e3, 48e3, 200_000), (-250e3, 1.2e6, 65_536)])
def test_rotator_nco_usb(f, fs, n):
    # default oscillator mode for these blocks: the reference recurrence replayed exactly -> bit-identical outputs
    x = noise_c64(n, seed=int(abs(f)))
    a, b = stream_pair(ob.Rotator(f, fs), oracle.Rotator(f, fs), x, np.complex64, [n // 2 + 3, n - n // 2 - 3])
    assert bit_equal(a, b), "rotate_block"
    a = ob.NcoMixer(f, fs).run(x)
    b = oracle.Nco(f, fs).mix(x)
    assert bit_equal(a, b), "mix_with_nco"
    a = ob.RotatorUsb(f, fs).run(x)
    b = oracle.Rotator(f, fs).mix_usb_block(x)
    assert bit_equal(a, b), "mix_usb_block"
    # the closed-form option (absolute-index phase, north_star's formulation) stays within tolerance on short streams
    g = ob.Rotator(f, fs)
    g.set_option(ob.OPT_EXACT_NCO, 0)
    assert_parity(g.run(x), oracle.Rotator(f, fs).rotate_block(x), what="rotate_block, closed-form phase")
    assert g.exact_host_ms == 0.0


def test_mixer_fir_chain_set_freq_and_reset_phase_mid_stream():
    """Rotator -> FirLowpassIq streaming with set_freq and reset_phase between calls: the FIR history holds samples
    that were mixed with the OLD phasors (the reference mixes before its delay line); the exact-replay oscillator keeps
    the phasors it applied to the history items."""
    fs, n = 48e3, 30_000
    x = noise_c64(3 * n, seed=77)
    taps = ob.kaiser_lowpass_taps(81, 0.1, 60.0)
    g = ob.Chain(mix=ob.MIX_ROTATE, mix_freq_hz=1.5e3, mix_fs=fs, fir=ob.FIR_IQ, taps=taps, decim=1)
    rot, fir = oracle.Rotator(1.5e3, fs), oracle.FirLowpassIq(taps=oracle.kaiser_lowpass_taps(81, 0.1, 60.0))
    outs, refs = [], []
    for i in range(3):
        seg = x[i * n:(i + 1) * n]
        outs.append(g.run(seg))
        refs.append(fir.run(rot.rotate_block(seg)))
        if i == 0:
            lib = ob.lib()
            assert lib.orion_b200_oscillator_set_freq(g._h, -4.2e3, fs) == 0
            rot.set_freq(-4.2e3, fs)
        if i == 1:
            assert ob.lib().orion_b200_oscillator_reset_phase(g._h) == 0
            rot.reset_phase()
    assert_parity(np.concatenate(outs), np.concatenate(refs), tol=2e-6, snr_db=120.0, what="mixer -> FIR with set_freq / reset_phase")


def test_chain_am_with_abs_approx_after_create_reselects_the_kernel():
    """with_abs_approx on a fused FIR/8 + AM chain must drop the power front AND the sqrt (the specialised
    instance bakes the demodulator kind in)."""
    fs, m, n = 384e3, 8, 400_000
    x = am_iq(n, fs)
    taps = ob.fir_lowpass_design(fs, 10e3, 6144.0)
    g = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_AM, fs_demod=48e3, audio_bw_hz=5e3)
    assert ob.lib().orion_b200_am_demod_with_abs_approx(g._h, 0.9482, 0.3920) == 0
    ref = oracle.AmEnvelopeDemod(48e3, 5e3, abs_approx=True).run(oracle.FirDecimator(fs, m, 10e3, 6144.0).run(x))
    assert_parity(g.run(x), ref, what="Chain(AM).with_abs_approx")


def test_decimator_with_m_1_consumes_all_input_when_output_is_short():
    """decim.rs:44-76 for m = 1: in_read = n whatever out.len() is; the filter state advances over all n."""
    taps = np.random.default_rng(3).standard_normal(31).astype(np.float32)
    x = noise_c64(5000, seed=5)
    g, r = ob.FirDecimator.from_taps(taps, 1), oracle.FirDecimator(taps=taps, m=1)
    og, orf = np.zeros(3000, np.complex64), np.zeros(3000, np.complex64)
    wg, wr = g.process(x, og), r.process(x, orf)
    assert tuple(wg) == tuple(wr) == (5000, 3000)
    assert_parity(og, orf)
    a, b = g.run(x[:2000]), r.run(x[:2000])                # continues from the state after ALL 5000 samples
    assert_parity(a, b)


def test_rotator_set_freq_keeps_phase():
    x = noise_c64(8192, seed=21)
    g, r = ob.Rotator(1.5e3, 48e3), oracle.Rotator(1.5e3, 48e3)
    a1, b1 = run_pair(g, r, x[:4000], np.complex64)
    g.set_freq(-3.1e3, 48e3)
    r.set_freq(-3.1e3, 48e3)
    a2, b2 = run_pair(g, r, x[4000:], np.complex64)
    assert bit_equal(np.concatenate([a1, a2]), np.concatenate([b1, b2]))
    g.reset_phase()
    r.reset_phase()
    a3, b3 = run_pair(g, r, x[:1000], np.complex64)
    assert bit_equal(a3, b3)


# ---- recursive sections -------------------------------------------------------------------------------
def test_iir_blocks():
    n = 400_000
    x = noise_f32(n, seed=31)
    chunks = [1, 5, 100_000, 7, 299_987]
    c = oracle.lp_biquad_coeffs(48e3, 3e3)
    pairs = [
        (ob.Biquad(*c), oracle.Biquad(*c)),
        (ob.LpCascade(48e3, 4.5e3), oracle.LpCascade(48e3, 4.5e3)),
        (ob.LpDcCascade(48e3, 4.5e3, 2.0), oracle.LpDcCascade(48e3, 4.5e3, 2.0)),
        (ob.DcBlocker(48e3, 2.0), oracle.DcBlocker(48e3, 2.0)),
        (ob.DcBlocker(1e6, 0.05), oracle.DcBlocker(1e6, 0.05)),          # pole clamped at 0.9999
    ]
    for g, r in pairs:
        a, b = stream_pair(g, r, x, np.float32, chunks)
        assert_parity(a, b, what=type(g).__name__)
    # process_mapped(x, sqrt) on a positive input (iir.rs:170-186)
    xp = np.abs(x) + 0.5
    a, b = stream_pair(ob.LpDcCascade(48e3, 4.5e3, 2.0, True), oracle.LpDcCascade(48e3, 4.5e3, 2.0, True), xp,
                       np.float32, chunks)
    assert_parity(a, b, what="LpDcCascade mapped")
    # N-section cascade == LpCascade twice
    sos = np.stack([c, c, c, c])
    a = ob.IirCascade(sos).run(x)
    r1, r2 = oracle.LpCascade(48e3, 3e3), oracle.LpCascade(48e3, 3e3)
    assert_parity(a, r2.run(r1.run(x)), what="4-section cascade")


def test_sqrt_of_negative_latches_nan_like_the_reference():
    # am.rs:55 / iir.rs:180: LR4 of |z|^2 can ring below zero; sqrt -> NaN; the DC blocker keeps it forever
    x = np.zeros(6000, np.complex64)
    x[:3000] = 1.0
    a = ob.AmEnvelopeDemod(48e3, 5e3).run(x)
    b = oracle.AmEnvelopeDemod(48e3, 5e3).run(x)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    if np.isnan(b).any():
        first = int(np.argmax(np.isnan(b)))
        assert np.isnan(a[first:]).all()
        assert_parity(a[:first], b[:first])


# ---- demodulators -------------------------------------------------------------------------------------
def _demod_cases():
    n = 300_000
    return [
        ("fm", lambda m: m.FmQuadratureDemod(48e3, 2.5e3, 5e3), lambda: fm_iq(n, 48e3, f_c=0.0, dev=2.5e3)),
        ("fm_translate", lambda m: m.FmQuadratureDemod(300e3, 25e3, 15e3).with_translate(100e3), lambda: fm_iq(n, 300e3)),
        ("pm", lambda m: m.PmQuadratureDemod(48e3, 1.25, 5e3), lambda: pm_iq(n, 48e3)),
        ("am", lambda m: m.AmEnvelopeDemod(48e3, 5e3), lambda: am_iq(n, 48e3)),
        ("am_abs", lambda m: m.AmEnvelopeDemod(48e3, 5e3).with_abs_approx(0.9482, 0.3920), lambda: am_iq(n, 48e3, f_off=50.0)),
        ("ssb", lambda m: m.SsbProductDemod(48e3, 1.5e3, 2.8e3), lambda: ssb_iq(n, 48e3, f_bfo=1.5e3)),
        ("cw", lambda m: m.CwEnvelopeDemod(48e3, 700.0, 50.0), lambda: cw_iq(n, 48e3)),
    ]


@pytest.mark.parametrize("case", _demod_cases(), ids=lambda c: c[0])
def test_demodulators(case):
    name, make, gen = case
    x = gen()
    g, r = make(ob), make(oracle)
    n = x.size
    a, b = stream_pair(g, r, x, np.float32, [4096, 1, 100_003, n - 4096 - 1 - 100_003])
    assert_parity(a, b, what=name)
    # state persists across calls and equals the oracle's
    sg, sr = g.state(), r.state()
    if name in ("fm", "pm"):
        assert np.allclose(sg[0:2], sr[15:17], rtol=0, atol=1e-4 * max(1.0, np.max(np.abs(sr[15:17]))))
    elif name == "fm_translate":
        # `prev` is the translated sample: its absolute phase carries the oscillator's (closed-form
        # vs f32-recurrence) common phase, which the discriminator cancels; the magnitude must agree
        assert abs(np.hypot(*sg[0:2]) - np.hypot(*sr[15:17])) <= 1e-4 * np.hypot(*sr[15:17])


def test_cw_set_gain():
    x = cw_iq(20_000, 48e3)
    g, r = ob.CwEnvelopeDemod(48e3, 700.0, 50.0), oracle.CwEnvelopeDemod(48e3, 700.0, 50.0)
    g.set_gain(2.5)
    r.set_gain(2.5)
    a, b = run_pair(g, r, x, np.float32)
    assert_parity(a, b)


@pytest.mark.parametrize("n", [0, 1, 2, 7, 8, 9, 1023, 1024, 1025])
def test_tiny_and_empty_inputs(n):
    x = fm_iq(max(n, 1), 48e3, f_c=0.0, dev=2.5e3)[:n]
    a, b = run_pair(ob.FmQuadratureDemod(48e3, 2.5e3, 5e3), oracle.FmQuadratureDemod(48e3, 2.5e3, 5e3), x, np.float32)
    assert a.size == n
    assert_parity(a, b)
    xd = noise_c64(n, seed=n + 1)
    a, b = run_pair(ob.FirDecimator(96e3, 4, 10e3, 2400.0), oracle.FirDecimator(96e3, 4, 10e3, 2400.0), xd,
                    np.complex64, ratio=4)
    assert a.size == -(-n // 4)
    assert_parity(a, b)


def test_rate_one_length_rule():
    # n = min(len(in), len(out)) (fm.rs:46,73-76)
    x = fm_iq(5000, 48e3, f_c=0.0, dev=2.5e3)
    g, r = ob.FmQuadratureDemod(48e3, 2.5e3, 5e3), oracle.FmQuadratureDemod(48e3, 2.5e3, 5e3)
    a, b = run_pair(g, r, x, np.float32, cap=3000)
    assert a.size == 3000
    assert_parity(a, b)
    a, b = run_pair(g, r, x[3000:], np.float32, cap=9000)
    assert a.size == 2000
    assert_parity(a, b)


# ---- the fused chain (north-star path) -----------------------------------------------------------------
def _c1_pair(fs=2.4e6, m=8):
    taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
    g = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3,
                 audio_bw_hz=15e3, translate_hz=100e3)
    dec = oracle.FirDecimator(fs, m, 100e3, 38400.0)
    fm = oracle.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3)
    return g, dec, fm, taps


def _ref_chain(dec, fm, x, m):
    mid = np.zeros(-(-x.size // m), np.complex64)
    wr = dec.process(x, mid)
    out = np.zeros(wr.out_written, np.float32)
    fm.process(mid[:wr.out_written], out)
    return out


def test_chain_c1_fir_decim_nco_fm():
    fs, m, n = 2.4e6, 8, 1_200_000                       # 0.5 s of the C1 stream
    x = fm_iq(n, fs)
    g, dec, fm, taps = _c1_pair(fs, m)
    assert taps.size == 63
    ref = _ref_chain(dec, fm, x, m)
    out = np.zeros(ref.size, np.float32)
    wr = g.process(x, out)
    assert tuple(wr) == (n, ref.size)
    e, snr = assert_parity(out, ref, what="C1 chain")
    print(f"C1 chain parity: max err {e:.2e} of full scale, SNR {snr:.1f} dB")
    # start-up transient included; streaming in ragged chunks gives the oracle's chunked result
    g, dec, fm, _ = _c1_pair(fs, m)
    outs, refs, pos = [], [], 0
    for c in [8 * 1000, 8 * 37 + 3, 500_001, n - 8 * 1000 - (8 * 37 + 3) - 500_001]:
        xc = x[pos:pos + c]
        pos += c
        refs.append(_ref_chain(dec, fm, xc, m))
        o = np.zeros(refs[-1].size, np.float32)
        g.process(xc, o)
        outs.append(o)
    assert_parity(np.concatenate(outs), np.concatenate(refs), what="C1 chain chunked")


def test_chain_c1_reference_order_fir_variant():
    fs, m, n = 2.4e6, 8, 160_000
    x = fm_iq(n, fs)
    g, dec, fm, _ = _c1_pair(fs, m)
    g.set_option(ob.OPT_FIR_GLOBAL, 1)
    ref = _ref_chain(dec, fm, x, m)
    out = g.run(x)
    assert_parity(out, ref, tol=2e-6, snr_db=120.0, what="C1 with bit-faithful FIR")


def test_chain_c2_mix_fir_decim_ssb():
    # Rotator(-250k) -> FirLowpassIq(201, 0.01, 60) -> keep every 25th -> SsbProductDemod(48k, 0, 2800)
    fs, m, n = 1.2e6, 25, 2 ** 17
    x = ssb_iq(n, fs)
    taps = ob.kaiser_lowpass_taps(201, 0.01, 60.0)
    g = ob.Chain(mix=ob.MIX_ROTATE, mix_freq_hz=-250e3, mix_fs=fs, fir=ob.FIR_IQ, taps=taps, decim=m,
                 demod=ob.DEMOD_SSB, fs_demod=fs / m, p0=0.0, audio_bw_hz=2800.0)
    rot, fir = oracle.Rotator(-250e3, fs), oracle.FirLowpassIq(201, 0.01, 60.0)
    ssb = oracle.SsbProductDemod(fs / m, 0.0, 2800.0)
    y = fir.run(rot.rotate_block(x))[::m]
    ref = ssb.run(np.ascontiguousarray(y))
    out = g.run(x)
    assert out.size == ref.size == -(-n // m)
    assert_parity(out, ref, what="C2 chain")


def test_chain_c3_fir_decim_am_four_sections():
    fs, m, n = 384e3, 8, 1_000_000
    x = am_iq(n, fs)
    taps = ob.fir_lowpass_design(fs, 10e3, 6144.0)
    extra = np.stack([oracle.lp_biquad_coeffs(48e3, 3e3)] * 2)
    g = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_AM, fs_demod=48e3, audio_bw_hz=5e3,
                 post_sos=extra)
    dec, am, lp = oracle.FirDecimator(fs, m, 10e3, 6144.0), oracle.AmEnvelopeDemod(48e3, 5e3), oracle.LpCascade(48e3, 3e3)
    ref = lp.run(am.run(dec.run(x)))
    out = g.run(x)
    assert_parity(out, ref, what="C3 chain")


# ---- src/core.rs chain wrappers ------------------------------------------------------------------------
def test_chain_wrappers_return_input_len_items():
    x = cw_iq(4096, 48e3)
    a = ob.IqToAudioChain(ob.CwEnvelopeDemod(48e3, 700.0, 50.0)).process(x)
    b = oracle.IqToAudioChain(oracle.CwEnvelopeDemod(48e3, 700.0, 50.0)).process(x)
    assert a.size == b.size == x.size
    assert_parity(a, b)
    # a decimating block behind IqToIqChain: input.len() items come back, tail stale (core.rs:70-77)
    xd = noise_c64(4096, seed=9)
    a = ob.IqToIqChain(ob.FirDecimator(96e3, 4, 10e3, 2400.0)).process(xd)
    b = oracle.IqToIqChain(oracle.FirDecimator(96e3, 4, 10e3, 2400.0)).process(xd)
    assert a.size == b.size == 4096
    assert_parity(a[:1024], b[:1024])
    assert not a[1024:].any() and not b[1024:].any()


# ---- back-to-back long calls: the launches overlap (programmatic dependent launch), the stream must not notice --
def test_overlapped_back_to_back_calls_keep_streaming_state():
    import os
    import torch
    fs, m = 2.4e6, 8
    n_call = 2048 * 256 * m + 8 * 1237            # >= 2048 warp tiles per call: the overlap path is taken
    calls = 3
    x = fm_iq(calls * n_call, fs)
    xd = torch.from_numpy(x).cuda()
    n_out = n_call // m

    def run(no_overlap):
        if no_overlap:
            os.environ["ORION_B200_NO_OVERLAP"] = "1"
        else:
            os.environ.pop("ORION_B200_NO_OVERLAP", None)
        taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
        ch = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3,
                      audio_bw_hz=15e3, translate_hz=100e3)
        yd = torch.zeros(calls * n_out, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        for c in range(calls):                     # enqueued back to back on the block's own stream
            wr = ch.process_dev(xd.data_ptr() + c * n_call * 8, n_call, yd.data_ptr() + c * n_out * 4, n_out)
            assert tuple(wr) == (n_call, n_out)
        ch.synchronize()
        os.environ.pop("ORION_B200_NO_OVERLAP", None)
        return yd.cpu().numpy()

    over, plain = run(False), run(True)
    assert bit_equal(over, plain), "overlapped launches must give exactly the serialised result"
    dec = oracle.FirDecimator(fs, m, 100e3, 38400.0)
    fm = oracle.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3)
    ref = np.concatenate([fm.run(dec.run(x[c * n_call:(c + 1) * n_call])) for c in range(calls)])
    assert_parity(over, ref, what="3 overlapped calls vs oracle streaming")
    # the host-pointer entry point pipelines a long call in chunks (copy in / kernel / copy out overlap): internally a
    # sequence of streaming launches, so it agrees with the one-launch result to the rounding of the carried start states
    taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
    ch = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3,
                  audio_bw_hz=15e3, translate_hz=100e3)
    host = np.concatenate([ch.run(x[c * n_call:(c + 1) * n_call]) for c in range(calls)])
    e, snr = assert_parity(host, over, what="pipelined host call vs device calls")
    assert e <= 5e-6 and snr >= 110.0, (e, snr)
    assert_parity(host, ref, what="pipelined host call vs oracle streaming")
    os.environ["ORION_B200_NO_PIPELINE"] = "1"            # one copy, one launch, one copy: exactly the device-call result
    try:
        ch.reset()
        host1 = np.concatenate([ch.run(x[c * n_call:(c + 1) * n_call]) for c in range(calls)])
    finally:
        os.environ.pop("ORION_B200_NO_PIPELINE", None)
    assert bit_equal(host1, over)


@pytest.mark.parametrize("ntiles,grid", [(1024, 64), (1500, 100), (2500, 40), (1030, 148), (1100, 148), (2368, 148)])
def test_overlapped_calls_near_the_threshold_with_small_grids(ntiles, grid):
    """at the overlap threshold every warp of a small grid claims its tile at kernel start, and the last tile
    of call N can finish before its first tiles have consumed what call N-1 handed over.  Hand-over buffers rotate
    through three and call N+1 waits for every CTA of call N-1 (a done counter), so five overlapped calls must equal
    the serialised run bit for bit whatever the grid."""
    import os
    import torch
    fs, m, calls = 2.4e6, 8, 5
    n_call = ntiles * 256 * m - 8 * 77
    x = fm_iq(calls * n_call, fs)
    xd = torch.from_numpy(x).cuda()
    n_out = -(-n_call // m)

    def run(no_overlap):
        os.environ["ORION_B200_GRID"] = str(grid)
        if no_overlap:
            os.environ["ORION_B200_NO_OVERLAP"] = "1"
        try:
            taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
            ch = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3,
                          audio_bw_hz=15e3, translate_hz=100e3)
            yd = torch.zeros(calls * n_out, dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            for c in range(calls):
                ch.process_dev(xd.data_ptr() + c * n_call * 8, n_call, yd.data_ptr() + c * n_out * 4, n_out)
            ch.synchronize()
            return yd.cpu().numpy()
        finally:
            os.environ.pop("ORION_B200_GRID", None)
            os.environ.pop("ORION_B200_NO_OVERLAP", None)

    over, plain = run(False), run(True)
    assert bit_equal(over, plain)
    # Round 2: 1100 tiles on 69 CTAs failed in ~70 % of the runs (look-back watchdog: call N found call N+2's link records)
    # while 1030 tiles failed once in a few hundred -- CTAs of call N+1 finishing before the last CTA of call N satisfied
    # call N+2's CTA-done target.  The counters are per launch parity now; repeat the run to catch a relapse.
    for _ in range(6):
        assert bit_equal(run(False), plain)
    dec = oracle.FirDecimator(fs, m, 100e3, 38400.0)
    fm = oracle.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3)
    ref = np.concatenate([fm.run(dec.run(x[c * n_call:(c + 1) * n_call])) for c in range(calls)])
    assert_parity(over, ref, what="5 overlapped calls, small grid")


@pytest.mark.parametrize("kind", ["decimator", "lp_cascade", "fm_rate1", "pm_rate1", "rotator", "am_rate1"])
def test_overlapped_calls_other_blocks(kind):
    """Three long back-to-back device calls per block kind: overlapped launches == serialised launches bit for bit,
    and both within tolerance of the oracle fed the same three chunks."""
    import os
    import torch
    n_call = 1_100_000 if kind != "decimator" else 1024 * 256 * 8 + 8 * 321
    calls = 3
    if kind == "decimator":
        mk_g, mk_r = (lambda: ob.FirDecimator(2.4e6, 8, 100e3, 38400.0)), (lambda: oracle.FirDecimator(2.4e6, 8, 100e3, 38400.0))
        x, ratio, odt = noise_c64(calls * n_call, seed=31), 8, np.complex64
    elif kind == "lp_cascade":
        mk_g, mk_r = (lambda: ob.LpCascade(48e3, 4.5e3)), (lambda: oracle.LpCascade(48e3, 4.5e3))
        x, ratio, odt = noise_f32(calls * n_call, seed=32), 1, np.float32
    elif kind == "fm_rate1":
        mk_g = lambda: ob.FmQuadratureDemod(300e3, 25e3, 15e3).with_translate(20e3)
        mk_r = lambda: oracle.FmQuadratureDemod(300e3, 25e3, 15e3).with_translate(20e3)
        x, ratio, odt = fm_iq(calls * n_call, 300e3, f_c=20e3), 1, np.float32
    elif kind == "pm_rate1":
        mk_g, mk_r = (lambda: ob.PmQuadratureDemod(48e3, 0.8, 5e3)), (lambda: oracle.PmQuadratureDemod(48e3, 0.8, 5e3))
        x, ratio, odt = pm_iq(calls * n_call, 48e3), 1, np.float32
    elif kind == "rotator":
        mk_g, mk_r = (lambda: ob.Rotator(1.5e3, 48e3)), (lambda: oracle.Rotator(1.5e3, 48e3))
        x, ratio, odt = noise_c64(calls * n_call, seed=33), 1, np.complex64
    else:                                                   # AM: the DC blocker's slow pole keeps the serialised launch path
        mk_g, mk_r = (lambda: ob.AmEnvelopeDemod(48e3, 5e3)), (lambda: oracle.AmEnvelopeDemod(48e3, 5e3))
        x, ratio, odt = am_iq(calls * n_call, 48e3), 1, np.float32
    xd = torch.from_numpy(x).cuda()
    n_out = -(-n_call // ratio)
    isz, osz = x.dtype.itemsize, np.dtype(odt).itemsize

    def run(no_overlap):
        if no_overlap:
            os.environ["ORION_B200_NO_OVERLAP"] = "1"
        else:
            os.environ.pop("ORION_B200_NO_OVERLAP", None)
        g = mk_g()
        yd = torch.zeros(calls * n_out * osz // 4, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        for c in range(calls):
            wr = g.process_dev(xd.data_ptr() + c * n_call * isz, n_call, yd.data_ptr() + c * n_out * osz, n_out)
            assert tuple(wr) == (n_call, n_out)
        g.synchronize()
        os.environ.pop("ORION_B200_NO_OVERLAP", None)
        return yd.cpu().numpy().view(odt)

    over, plain = run(False), run(True)
    assert bit_equal(over, plain)          # also for AM: the chained look-back of the DC blocker follows a fixed recipe
    r = mk_r()
    ref = np.concatenate([r.run(x[c * n_call:(c + 1) * n_call]) for c in range(calls)])
    if kind == "rotator":
        # absolute phase: the exact-replay oscillator reproduces the reference's f32 phasor recurrence bit for bit
        assert bit_equal(over, ref)
    else:
        assert_parity(over, ref, what=f"{kind}: 3 back-to-back calls vs oracle streaming")


# ---- HalfCosineMf (src/dsp/fir.rs:317-376) ---------------------------------------------
@pytest.mark.parametrize("sps,n", [(32, 40_000), (1, 1000), (256, 100_001)])
def test_half_cosine_mf(sps, n):
    x = noise_c64(n, seed=sps)
    a, b = run_pair(ob.HalfCosineMf(sps), oracle.HalfCosineMf(sps), x, np.complex64)
    assert_parity(a, b, what="staged")
    g = ob.HalfCosineMf(sps)
    g.set_option(ob.OPT_FIR_GLOBAL, 1)
    a2, _ = run_pair(g, oracle.HalfCosineMf(sps), x, np.complex64)
    assert bit_equal(a2, b), "reference-order (unfused) matched filter must be bit-identical"


# ---- checkpoint / resume (the reference's blocks are Clone) -------------------------------------------------------
@pytest.mark.parametrize("kind", ["c1_chain", "ssb", "am", "fir_iq", "rotator"])
def test_snapshot_restore_continues_the_stream_bit_for_bit(kind):
    n = 120_000
    if kind == "c1_chain":
        taps = ob.fir_lowpass_design(2.4e6, 100e3, 38400.0)
        mk = lambda: ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=8, demod=ob.DEMOD_FM, fs_demod=3e5, p0=25e3,
                              audio_bw_hz=15e3, translate_hz=100e3)
        x = fm_iq(2 * n, 2.4e6)
    elif kind == "ssb":
        mk, x = (lambda: ob.SsbProductDemod(48e3, 1.5e3, 2800.0)), ssb_iq(2 * n, 48e3, f_bfo=1.5e3)
    elif kind == "am":
        mk, x = (lambda: ob.AmEnvelopeDemod(48e3, 5e3)), am_iq(2 * n, 48e3)
    elif kind == "fir_iq":
        mk, x = (lambda: ob.FirLowpassIq(81, 0.1, 60.0)), noise_c64(2 * n, seed=5)
    else:
        mk, x = (lambda: ob.Rotator(1.5e3, 48e3)), noise_c64(2 * n, seed=6)
    a = mk()
    a.run(x[:n])
    blob = a.snapshot()
    tail_a = a.run(x[n:])
    b = mk()
    b.run(x[:777])                         # some unrelated history that restore must wipe
    b.restore(blob)
    tail_b = b.run(x[n:])
    assert bit_equal(tail_a, tail_b)
    with pytest.raises(ob.OrionB200Error):
        ob.LpCascade(48e3, 4.5e3).restore(blob)          # a block of a different shape refuses the blob


# ---- modulators on the GPU (next-row scope: AmDsbMod, PmDirectPhaseMod) ---------------------------------------------
@pytest.mark.parametrize("rf", [0.0, 1.5e3])
def test_am_pm_modulators(rf):
    fs, n = 48e3, 65_536
    audio = (0.5 * np.sin(2 * np.pi * 1e3 * np.arange(n) / fs)).astype(np.float32)
    a, b = run_pair(ob.AmDsbMod(fs, rf, 0.8, 0.5), oracle.AmDsbMod(fs, rf, 0.8, 0.5), audio, np.complex64)
    assert_parity(a, b, what="AmDsbMod")
    g, r = ob.AmDsbMod(fs, rf, 0.8, 0.9), oracle.AmDsbMod(fs, rf, 0.8, 0.9)
    g.set_clamp(True); r.set_clamp(True); g.set_gain(0.7); r.set_gain(0.7)
    a, b = run_pair(g, r, (3.0 * audio).astype(np.float32), np.complex64)
    assert_parity(a, b, what="AmDsbMod clamped")
    a, b = run_pair(ob.PmDirectPhaseMod(fs, 0.9, rf), oracle.PmDirectPhaseMod(fs, 0.9, rf), audio, np.complex64)
    assert_parity(a, b, what="PmDirectPhaseMod")
    # streaming: two calls continue the carrier phase
    g, r = ob.PmDirectPhaseMod(fs, 0.9, rf), oracle.PmDirectPhaseMod(fs, 0.9, rf)
    a, b = stream_pair(g, r, audio, np.complex64, [10_001, n - 10_001])
    assert_parity(a, b, what="PmDirectPhaseMod two calls")


def test_gpu_modulator_into_gpu_demodulator_roundtrip():
    fs, n = 48e3, 32_768
    k = np.arange(n, dtype=np.float32)
    audio = (np.float32(0.5) * np.sin(np.float32(2 * np.pi) * np.float32(1e3) * k / np.float32(fs))).astype(np.float32)
    iq = ob.AmDsbMod(fs, 0.0, 0.8, 0.5).run(audio)
    y = ob.AmEnvelopeDemod(fs, 5e3).run(iq)[n // 4:]
    t = np.arange(y.size) / fs
    p = lambda f: np.abs(np.sum(y * np.exp(-2j * np.pi * f * t))) ** 2
    assert 10 * np.log10(p(1e3) / max(p(730.0), 1e-30)) > 24.0          # tests/roundtrip/am.rs:26


def test_first_long_launch_is_ordered_after_the_blocks_own_initialisation():
    """Regression: the link records of a long call are cleared on the block's stream.  A plain cudaMemset (legacy default
    stream) queued behind a caller's default-stream work once ran while the first launches were already publishing."""
    import torch
    fs, m, n = 2.4e6, 8, 40_000_000                       # 19 532 tiles: a 19 MB link array to clear at the first launch
    x = torch.from_numpy(fm_iq(4_000_000, fs)).cuda().repeat(10)
    taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
    ch = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3,
                  audio_bw_hz=15e3, translate_hz=100e3)
    y = torch.zeros(3 * (n // m), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    junk = [torch.randn(200_000_000, device="cuda") for _ in range(6)]        # several ms of default-stream work in the queue
    for c in range(3):
        ch.process_dev(x.data_ptr(), n, y.data_ptr() + c * (n // m) * 4, n // m)
    ch.synchronize()                                       # raises if a device watchdog tripped
    del junk
    head = y[:25_000].cpu().numpy()
    dec = oracle.FirDecimator(fs, m, 100e3, 38400.0)
    fm = oracle.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3)
    assert_parity(head, fm.run(dec.run(x[:200_000].cpu().numpy())), what="head of the first long call")


# ---- round-2 host-side features -----------------------------------------------------------------------------------
def test_prepare_oscillator_walks_ahead_without_changing_results():
    """orion_b200_block_prepare_oscillator: the exact-replay walk done ahead of the stream; process() then only copies anchors."""
    fs, n, calls = 2.4e6, 300_000, 3
    x = noise_c64(n * calls, seed=77)
    a, b = ob.Rotator(1e5, fs), ob.Rotator(1e5, fs)
    b.prepare_oscillator(n, calls)
    walked = b.exact_host_ms
    assert walked > 0.0
    ya = np.concatenate([a.run(x[i * n:(i + 1) * n]) for i in range(calls)])
    yb = np.concatenate([b.run(x[i * n:(i + 1) * n]) for i in range(calls)])
    assert bit_equal(ya, yb)
    assert b.exact_host_ms - walked < 0.5 * max(a.exact_host_ms, 1e-3) + 0.2     # nothing left to walk inside the calls
    ref = oracle.Rotator(1e5, fs).run(x)
    assert bit_equal(ya, ref)


@pytest.mark.parametrize("pageable", [True, False])
def test_pipelined_host_call_matches_the_oracle(pageable):
    """Host-pointer calls longer than two chunks are pipelined (copy in / kernel / copy out overlap); pageable caller
    buffers go through the pinned staging ring.  WorkReport and values as for one launch."""
    import torch
    fs, m = 2.4e6, 8
    n = 5_000_003                                              # ragged: the last chunk is short and not a multiple of m
    x = fm_iq(n, fs)
    taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
    ch = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3,
                  audio_bw_hz=15e3, translate_hz=100e3)
    n_out = -(-n // m)
    if pageable:
        xin, out = x, np.zeros(n_out, np.float32)
    else:
        xin = torch.from_numpy(x).pin_memory().numpy()
        out = torch.zeros(n_out, dtype=torch.float32).pin_memory().numpy()
    wr = ch.process(xin, out)
    assert tuple(wr) == (n, n_out)
    dec = oracle.FirDecimator(fs, m, 100e3, 38400.0)
    fm = oracle.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3)
    assert_parity(out, fm.run(dec.run(x)), what=f"pipelined host call, pageable={pageable}")
    # a rate-1 block with an exact oscillator through the same path: still bit for bit
    r = ob.Rotator(-2.5e5, 1.2e6)
    xr = noise_c64(3_000_001, seed=5)
    assert bit_equal(r.run(xr), oracle.Rotator(-2.5e5, 1.2e6).run(xr))
