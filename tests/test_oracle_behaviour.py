"""The reference's own behavioural tests for the hot path, transcribed against the CPU oracle.

The reference holds no golden vectors for this path; what it does assert --
lengths, SNR thresholds, filter-response bounds, filter_aligned == streamed, instance determinism --
is restated here assertion for assertion, with the reference's signal generators (f32 arithmetic)
and thresholds.  This is what pins the oracle to the reference's observable behaviour; the
independent numpy restatement (test_oracle_crosscheck.py) pins its arithmetic.

Sources: /root/reference/tests/unit/{dsp,fm,pm,ssb,chains}.rs, tests/roundtrip/{fm,am,pm,ssb,cw}.rs,
tests/common/mod.rs, tests/roundtrip/helpers.rs, python/tests/test_unit.py.
"""
import numpy as np
import pytest

import oracle

F32 = np.float32
TAU = F32(6.28318530717958647692)
PI = F32(3.14159265358979323846)


# ---- helpers, tests/common/mod.rs:9-24, tests/unit/helpers.rs:5-16, tests/roundtrip/helpers.rs:7-23 ----
def snr_db_at(fs, f0, x):
    """Power of the projection on f0 against the projection on 0.73*f0 (common/mod.rs:9-24)."""
    x = np.asarray(x, np.float64)
    n = max(x.size, 1)
    k = np.arange(x.size, dtype=np.float64)

    def proj(f):
        t = (-2.0 * np.pi * f / fs) * k
        re, im = np.sum(x * np.cos(t)), np.sum(x * np.sin(t))
        return (re * re + im * im) / (n * n)
    return 10.0 * np.log10(proj(f0) / (proj(f0 * 0.73) + 1e-20))


def dft_power(x, fs, f):                                   # unit/helpers.rs:5-16
    x = np.asarray(x, np.float64)
    k = np.arange(x.size, dtype=np.float64)
    w = 2.0 * np.pi * f / fs
    re, im = np.sum(x * np.cos(w * k)), -np.sum(x * np.sin(w * k))
    return (re * re + im * im) / (x.size * x.size)


def real_tone(fs, f_hz, n, amp):                           # roundtrip/helpers.rs:19-23
    k = np.arange(n, dtype=F32)
    return (F32(amp) * np.sin(TAU * F32(f_hz) * k / F32(fs))).astype(F32)


def tail(x):                                               # roundtrip/helpers.rs:15-17
    return x[x.size // 4:]


def response_db(taps, f):                                  # unit/dsp.rs:33-41
    n = np.arange(taps.size, dtype=np.float64)
    h = np.sum(taps.astype(np.float64) * np.exp(-2j * np.pi * f * n))
    return 20.0 * np.log10(max(abs(h), 1e-12))


# ---- tests/unit/dsp.rs ------------------------------------------------------------------------------
def test_decimator_reduces_length_and_preserves_tone():    # dsp.rs:12-27
    fs, m, n = 96_000.0, 4, 4096
    dec = oracle.FirDecimator(fs, m, fs / m * 0.45, fs / m * 0.10)
    iq = oracle.Nco(2_000.0, fs).mix(np.ones(n, np.complex64))
    out = np.zeros(n // m, np.complex64)
    w = dec.process(iq, out)
    assert w.out_written == n // m
    assert w.in_read == n                                  # decim.rs:72-75


def test_kaiser_taps_are_linear_phase_and_unit_dc_gain():  # dsp.rs:44-60
    for req in (3, 16, 31, 64, 101):
        taps = oracle.kaiser_lowpass_taps(req, 0.2, 60.0)
        assert taps.size == max(req, 3) | 1
        m = taps.size
        for i in range(m // 2):
            assert abs(taps[i] - taps[m - 1 - i]) < 1e-6
        assert abs(float(np.sum(taps, dtype=np.float32)) - 1.0) < 1e-5


def test_kaiser_lowpass_meets_its_stopband_target():       # dsp.rs:63-91
    num_taps, cutoff, a_db = 101, 0.2, 60.0
    taps = oracle.kaiser_lowpass_taps(num_taps, cutoff, a_db)
    half = 0.5 * oracle.kaiser_transition_norm(num_taps, a_db)
    for f in (0.0, 0.05, 0.1, cutoff - half):
        assert abs(response_db(taps, f)) < 0.5
    assert abs(response_db(taps, cutoff) + 6.0) < 1.0
    for f in (cutoff + half, 0.3, 0.4, 0.5):
        assert response_db(taps, f) < -(a_db - 5.0)


def test_kaiser_num_taps_inverts_transition_norm():        # dsp.rs:94-112
    for transition, a_db in ((0.02, 60.0), (0.05, 40.0), (0.084, 60.0)):
        m = oracle.kaiser_num_taps(transition, a_db)
        assert m % 2 == 1
        assert oracle.kaiser_transition_norm(m, a_db) <= transition * 1.001
        assert oracle.kaiser_transition_norm(max(m - 2, 0), a_db) > transition * 0.999


def test_fir_lowpass_iq_passes_in_band_and_rejects_out_of_band():   # dsp.rs:115-144
    num_taps, cutoff, a_db, n = 81, 0.2, 60.0, 2048

    def amplitude(f):
        i = np.arange(n, dtype=F32)
        ph = TAU * F32(f) * i
        x = (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)
        y = oracle.FirLowpassIq(num_taps, cutoff, a_db).run(x)
        return float(np.max(np.abs(y[2 * num_taps + 1:])))
    in_band, out_band = amplitude(0.1), amplitude(0.35)
    assert abs(in_band - 1.0) < 0.02
    assert 20.0 * np.log10(max(out_band, 1e-12) / in_band) < -(a_db - 5.0)


def test_filter_aligned_is_same_length_and_group_delay_compensated():   # dsp.rs:147-198
    num_taps, cutoff, n = 31, 0.2, 512
    i = np.arange(n, dtype=F32)
    p = TAU * F32(0.03) * i
    env = np.exp(-(((i - F32(200.0)) / F32(60.0)) ** 2)).astype(F32)
    x = (env * np.cos(p) + 1j * env * np.sin(p)).astype(np.complex64)
    d = oracle.FirLowpassIq(num_taps, cutoff, 60.0).group_delay()
    assert d == (num_taps - 1) // 2
    streamed = oracle.FirLowpassIq(num_taps, cutoff, 60.0).run(np.concatenate([x, np.zeros(d, np.complex64)]))
    aligned = x.copy()
    oracle.FirLowpassIq(num_taps, cutoff, 60.0).filter_aligned(aligned)
    assert aligned.size == n
    assert np.max(np.abs(aligned - streamed[d:d + n])) < 1e-5
    assert abs(int(np.argmax(np.abs(aligned))) - int(np.argmax(np.abs(x)))) <= 1


# ---- tests/unit/{fm,pm,ssb,chains}.rs -----------------------------------------------------------------
def test_fm_quadrature_demod_recovers_tone():              # unit/fm.rs:10-28
    fs, n, f_mod, dev = F32(48_000.0), 16_384, F32(1_000.0), F32(2_500.0)
    t = np.arange(n, dtype=F32) / fs
    f_inst = dev * np.sin(F32(2.0) * PI * f_mod * t)
    phi = np.zeros(n, F32)
    acc = F32(0.0)
    inc = (F32(2.0) * PI * f_inst / fs).astype(F32)
    for k in range(n):                                     # sequential f32 accumulation, as in the test
        acc = F32(acc + inc[k])
        phi[k] = acc
    iq = (np.cos(phi) + 1j * np.sin(phi)).astype(np.complex64)
    y = oracle.FmQuadratureDemod(48_000.0, 2_500.0, 5_000.0).run(iq)
    assert snr_db_at(48_000.0, 1_000.0, y) > 20.0


def test_pm_quadrature_demod_recovers_tone():              # unit/pm.rs:10-26
    fs, n, f_mod, beta = F32(48_000.0), 16_384, F32(1_000.0), F32(0.8)
    t = np.arange(n, dtype=F32) / fs
    phi = beta * np.sin(F32(2.0) * PI * f_mod * t)
    iq = (np.cos(phi) + 1j * np.sin(phi)).astype(np.complex64)
    y = oracle.PmQuadratureDemod(48_000.0, 0.8, 5_000.0).run(iq)
    assert snr_db_at(48_000.0, 1_000.0, y) > 20.0


def gen_complex_tone(fs, f_hz, n):                         # src/util.rs:33-40
    ph = TAU * F32(f_hz) * np.arange(n, dtype=F32) / F32(fs)
    return (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)


def test_ssb_product_demod_yields_strong_tone_and_low_dc():   # unit/ssb.rs:10-34
    fs, n, f_tone = 48_000.0, 16_384, 1_000.0
    audio = oracle.SsbProductDemod(fs, 0.0, 2_800.0).run(gen_complex_tone(fs, f_tone, n))
    assert abs(float(np.sum(audio, dtype=np.float32)) / n) < 1e-3
    assert 10.0 * np.log10(dft_power(audio, fs, f_tone) / (dft_power(audio, fs, 700.0) + 1e-20)) > 25.0


@pytest.mark.parametrize("mk", [lambda: oracle.CwEnvelopeDemod(48_000.0, 700.0, 300.0),
                                lambda: oracle.AmEnvelopeDemod(48_000.0, 5_000.0),
                                lambda: oracle.SsbProductDemod(48_000.0, 1_500.0, 2_800.0)])
def test_chains_return_input_len_items(mk):                # unit/chains.rs:10-33
    iq = gen_complex_tone(48_000.0, 700.0, 4096)
    assert oracle.IqToAudioChain(mk()).process(iq).size == iq.size


# ---- tests/roundtrip/*.rs ------------------------------------------------------------------------------
def test_roundtrip_fm_quadrature():                        # roundtrip/fm.rs:11-27
    fs, n, f_mod = 48_000.0, 32_768, 1_000.0
    iq = oracle.AudioToIqChain(oracle.FmPhaseAccumMod(fs, 2_500.0, 0.0)).process(real_tone(fs, f_mod, n, 0.5))
    out = oracle.IqToAudioChain(oracle.FmQuadratureDemod(fs, 2_500.0, 5_000.0)).process(iq)
    assert snr_db_at(fs, f_mod, tail(out)) > 20.0


def test_roundtrip_am_envelope():                          # roundtrip/am.rs:11-27
    fs, n, f_mod = 48_000.0, 32_768, 1_000.0
    iq = oracle.AudioToIqChain(oracle.AmDsbMod(fs, 0.0, 0.8, 0.5)).process(real_tone(fs, f_mod, n, 0.5))
    out = oracle.IqToAudioChain(oracle.AmEnvelopeDemod(fs, 5_000.0)).process(iq)
    assert snr_db_at(fs, f_mod, tail(out)) > 24.0


def test_roundtrip_am_abs_approx():                        # python/tests/test_roundtrip.py:65-74
    fs, n, f_mod = 48_000.0, 32_768, 1_000.0
    iq = oracle.AmDsbMod(fs, 0.0, 0.8, 0.5).run(real_tone(fs, f_mod, n, 0.5))
    out = oracle.AmEnvelopeDemod(fs, 5_000.0, abs_approx=True).run(iq)
    assert snr_db_at(fs, f_mod, tail(out)) > 20.0


def test_roundtrip_pm_quadrature():                        # roundtrip/pm.rs:11-27
    fs, n, f_mod = 48_000.0, 32_768, 900.0
    iq = oracle.AudioToIqChain(oracle.PmDirectPhaseMod(fs, 0.9, 0.0)).process(real_tone(fs, f_mod, n, 0.5))
    out = oracle.IqToAudioChain(oracle.PmQuadratureDemod(fs, 0.9, 5_000.0)).process(iq)
    assert snr_db_at(fs, f_mod, tail(out)) > 18.0


def test_roundtrip_ssb_usb_product():                      # roundtrip/ssb.rs:10-33
    fs, n, f_audio = 48_000.0, 32_768, 1_200.0
    audio_in = real_tone(fs, f_audio, n, 0.4)
    iq = oracle.AudioToIqChain(oracle.SsbPhasingMod(fs, 2_800.0, 1_500.0, 0.0, True)).process(audio_in)
    out = oracle.IqToAudioChain(oracle.SsbProductDemod(fs, 1_500.0, 2_800.0)).process(iq)
    assert snr_db_at(fs, f_audio, out[int(0.120 * fs):]) > 18.0


def test_roundtrip_cw_envelope():                          # roundtrip/cw.rs:10-48
    fs, n, pitch = 48_000.0, 24_000, 700.0
    k = np.arange(n, dtype=F32)
    key_env = ((k * F32(5.0) / F32(fs)) % F32(1.0) < F32(0.5)).astype(F32)
    iq = oracle.AudioToIqChain(oracle.CwKeyedMod(fs, pitch, 3.0, 3.0)).process(key_env)
    audio = oracle.IqToAudioChain(oracle.CwEnvelopeDemod(fs, pitch, 300.0)).process(iq)
    skip = int(0.100 * fs)
    a, ke = audio[skip:], key_env[skip:]
    rms = lambda v: float(np.sqrt(np.mean(np.square(v, dtype=np.float64)))) if v.size else 0.0
    contrast = 20.0 * np.log10(rms(a[ke > 0.5]) / (rms(a[ke <= 0.5]) + 1e-12))
    assert contrast > 14.0


# ---- python/tests/test_unit.py:284-303 -----------------------------------------------------------------
def test_two_instances_are_independent_and_state_persists():
    iq = gen_complex_tone(48_000.0, 1_000.0, 4096)
    a = oracle.FmQuadratureDemod(48_000.0, 2_500.0, 5_000.0)
    b = oracle.FmQuadratureDemod(48_000.0, 2_500.0, 5_000.0)
    np.testing.assert_array_equal(a.run(iq), b.run(iq))
    # state persists: a second call continues the stream rather than restarting it
    c = oracle.FmQuadratureDemod(48_000.0, 2_500.0, 5_000.0)
    whole = c.run(np.concatenate([iq, iq]))
    np.testing.assert_array_equal(a.run(iq), whole[4096:])


def test_agc_rms_converges_on_iq():                           # tests/unit/agc.rs:8-32
    fs, n = 48_000.0, 8_000
    x = np.where(np.arange(n) < n // 2, 0.02, 1.0).astype(np.float32).astype(np.complex64)
    out = oracle.AgcRmsIq(fs, 0.2, 5.0, 0.2).run(x)
    tail = out[n - 1000:]
    rms_tail = np.sqrt(np.mean(np.abs(tail).astype(np.float64) ** 2))
    assert abs(rms_tail - 0.2) < 0.03, rms_tail


# ---- reference-held known answers ----------------------------------------------------------------------------------------
# The only literal input/output vectors the reference's own tests hold for anything on (or next to) the path are the
# constellation points of its mappers (tests/unit/bpsk.rs:9-19, qpsk.rs:9-18, qam.rs:9-41).  The deciders are their
# inverses (roundtrip tests feed one into the other), so these vectors pin the decider restatement to numbers written
# down by the reference itself; the GPU deciders are checked against the same vectors in tests/test_gpu_next_rows.py.
REFERENCE_MAPPER_VECTORS = {
    1: ([0, 1, 0, 1, 1, 0], [(1.0, 0.0), (-1.0, 0.0), (1.0, 0.0), (-1.0, 0.0), (-1.0, 0.0), (1.0, 0.0)]),
    2: ([0, 0, 0, 1, 1, 0, 1, 1], [(np.sqrt(0.5), np.sqrt(0.5)), (np.sqrt(0.5), -np.sqrt(0.5)),
                                   (-np.sqrt(0.5), np.sqrt(0.5)), (-np.sqrt(0.5), -np.sqrt(0.5))]),
    4: ([0, 0, 0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0],
        [(-3.0 * np.sqrt(0.1), -3.0 * np.sqrt(0.1)), (-np.sqrt(0.1), -3.0 * np.sqrt(0.1)),
         (np.sqrt(0.1), -3.0 * np.sqrt(0.1)), (3.0 * np.sqrt(0.1), -3.0 * np.sqrt(0.1))]),
}


def reference_mapper_vector(bits):
    b, pts = REFERENCE_MAPPER_VECTORS[bits]
    return np.array(b, np.uint8), np.array([complex(np.float32(re), np.float32(im)) for re, im in pts], np.complex64)


@pytest.mark.parametrize("bits", [1, 2, 4])
def test_deciders_invert_the_reference_mapper_known_answers(bits):
    from oracle import np_oracle as npo
    want, syms = reference_mapper_vector(bits)
    fn = {1: npo.bpsk_decide, 2: npo.qpsk_decide}.get(bits, lambda v: npo.qam_decide(v, bits))
    assert np.array_equal(fn(syms), want)
    # ... and with the symbols pulled 30 % of the way towards the neighbouring decision boundary (the round-trip tests add noise)
    assert np.array_equal(fn((syms * np.float32(0.85)).astype(np.complex64)), want)
    if bits == 4:                                        # qam.rs:11: the axis scale the test itself computes, (1/10).sqrt() in f32
        assert npo.qam_axis_scale(4) == np.sqrt(np.float32(1.0) / np.float32(10.0))
