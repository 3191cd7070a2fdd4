"""The reference's own Python tests for the five analog demodulators, run against the drop-in
`orion_sdr` shim (python/tests/test_unit.py:37-60,93-129,136-145,284-303 and
python/tests/test_roundtrip.py thresholds), plus the Rust behavioural thresholds of
tests/unit/{fm,pm}.rs and tests/roundtrip/am.rs restated on numpy-generated signals."""
import numpy as np
import pytest

import orion_sdr
from signals import am_iq, fm_iq, pm_iq

pytestmark = pytest.mark.gpu
FS = 48_000.0


def _all():
    return [orion_sdr.CwEnvelopeDemod(FS, 700.0, 50.0), orion_sdr.AmEnvelopeDemod(FS, 5e3),
            orion_sdr.AmEnvelopeDemod(FS, 5e3, abs_approx=True), orion_sdr.SsbProductDemod(FS, 1.5e3, 2.8e3),
            orion_sdr.FmQuadratureDemod(FS, 2.5e3, 5e3), orion_sdr.PmQuadratureDemod(FS, 1.0, 5e3)]


def test_shape_and_dtype():                                   # test_unit.py:37-60
    iq = np.zeros(4096, np.complex64)
    for d in _all():
        out = d.process(iq)
        assert out.shape == (4096,) and out.dtype == np.float32


def test_input_validation():                                  # test_unit.py:93-129
    d = orion_sdr.FmQuadratureDemod(FS, 2.5e3, 5e3)
    with pytest.raises((ValueError, TypeError)):
        d.process(np.zeros(16, np.complex128))
    with pytest.raises((ValueError, TypeError)):
        d.process(np.zeros(16, np.float32))
    with pytest.raises((ValueError, TypeError)):
        d.process(np.zeros((4, 4), np.complex64))
    with pytest.raises((ValueError, TypeError)):
        d.process(np.zeros(32, np.complex64)[::2])
    with pytest.raises((ValueError, TypeError)):
        d.process([0j, 1j])


def test_set_gain():                                          # test_unit.py:136-145
    iq = np.full(2048, 0.5 + 0j, np.complex64)
    a, b = orion_sdr.CwEnvelopeDemod(FS, 700.0, 50.0), orion_sdr.CwEnvelopeDemod(FS, 700.0, 50.0)
    b.set_gain(2.0)
    ya, yb = a.process(iq), b.process(iq)
    assert np.allclose(yb, 2.0 * ya, rtol=1e-6, atol=1e-9)


def test_instances_independent_and_state_persists():          # test_unit.py:284-303
    iq = fm_iq(8192, FS, f_c=0.0, dev=2.5e3)
    a, b = orion_sdr.FmQuadratureDemod(FS, 2.5e3, 5e3), orion_sdr.FmQuadratureDemod(FS, 2.5e3, 5e3)
    np.testing.assert_array_equal(a.process(iq), b.process(iq))
    one = orion_sdr.FmQuadratureDemod(FS, 2.5e3, 5e3).process(iq)
    c = orion_sdr.FmQuadratureDemod(FS, 2.5e3, 5e3)
    two = np.concatenate([c.process(iq[:5000]), c.process(iq[5000:])])
    assert np.allclose(one, two, rtol=0, atol=1e-4 * np.max(np.abs(one)))


def _snr_db_at(y, fs, f0):                                    # tests/common/mod.rs:9-24
    n = y.size
    t = np.arange(n) / fs

    def p(f):
        return np.abs(np.sum(y * np.exp(-2j * np.pi * f * t))) ** 2
    return 10 * np.log10(p(f0) / max(p(0.73 * f0), 1e-30))


def test_fm_pm_am_tone_snr_thresholds():                      # tests/unit/fm.rs:10-28, pm.rs:10-26, roundtrip/am.rs
    n = 16_384
    y = orion_sdr.FmQuadratureDemod(FS, 2.5e3, 5e3).process(
        fm_iq(n, FS, f_c=0.0, dev=2.5e3, tones=((1e3, 0.8),), sigma=0.0))
    assert _snr_db_at(y[n // 4:], FS, 1e3) > 20.0
    y = orion_sdr.PmQuadratureDemod(FS, 1.0, 5e3).process(pm_iq(n, FS, k=0.8, tone=1e3, sigma=0.0))
    assert _snr_db_at(y[n // 4:], FS, 1e3) > 18.0
    y = orion_sdr.AmEnvelopeDemod(FS, 5e3).process(am_iq(2 * n, FS, tones=(1e3,), sigma=0.0))
    assert _snr_db_at(y[n:], FS, 1e3) > 24.0
    y = orion_sdr.AmEnvelopeDemod(FS, 5e3, abs_approx=True).process(am_iq(2 * n, FS, tones=(1e3,), sigma=0.0))
    assert _snr_db_at(y[n:], FS, 1e3) > 20.0


# ---- the reference's round-trip tests (tests/roundtrip/*.rs, python/tests/test_roundtrip.py:25-123): the reference's own
# ---- modulators (CPU restatement in the oracle -- the step BEFORE the path) feeding the GPU demodulators of the shim ----
def _real_tone(fs, f_hz, n, amp):                             # tests/roundtrip/helpers.rs:19-23
    k = np.arange(n, dtype=np.float32)
    return (np.float32(amp) * np.sin(np.float32(2 * np.pi) * np.float32(f_hz) * k / np.float32(fs))).astype(np.float32)


def test_roundtrips_with_reference_modulators():
    import oracle
    n = 32_768
    iq = oracle.FmPhaseAccumMod(FS, 2_500.0, 0.0).run(_real_tone(FS, 1_000.0, n, 0.5))          # roundtrip/fm.rs:11-27
    assert _snr_db_at(orion_sdr.FmQuadratureDemod(FS, 2_500.0, 5_000.0).process(iq)[n // 4:], FS, 1_000.0) > 20.0
    iq = oracle.AmDsbMod(FS, 0.0, 0.8, 0.5).run(_real_tone(FS, 1_000.0, n, 0.5))                  # roundtrip/am.rs:11-27
    assert _snr_db_at(orion_sdr.AmEnvelopeDemod(FS, 5_000.0).process(iq)[n // 4:], FS, 1_000.0) > 24.0
    assert _snr_db_at(orion_sdr.AmEnvelopeDemod(FS, 5_000.0, abs_approx=True).process(iq)[n // 4:], FS, 1_000.0) > 20.0
    iq = oracle.PmDirectPhaseMod(FS, 0.9, 0.0).run(_real_tone(FS, 900.0, n, 0.5))                 # roundtrip/pm.rs:11-27
    assert _snr_db_at(orion_sdr.PmQuadratureDemod(FS, 0.9, 5_000.0).process(iq)[n // 4:], FS, 900.0) > 18.0
    iq = oracle.SsbPhasingMod(FS, 2_800.0, 1_500.0, 0.0, True).run(_real_tone(FS, 1_200.0, n, 0.4))   # roundtrip/ssb.rs:10-33
    assert _snr_db_at(orion_sdr.SsbProductDemod(FS, 1_500.0, 2_800.0).process(iq)[int(0.120 * FS):], FS, 1_200.0) > 18.0
    nk = 24_000                                                                                   # roundtrip/cw.rs:10-48
    key = ((np.arange(nk, dtype=np.float32) * np.float32(5.0) / np.float32(FS)) % np.float32(1.0) < 0.5).astype(np.float32)
    iq = oracle.CwKeyedMod(FS, 700.0, 3.0, 3.0).run(key)
    audio = orion_sdr.CwEnvelopeDemod(FS, 700.0, 300.0).process(iq)
    skip = int(0.100 * FS)
    a, ke = audio[skip:], key[skip:]
    rms = lambda v: float(np.sqrt(np.mean(np.square(v, dtype=np.float64))))
    assert 20.0 * np.log10(rms(a[ke > 0.5]) / (rms(a[ke <= 0.5]) + 1e-12)) > 14.0
    # and every one of them equals the oracle's demodulator on the same IQ within the parity bar
    from signals import assert_parity
    assert_parity(audio, oracle.CwEnvelopeDemod(FS, 700.0, 300.0).run(iq), what="cw roundtrip")
