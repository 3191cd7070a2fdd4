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
