"""Drop-in for the five analog demodulator classes of the reference's Python package
(`python/orion_sdr/__init__.py:2-6`, PyO3 wrappers `src/python/demodulate.rs:11-148`): same
class names, constructor arguments, `process(iq: complex64[n]) -> float32[n]` (a new array per
call), same input validation behaviour (`ValueError`/`TypeError` for wrong dtype, ndim or a
non-contiguous view -- `python/tests/test_unit.py:93-129`).  Each class forwards to the same
GPU block the C ABI exposes; nothing is computed on the CPU.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

import orion_b200 as _ob  # noqa: E402

__all__ = ["CwEnvelopeDemod", "AmEnvelopeDemod", "SsbProductDemod", "FmQuadratureDemod", "PmQuadratureDemod"]


def _as_slice(iq):
    """PyReadonlyArray1<Complex32>::as_slice()? -- src/python/demodulate.rs:32."""
    if not isinstance(iq, np.ndarray):
        raise TypeError("argument 'iq': expected a numpy.ndarray of complex64")
    if iq.dtype != np.complex64:
        raise TypeError(f"argument 'iq': dtype {iq.dtype} cannot be converted to complex64 array")
    if iq.ndim != 1:
        raise TypeError(f"argument 'iq': expected a 1-D array, got {iq.ndim}-D")
    if not iq.flags.c_contiguous:
        raise ValueError("The given array is not contiguous")
    return iq


class _Demod:
    _block: _ob.Block

    def process(self, iq: np.ndarray) -> np.ndarray:
        x = _as_slice(iq)
        out = np.zeros(x.size, np.float32)                            # vec![0.0f32; n]
        if x.size:
            self._block.process(x, out)
        return out


class CwEnvelopeDemod(_Demod):                                        # src/python/demodulate.rs:11-37
    def __init__(self, sample_rate: float, tone_hz: float, env_bw_hz: float):
        self._block = _ob.CwEnvelopeDemod(sample_rate, tone_hz, env_bw_hz)

    def set_gain(self, g: float) -> None:
        self._block.set_gain(g)


class AmEnvelopeDemod(_Demod):                                        # src/python/demodulate.rs:41-67
    def __init__(self, fs: float, audio_bw_hz: float, abs_approx: bool = False):
        self._block = _ob.AmEnvelopeDemod(fs, audio_bw_hz)
        if abs_approx:
            self._block.with_abs_approx(0.9482, 0.3920)               # demodulate.rs:50


class SsbProductDemod(_Demod):                                        # src/python/demodulate.rs:71-94
    def __init__(self, fs: float, bfo_hz: float, audio_bw_hz: float):
        self._block = _ob.SsbProductDemod(fs, bfo_hz, audio_bw_hz)


class FmQuadratureDemod(_Demod):                                      # src/python/demodulate.rs:98-121
    def __init__(self, fs: float, dev_hz: float, audio_bw_hz: float):
        self._block = _ob.FmQuadratureDemod(fs, dev_hz, audio_bw_hz)


class PmQuadratureDemod(_Demod):                                      # src/python/demodulate.rs:125-148
    def __init__(self, fs: float, k: float, audio_bw_hz: float):
        self._block = _ob.PmQuadratureDemod(fs, k, audio_bw_hz)
