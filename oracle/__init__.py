"""ctypes front end of the CPU oracle (oracle/orion_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the shipped GPU path never imports it.
PARITY UNPINNED by reference golden vectors (none exist; see orion_oracle.h).

The classes mirror the reference's Rust blocks one-for-one (same constructor
arguments, `process(input, output) -> WorkReport` as in src/core.rs:12-22).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborion_oracle.so")

WorkReport = namedtuple("WorkReport", ["in_read", "out_written"])


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc; a few hundred ms)."""
    src = os.path.join(_HERE, "orion_oracle.c")
    hdr = os.path.join(_HERE, "orion_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


class _WR(C.Structure):
    _fields_ = [("in_read", C.c_size_t), ("out_written", C.c_size_t)]


class _C32(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    f, sz, vp, i = C.c_float, C.c_size_t, C.c_void_p, C.c_int
    fp = C.POINTER(C.c_float)

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("oo_fir_lowpass_ntaps", sz, f, f, f)
    sig("oo_fir_lowpass_design", sz, f, f, f, vp, sz)
    sig("oo_kaiser_beta", f, f)
    sig("oo_bessel_i0", f, f)
    sig("oo_kaiser_lowpass_taps", sz, sz, f, f, vp, sz)
    sig("oo_kaiser_transition_norm", f, sz, f)
    sig("oo_kaiser_num_taps", sz, f, f)
    sig("oo_lp_biquad_design", None, f, f, vp)
    sig("oo_dc_pole", f, f, f)
    sig("oo_cw_alpha", f, f, f)
    sig("oo_atan2_approx", f, f, f)
    sig("oo_free", None, vp)
    sig("oo_reset", None, vp)
    sig("oo_process", _WR, vp, vp, sz, vp, sz)
    sig("oo_fir_lowpass_new", vp, f, f, f)
    sig("oo_fir_lowpass_from_taps", vp, vp, sz)
    sig("oo_fir_decimator_new", vp, f, sz, f, f)
    sig("oo_fir_decimator_from_taps", vp, vp, sz, sz)
    sig("oo_fir_decim_kept", None, vp, sz, sz, vp, sz, vp, sz, sz)
    sig("oo_fir_iq_kept", None, vp, sz, sz, vp, sz, vp, sz, sz)
    sig("oo_fir_iq_design", vp, sz, f, f)
    sig("oo_fir_iq_from_taps", vp, vp, sz)
    sig("oo_fir_iq_group_delay", sz, vp)
    sig("oo_fir_iq_filter_aligned", None, vp, vp, sz)
    sig("oo_get_taps", sz, vp, vp, sz)
    sig("oo_rotator_new", vp, f, f)
    sig("oo_rotator_set_freq", None, vp, f, f)
    sig("oo_rotator_reset_phase", None, vp)
    sig("oo_rotator_next", _C32, vp)
    sig("oo_rotator_phasors", None, vp, vp, sz)
    sig("oo_rotator_rotate_block", None, vp, vp, vp, sz)
    sig("oo_rotator_mix_usb_block", None, vp, vp, vp, sz)
    sig("oo_nco_new", vp, f, f)
    sig("oo_nco_set_freq", None, vp, f)
    sig("oo_nco_mix", None, vp, vp, vp, sz)
    sig("oo_biquad_new", vp, f, f, f, f, f)
    sig("oo_lp_cascade_new", vp, f, f)
    sig("oo_lp_dc_cascade_new", vp, f, f, f, i)
    sig("oo_dc_blocker_new", vp, f, f)
    sig("oo_fm_demod_new", vp, f, f, f)
    sig("oo_fm_demod_with_translate", None, vp, f)
    sig("oo_pm_demod_new", vp, f, f, f)
    sig("oo_am_demod_new", vp, f, f)
    sig("oo_am_demod_with_abs_approx", None, vp, f, f)
    sig("oo_ssb_demod_new", vp, f, f, f)
    sig("oo_cw_demod_new", vp, f, f, f)
    sig("oo_cw_demod_set_gain", None, vp, f)
    sig("oo_get_state", sz, vp, vp, sz)
    sig("oo_half_cosine_taps", sz, sz, vp, sz)
    sig("oo_half_cosine_mf_new", vp, sz)
    sig("oo_fm_mod_new", vp, f, f, f)
    sig("oo_pm_mod_new", vp, f, f, f)
    sig("oo_am_mod_new", vp, f, f, f, f)
    sig("oo_am_mod_set_clamp", None, vp, i)
    sig("oo_ssb_mod_new", vp, f, f, f, f, i)
    sig("oo_cw_mod_new", vp, f, f, f, f)
    sig("oo_mod_set_gain", None, vp, f)
    sig("oo_agc_rms_new", vp, f, f, f, f)
    sig("oo_agc_rms_iq_new", vp, f, f, f, f)
    sig("oo_agc_env", f, vp)
    _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _as(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


# ---- design-time helpers ----------------------------------------------------

def fir_lowpass_taps(fs, pass_hz, trans_hz) -> np.ndarray:
    n = lib().oo_fir_lowpass_ntaps(fs, pass_hz, trans_hz)
    t = np.empty(n, np.float32)
    lib().oo_fir_lowpass_design(fs, pass_hz, trans_hz, _ptr(t), n)
    return t


def kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db) -> np.ndarray:
    n = lib().oo_kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db, None, 0)
    t = np.empty(n, np.float32)
    lib().oo_kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db, _ptr(t), n)
    return t


def kaiser_transition_norm(num_taps, stopband_db) -> float:
    return float(lib().oo_kaiser_transition_norm(num_taps, stopband_db))


def kaiser_num_taps(transition_norm, stopband_db) -> int:
    return int(lib().oo_kaiser_num_taps(transition_norm, stopband_db))


def lp_biquad_coeffs(fs, fc) -> np.ndarray:
    c = np.empty(5, np.float32)
    lib().oo_lp_biquad_design(fs, fc, _ptr(c))
    return c


def dc_pole(fs, cut_hz) -> float:
    return float(lib().oo_dc_pole(fs, cut_hz))


def cw_alpha(fs, env_bw_hz) -> float:
    return float(lib().oo_cw_alpha(fs, env_bw_hz))


def atan2_approx(y, x):
    y = np.asarray(y, np.float32)
    x = np.asarray(x, np.float32)
    yb, xb = np.broadcast_arrays(y, x)
    out = np.empty(yb.shape, np.float32)
    fn = lib().oo_atan2_approx
    for idx in np.ndindex(yb.shape):
        out[idx] = fn(float(yb[idx]), float(xb[idx]))
    return out


# ---- blocks -------------------------------------------------------------------

class Block:
    """Mirror of `trait Block` (src/core.rs:12-22) over the C oracle."""
    In = np.complex64
    Out = np.float32
    ratio = 1  # output items per input item is 1/ratio (decimator overrides)

    def __init__(self, handle):
        if not handle:
            raise MemoryError("oracle allocation failed")
        self._h = C.c_void_p(handle)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.oo_free(h)
            self._h = None

    def reset(self):
        lib().oo_reset(self._h)

    def process(self, input, output) -> WorkReport:
        x = _as(input, self.In)
        assert isinstance(output, np.ndarray) and output.dtype == self.Out and output.flags.c_contiguous
        wr = lib().oo_process(self._h, _ptr(x), x.size, _ptr(output), output.size)
        return WorkReport(wr.in_read, wr.out_written)

    process_into = process

    def run(self, input) -> np.ndarray:
        """Convenience: allocate the output like util::run_block_vec (src/util.rs:80-86)."""
        x = _as(input, self.In)
        out = np.zeros(x.size, self.Out)
        wr = self.process(x, out)
        return out[: wr.out_written]

    def state(self) -> np.ndarray:
        s = np.zeros(18, np.float32)
        lib().oo_get_state(self._h, _ptr(s), s.size)
        return s


class FirLowpass(Block):
    In = np.float32
    Out = np.float32

    def __init__(self, fs=None, pass_hz=None, trans_hz=None, taps=None):
        if taps is not None:
            t = _as(taps, np.float32)
            super().__init__(lib().oo_fir_lowpass_from_taps(_ptr(t), t.size))
        else:
            super().__init__(lib().oo_fir_lowpass_new(fs, pass_hz, trans_hz))

    def taps(self):
        n = lib().oo_get_taps(self._h, None, 0)
        t = np.empty(n, np.float32)
        lib().oo_get_taps(self._h, _ptr(t), n)
        return t


class FirDecimator(Block):
    In = np.complex64
    Out = np.complex64

    def __init__(self, fs=None, m=1, cutoff_hz=None, trans_hz=None, taps=None):
        self.m = max(int(m), 1)
        if taps is not None:
            t = _as(taps, np.float32)
            super().__init__(lib().oo_fir_decimator_from_taps(_ptr(t), t.size, self.m))
        else:
            super().__init__(lib().oo_fir_decimator_new(fs, self.m, cutoff_hz, trans_hz))

    taps = FirLowpass.taps

    def run(self, input):
        x = _as(input, self.In)
        out = np.zeros(-(-x.size // self.m), self.Out)
        wr = self.process(x, out)
        return out[: wr.out_written]


def _kept(fn_name, taps, m, x, threads=None):
    """Outputs y[j*m] of a FRESH decimating FIR over x, evaluated directly (oo_fir_*_kept; bit-identical to the
    loop-for-loop block).  Output ranges are spread over host threads (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    t = _as(taps, np.float32)
    x = _as(x, np.complex64)
    m = max(int(m), 1)
    n_out = -(-x.size // m)
    out = np.zeros(n_out, np.complex64)
    threads = threads or min(os.cpu_count() or 1, 16)
    fn = getattr(lib(), fn_name)
    step = max(-(-n_out // (threads * 4)), 1)

    def work(j0):
        j1 = min(n_out, j0 + step)
        fn(_ptr(t), t.size, m, _ptr(x), x.size, out[j0:j1].ctypes.data, j0, j1)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(0, n_out, step)))
    return out


def fir_decim_kept(taps, m, x, threads=None) -> np.ndarray:
    """== FirDecimator(taps=taps, m=m).run(x) on a fresh block, m times cheaper."""
    return _kept("oo_fir_decim_kept", taps, m, x, threads)


def fir_iq_kept(taps, m, x, threads=None) -> np.ndarray:
    """== FirLowpassIq(taps=taps).run(x)[::m] on a fresh block, m times cheaper."""
    return _kept("oo_fir_iq_kept", taps, m, x, threads)


class FirLowpassIq(Block):
    In = np.complex64
    Out = np.complex64

    def __init__(self, num_taps=None, cutoff_norm=None, stopband_db=None, taps=None):
        if taps is not None:
            t = _as(taps, np.float32)
            super().__init__(lib().oo_fir_iq_from_taps(_ptr(t), t.size))
        else:
            super().__init__(lib().oo_fir_iq_design(num_taps, cutoff_norm, stopband_db))

    taps = FirLowpass.taps

    def group_delay(self):
        return int(lib().oo_fir_iq_group_delay(self._h))

    def filter_aligned(self, io: np.ndarray):
        assert io.dtype == np.complex64 and io.flags.c_contiguous
        lib().oo_fir_iq_filter_aligned(self._h, _ptr(io), io.size)
        return io


class Rotator(Block):
    In = np.complex64
    Out = np.complex64

    def __init__(self, freq_hz, fs):
        super().__init__(lib().oo_rotator_new(freq_hz, fs))

    def set_freq(self, freq_hz, fs):
        lib().oo_rotator_set_freq(self._h, freq_hz, fs)

    def reset_phase(self):
        lib().oo_rotator_reset_phase(self._h)

    def next(self) -> complex:
        p = lib().oo_rotator_next(self._h)
        return complex(p.re, p.im)

    def phasors(self, n) -> np.ndarray:
        out = np.empty(n, np.complex64)
        lib().oo_rotator_phasors(self._h, _ptr(out), n)
        return out

    def rotate_block(self, input) -> np.ndarray:
        x = _as(input, np.complex64)
        out = np.empty_like(x)
        lib().oo_rotator_rotate_block(self._h, _ptr(x), _ptr(out), x.size)
        return out

    def mix_usb_block(self, input) -> np.ndarray:
        x = _as(input, np.complex64)
        out = np.empty(x.size, np.float32)
        lib().oo_rotator_mix_usb_block(self._h, _ptr(x), _ptr(out), x.size)
        return out


class Nco(Block):
    In = np.complex64
    Out = np.complex64

    def __init__(self, freq_hz, fs):
        super().__init__(lib().oo_nco_new(freq_hz, fs))

    def set_freq(self, freq_hz):
        lib().oo_nco_set_freq(self._h, freq_hz)

    def mix(self, input) -> np.ndarray:          # mix_with_nco per sample
        x = _as(input, np.complex64)
        out = np.empty_like(x)
        lib().oo_nco_mix(self._h, _ptr(x), _ptr(out), x.size)
        return out


class Biquad(Block):
    In = np.float32
    Out = np.float32

    def __init__(self, b0, b1, b2, a1, a2):
        super().__init__(lib().oo_biquad_new(b0, b1, b2, a1, a2))


class LpCascade(Block):
    In = np.float32
    Out = np.float32

    def __init__(self, fs, fc):
        super().__init__(lib().oo_lp_cascade_new(fs, fc))


class LpDcCascade(Block):
    In = np.float32
    Out = np.float32

    def __init__(self, fs, lp_fc, dc_cut_hz, map_sqrt=False):
        super().__init__(lib().oo_lp_dc_cascade_new(fs, lp_fc, dc_cut_hz, int(bool(map_sqrt))))


class DcBlocker(Block):
    In = np.float32
    Out = np.float32

    def __init__(self, fs, cut_hz):
        super().__init__(lib().oo_dc_blocker_new(fs, cut_hz))


class FmQuadratureDemod(Block):
    def __init__(self, fs, dev_hz, audio_bw_hz):
        super().__init__(lib().oo_fm_demod_new(fs, dev_hz, audio_bw_hz))

    def with_translate(self, freq_hz):
        lib().oo_fm_demod_with_translate(self._h, freq_hz)
        return self


class PmQuadratureDemod(Block):
    def __init__(self, fs, k, audio_bw_hz):
        super().__init__(lib().oo_pm_demod_new(fs, k, audio_bw_hz))


class AmEnvelopeDemod(Block):
    def __init__(self, fs, audio_bw_hz, abs_approx=False):
        super().__init__(lib().oo_am_demod_new(fs, audio_bw_hz))
        if abs_approx:                    # src/python/demodulate.rs:46-56
            self.with_abs_approx(0.9482, 0.3920)

    def with_abs_approx(self, k1, k2):
        lib().oo_am_demod_with_abs_approx(self._h, k1, k2)
        return self


class SsbProductDemod(Block):
    def __init__(self, fs, bfo_hz, audio_bw_hz):
        super().__init__(lib().oo_ssb_demod_new(fs, bfo_hz, audio_bw_hz))


class CwEnvelopeDemod(Block):
    def __init__(self, sample_rate, tone_hz, env_bw_hz):
        super().__init__(lib().oo_cw_demod_new(sample_rate, tone_hz, env_bw_hz))

    def set_gain(self, g):
        lib().oo_cw_demod_set_gain(self._h, g)


def half_cosine_taps(sps) -> np.ndarray:                # src/dsp/fir.rs:325-346
    n = lib().oo_half_cosine_taps(int(sps), None, 0)
    t = np.zeros(n, np.float32)
    lib().oo_half_cosine_taps(int(sps), _ptr(t), n)
    return t


class HalfCosineMf(Block):                            # src/dsp/fir.rs:317-376 (push per sample)
    In = np.complex64
    Out = np.complex64

    def __init__(self, sps):
        super().__init__(lib().oo_half_cosine_mf_new(int(sps)))


# ---- modulators, f32 -> c32 (src/modulate/*.rs; SURVEY.md 8(f) row 1, CPU oracle only) ----

class _Mod(Block):
    In = np.float32
    Out = np.complex64

    def set_gain(self, g):
        lib().oo_mod_set_gain(self._h, g)


class FmPhaseAccumMod(_Mod):                          # src/modulate/fm.rs:11-75
    def __init__(self, sample_rate, deviation_hz, rf_hz):
        super().__init__(lib().oo_fm_mod_new(sample_rate, deviation_hz, rf_hz))


class PmDirectPhaseMod(_Mod):                         # src/modulate/pm.rs:10-47
    def __init__(self, sample_rate, kp_rad_per_unit, rf_hz):
        super().__init__(lib().oo_pm_mod_new(sample_rate, kp_rad_per_unit, rf_hz))


class AmDsbMod(_Mod):                                 # src/modulate/am.rs:10-120
    def __init__(self, fs, rf_hz, carrier_level, modulation_index):
        super().__init__(lib().oo_am_mod_new(fs, rf_hz, carrier_level, modulation_index))

    def set_clamp(self, on):
        lib().oo_am_mod_set_clamp(self._h, int(bool(on)))


class SsbPhasingMod(_Mod):                            # src/modulate/ssb.rs:11-114
    def __init__(self, fs, audio_bw_hz, audio_if_hz, rf_hz, usb):
        super().__init__(lib().oo_ssb_mod_new(fs, audio_bw_hz, audio_if_hz, rf_hz, int(bool(usb))))


class CwKeyedMod(_Mod):                               # src/modulate/cw.rs:10-102
    def __init__(self, sample_rate, tone_hz, rise_ms, fall_ms):
        super().__init__(lib().oo_cw_mod_new(sample_rate, tone_hz, rise_ms, fall_ms))


class AgcRms(Block):                                  # src/dsp/agc.rs:8-75
    In = np.float32
    Out = np.float32

    def __init__(self, fs, attack_ms, release_ms, target_rms):
        super().__init__(lib().oo_agc_rms_new(fs, attack_ms, release_ms, target_rms))

    @property
    def env(self):
        return float(lib().oo_agc_env(self._h))


class AgcRmsIq(Block):                                # src/dsp/agc.rs:81-150
    In = np.complex64
    Out = np.complex64

    def __init__(self, fs, attack_ms, release_ms, target_rms):
        super().__init__(lib().oo_agc_rms_iq_new(fs, attack_ms, release_ms, target_rms))

    @property
    def env(self):
        return float(lib().oo_agc_env(self._h))


# ---- chain wrappers (src/core.rs:25-109): single-block, return input.len() items ----

class _Chain:
    def __init__(self, block: Block):
        self.block = block
        self.out = np.zeros(0, block.Out)

    def process(self, input) -> np.ndarray:
        return self.process_ref(input)

    def process_ref(self, input) -> np.ndarray:
        x = _as(input, self.block.In)
        if self.out.size < x.size:                       # grow-only scratch, core.rs:99-101
            grown = np.zeros(x.size, self.block.Out)
            grown[: self.out.size] = self.out
            self.out = grown
        n = x.size
        self.block.process_into(x, self.out[:n])
        return self.out[:n].copy()                       # WorkReport ignored, core.rs:103-104

    def process_into(self, input, output) -> WorkReport:
        return self.block.process_into(input, output)


class IqToAudioChain(_Chain):
    pass


class IqToIqChain(_Chain):
    pass


class AudioToIqChain(_Chain):
    pass
