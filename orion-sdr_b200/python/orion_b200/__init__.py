"""orion_b200 -- ctypes binding of liborion_b200.so (include/orion_b200.h).

Host-side mirror of the reference's block interface for the sample-stream front end:
every class below corresponds to one reference `impl Block` (src/core.rs:12-22) and keeps
its constructor arguments and its `process(input, output) -> WorkReport` contract.  All
arithmetic happens in the sm_100a kernels behind the C ABI; there is no CPU fallback --
constructing a block without a CUDA device raises `OrionB200Error`.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))           # .../orion-sdr_b200
LIB_PATH = os.environ.get("ORION_B200_LIB") or os.path.join(_PKG_ROOT, "lib", "liborion_b200.so")

WorkReport = namedtuple("WorkReport", ["in_read", "out_written"])   # src/core.rs:7-10

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_ALLOC, ERR_UNSUPPORTED, ERR_INTERNAL = range(7)
ITEM_F32, ITEM_C32, ITEM_U8 = 1, 2, 3
MIX_NONE, MIX_ROTATE, MIX_NCO = 0, 1, 2
FIR_NONE, FIR_DECIM, FIR_IQ = 0, 1, 2
DEMOD_NONE, DEMOD_FM, DEMOD_PM, DEMOD_AM, DEMOD_AM_ABS, DEMOD_SSB, DEMOD_CW, DEMOD_USB = range(8)
OPT_FIR_GLOBAL, OPT_USE_TMA, OPT_SERIAL_TILES, OPT_OVERLAP_LAUNCHES, OPT_EXACT_NCO = 1, 2, 3, 4, 5


class OrionB200Error(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__(f"orion_b200 status {status}: {msg}")


class _WR(C.Structure):
    _fields_ = [("in_read", C.c_size_t), ("out_written", C.c_size_t)]


class ChainSpec(C.Structure):
    """orion_b200_chain_spec (include/orion_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("mix", C.c_int32), ("mix_freq_hz", C.c_float), ("mix_fs", C.c_float),
        ("fir", C.c_int32), ("taps", C.POINTER(C.c_float)), ("ntaps", C.c_size_t), ("decim", C.c_size_t),
        ("demod", C.c_int32), ("fs_demod", C.c_float), ("p0", C.c_float), ("p1", C.c_float),
        ("audio_bw_hz", C.c_float), ("translate", C.c_int32), ("translate_hz", C.c_float),
        ("post_sos", C.POINTER(C.c_float)), ("n_post", C.c_size_t),
    ]


_lib = None


def lib():
    """Load the C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    f, sz, vp, i, d = C.c_float, C.c_size_t, C.c_void_p, C.c_int, C.c_double
    pp = C.POINTER(vp)
    psz = C.POINTER(sz)

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("orion_b200_abi_version", i)
    sig("orion_b200_build_info", C.c_char_p)
    sig("orion_b200_device_count", i)
    sig("orion_b200_set_device", i, i)
    sig("orion_b200_status_string", C.c_char_p, i)
    sig("orion_b200_host_alloc", i, pp, sz)
    sig("orion_b200_host_free", None, vp)
    sig("orion_b200_fir_lowpass_design", sz, f, f, f, vp, sz)
    sig("orion_b200_kaiser_lowpass_taps", sz, sz, f, f, vp, sz)
    sig("orion_b200_kaiser_transition_norm", f, sz, f)
    sig("orion_b200_kaiser_num_taps", sz, f, f)
    sig("orion_b200_lp_biquad_design", None, f, f, vp)
    sig("orion_b200_dc_pole", f, f, f)
    sig("orion_b200_cw_alpha", f, f, f)
    sig("orion_b200_fir_decimator_create", i, f, sz, f, f, pp)
    sig("orion_b200_fir_decimator_create_taps", i, vp, sz, sz, pp)
    sig("orion_b200_fir_lowpass_iq_create", i, sz, f, f, pp)
    sig("orion_b200_fir_lowpass_iq_create_taps", i, vp, sz, pp)
    sig("orion_b200_fir_lowpass_iq_filter_aligned", i, vp, vp, sz)
    sig("orion_b200_rotator_create", i, f, f, pp)
    sig("orion_b200_rotator_usb_create", i, f, f, pp)
    sig("orion_b200_nco_mixer_create", i, f, f, pp)
    sig("orion_b200_oscillator_set_freq", i, vp, f, f)
    sig("orion_b200_oscillator_reset_phase", i, vp)
    sig("orion_b200_biquad_create", i, f, f, f, f, f, pp)
    sig("orion_b200_lp_cascade_create", i, f, f, pp)
    sig("orion_b200_agc_rms_create", i, f, f, f, f, pp)
    sig("orion_b200_agc_rms_iq_create", i, f, f, f, f, pp)
    sig("orion_b200_agc_env", f, vp)
    sig("orion_b200_bank_set_stream", i, vp, vp)
    sig("orion_b200_fm_mod_create", i, f, f, f, pp)
    sig("orion_b200_fm_mod_set_deviation", i, vp, f)
    sig("orion_b200_cw_mod_create", i, f, f, f, f, pp)
    sig("orion_b200_ssb_mod_create", i, f, f, f, f, i, pp)
    sig("orion_b200_symbol_gain_create", i, f, pp)
    sig("orion_b200_symbol_gain_set", i, vp, f)
    sig("orion_b200_decider_create", i, i, pp)
    sig("orion_b200_cfo_derotate", i, f, f, vp, vp, sz)
    sig("orion_b200_lp_dc_cascade_create", i, f, f, f, i, pp)
    sig("orion_b200_dc_blocker_create", i, f, f, pp)
    sig("orion_b200_iir_cascade_create", i, vp, sz, pp)
    sig("orion_b200_fm_demod_create", i, f, f, f, pp)
    sig("orion_b200_fm_demod_with_translate", i, vp, f)
    sig("orion_b200_pm_demod_create", i, f, f, f, pp)
    sig("orion_b200_am_demod_create", i, f, f, pp)
    sig("orion_b200_am_demod_with_abs_approx", i, vp, f, f)
    sig("orion_b200_ssb_demod_create", i, f, f, f, pp)
    sig("orion_b200_cw_demod_create", i, f, f, f, pp)
    sig("orion_b200_cw_demod_set_gain", i, vp, f)
    sig("orion_b200_chain_create", i, C.POINTER(ChainSpec), pp)
    sig("orion_b200_block_destroy", None, vp)
    sig("orion_b200_block_reset", i, vp)
    sig("orion_b200_block_last_error", C.c_char_p, vp)
    sig("orion_b200_block_in_item", i, vp)
    sig("orion_b200_block_out_item", i, vp)
    sig("orion_b200_block_decimation", sz, vp)
    sig("orion_b200_block_plan", _WR, vp, sz, sz)
    sig("orion_b200_block_process", i, vp, vp, sz, vp, sz, psz, psz)
    sig("orion_b200_block_process_dev", i, vp, vp, sz, vp, sz, psz, psz)
    sig("orion_b200_block_synchronize", i, vp)
    sig("orion_b200_block_set_stream", i, vp, vp)
    sig("orion_b200_block_set_option", i, vp, i, d)
    sig("orion_b200_block_get_state", sz, vp, vp, sz)
    sig("orion_b200_block_launch_count", C.c_uint64, vp)
    sig("orion_b200_block_exact_host_ms", d, vp)
    sig("orion_b200_block_prepare_oscillator", i, vp, sz, sz)
    sig("orion_b200_last_create_error", C.c_char_p)
    sig("orion_b200_block_snapshot_size", sz, vp)
    sig("orion_b200_block_snapshot", i, vp, vp, sz)
    sig("orion_b200_block_restore", i, vp, vp, sz)
    sig("orion_b200_debug_set_trace", i, vp, vp)
    sig("orion_b200_half_cosine_mf_taps", sz, sz, vp, sz)
    sig("orion_b200_half_cosine_mf_create", i, sz, vp)
    sig("orion_b200_am_mod_create", i, f, f, f, f, vp)
    sig("orion_b200_am_mod_set_clamp", i, vp, i)
    sig("orion_b200_pm_mod_create", i, f, f, f, vp)
    sig("orion_b200_mod_set_gain", i, vp, f)
    sig("orion_b200_bank_create", i, vp, sz, vp)
    sig("orion_b200_bank_destroy", None, vp)
    sig("orion_b200_bank_reset", i, vp)
    sig("orion_b200_bank_channels", sz, vp)
    sig("orion_b200_bank_last_error", C.c_char_p, vp)
    sig("orion_b200_bank_process", i, vp, vp, sz, vp, sz, vp, vp)
    sig("orion_b200_bank_process_dev", i, vp, vp, sz, vp, sz, vp, vp)
    sig("orion_b200_bank_synchronize", i, vp)
    sig("orion_b200_bank_launch_count", C.c_uint64, vp)
    sig("orion_b200_debug_fir_plan", sz, i, vp, sz, sz, vp, vp, sz, vp, sz)
    sig("orion_b200_debug_group_tables", sz, vp, sz, i, vp, sz)
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "orion_b200_abi_version", "orion_b200_build_info", "orion_b200_device_count", "orion_b200_set_device",
    "orion_b200_status_string", "orion_b200_host_alloc", "orion_b200_host_free",
    "orion_b200_fir_lowpass_design", "orion_b200_kaiser_lowpass_taps", "orion_b200_kaiser_transition_norm",
    "orion_b200_kaiser_num_taps", "orion_b200_lp_biquad_design", "orion_b200_dc_pole", "orion_b200_cw_alpha",
    "orion_b200_fir_decimator_create", "orion_b200_fir_decimator_create_taps", "orion_b200_fir_lowpass_iq_create",
    "orion_b200_fir_lowpass_iq_create_taps", "orion_b200_fir_lowpass_iq_filter_aligned",
    "orion_b200_rotator_create", "orion_b200_rotator_usb_create", "orion_b200_nco_mixer_create",
    "orion_b200_oscillator_set_freq", "orion_b200_oscillator_reset_phase", "orion_b200_biquad_create",
    "orion_b200_lp_cascade_create", "orion_b200_lp_dc_cascade_create", "orion_b200_dc_blocker_create",
    "orion_b200_iir_cascade_create", "orion_b200_fm_demod_create", "orion_b200_fm_demod_with_translate",
    "orion_b200_pm_demod_create", "orion_b200_am_demod_create", "orion_b200_am_demod_with_abs_approx",
    "orion_b200_ssb_demod_create", "orion_b200_cw_demod_create", "orion_b200_cw_demod_set_gain",
    "orion_b200_chain_create", "orion_b200_block_destroy", "orion_b200_block_reset",
    "orion_b200_block_last_error", "orion_b200_block_in_item", "orion_b200_block_out_item",
    "orion_b200_block_decimation", "orion_b200_block_plan", "orion_b200_block_process",
    "orion_b200_block_process_dev", "orion_b200_block_synchronize", "orion_b200_block_set_stream",
    "orion_b200_block_set_option", "orion_b200_block_get_state", "orion_b200_block_launch_count",
    "orion_b200_debug_fir_plan", "orion_b200_debug_group_tables", "orion_b200_debug_set_trace",
    "orion_b200_half_cosine_mf_taps", "orion_b200_half_cosine_mf_create",
    "orion_b200_block_snapshot_size", "orion_b200_block_snapshot", "orion_b200_block_restore",
    "orion_b200_am_mod_create", "orion_b200_am_mod_set_clamp", "orion_b200_pm_mod_create", "orion_b200_mod_set_gain",
    "orion_b200_bank_create", "orion_b200_bank_destroy", "orion_b200_bank_reset", "orion_b200_bank_channels",
    "orion_b200_bank_last_error", "orion_b200_bank_process", "orion_b200_bank_process_dev",
    "orion_b200_bank_synchronize", "orion_b200_bank_launch_count",
    "orion_b200_block_exact_host_ms", "orion_b200_last_create_error", "orion_b200_block_prepare_oscillator",
    "orion_b200_agc_rms_create", "orion_b200_agc_rms_iq_create", "orion_b200_agc_env", "orion_b200_bank_set_stream",
    "orion_b200_fm_mod_create", "orion_b200_fm_mod_set_deviation", "orion_b200_cw_mod_create", "orion_b200_ssb_mod_create",
    "orion_b200_symbol_gain_create", "orion_b200_symbol_gain_set", "orion_b200_decider_create", "orion_b200_cfo_derotate",
]


def _check(status, handle=None):
    if status != OK:
        L = lib()
        msg = L.orion_b200_status_string(status).decode()
        if handle:
            extra = L.orion_b200_block_last_error(handle)
            if extra:
                msg += " -- " + extra.decode()
        raise OrionB200Error(status, msg)


def device_count() -> int:
    return lib().orion_b200_device_count()


def set_device(ordinal: int) -> None:
    _check(lib().orion_b200_set_device(int(ordinal)))


# ---- design helpers (host-only; reference design math, quirks included) ---------------------------
def fir_lowpass_design(fs, pass_hz, trans_hz) -> np.ndarray:          # src/dsp/fir.rs:16-44
    L = lib()
    n = L.orion_b200_fir_lowpass_design(fs, pass_hz, trans_hz, None, 0)
    t = np.zeros(n, np.float32)
    L.orion_b200_fir_lowpass_design(fs, pass_hz, trans_hz, t.ctypes.data, n)
    return t


def kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db) -> np.ndarray:   # src/dsp/fir.rs:113-141
    L = lib()
    n = L.orion_b200_kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db, None, 0)
    t = np.zeros(n, np.float32)
    L.orion_b200_kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db, t.ctypes.data, n)
    return t


def kaiser_transition_norm(num_taps, stopband_db) -> float:
    return float(lib().orion_b200_kaiser_transition_norm(num_taps, stopband_db))


def kaiser_num_taps(transition_norm, stopband_db) -> int:
    return int(lib().orion_b200_kaiser_num_taps(transition_norm, stopband_db))


def lp_biquad_design(fs, fc) -> np.ndarray:                            # src/dsp/iir.rs:49-71
    c = np.zeros(5, np.float32)
    lib().orion_b200_lp_biquad_design(fs, fc, c.ctypes.data)
    return c


def dc_pole(fs, cut_hz) -> float:
    return float(lib().orion_b200_dc_pole(fs, cut_hz))


def cw_alpha(fs, env_bw_hz) -> float:
    return float(lib().orion_b200_cw_alpha(fs, env_bw_hz))


def debug_fir_plan(fir_kind, taps, m):
    """Host-side launch plan of the staged FIR (tests only)."""
    L = lib()
    taps = np.ascontiguousarray(taps, np.float32)
    info = (C.c_int * 12)()
    g = np.zeros(len(taps), np.float32)
    nf = L.orion_b200_debug_fir_plan(fir_kind, taps.ctypes.data, len(taps), m, info, None, 0, g.ctypes.data, len(g))
    table = np.zeros(max(nf, 1), np.float32)
    L.orion_b200_debug_fir_plan(fir_kind, taps.ctypes.data, len(taps), m, info, table.ctypes.data, nf,
                                g.ctypes.data, len(g))
    keys = ["front", "R", "U", "Mb", "O", "P", "P_pad", "HR", "row_samples", "row_pitch", "rows", "H"]
    plan = dict(zip(keys, list(info)))
    plan["table"] = table[:nf].reshape(-1, 2)
    plan["g"] = g
    return plan


def debug_group_tables(sections, npt):
    """Scan tables of one section group (tests only).  sections = [(type, [coefficients]), ...]."""
    L = lib()
    sec = np.zeros((len(sections), 6), np.float32)
    for q, (t, c) in enumerate(sections):
        sec[q, 0] = t
        sec[q, 1:1 + len(c)] = c
    nf = L.orion_b200_debug_group_tables(sec.ctypes.data, len(sections), npt, None, 0)
    out = np.zeros(nf, np.float32)
    L.orion_b200_debug_group_tables(sec.ctypes.data, len(sections), npt, out.ctypes.data, nf)
    D = int(out[0])
    o = 4
    imp = out[o:o + 16 * 4].reshape(16, 4)[:npt, :D]; o += 16 * 4

    def mats(n):
        nonlocal o
        m = out[o:o + n * 16].reshape(n, 16)[:, :D * D].reshape(n, D, D)
        o += n * 16
        return m
    lv, lane, lb = mats(5), mats(32), mats(32)
    lb32, tile = mats(1)[0], mats(1)[0]
    return {"D": D, "depth": int(out[1]), "agg_only": bool(out[2]), "imp": imp, "lv": lv, "lane": lane, "lb": lb,
            "lb32": lb32, "tile": tile}


_DT = {ITEM_F32: np.float32, ITEM_C32: np.complex64, ITEM_U8: np.uint8}


class Block:
    """One reference `Block` instance living on the GPU (src/core.rs:12-22)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        L = lib()
        self.in_dtype = _DT[L.orion_b200_block_in_item(self._h)]
        self.out_dtype = _DT[L.orion_b200_block_out_item(self._h)]

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib().orion_b200_block_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Block::process ---------------------------------------------------------------------------
    def process(self, inp: np.ndarray, out: np.ndarray) -> WorkReport:
        """`fn process(&mut self, input: &[In], output: &mut [Out]) -> WorkReport` with host slices."""
        if inp.dtype != self.in_dtype or out.dtype != self.out_dtype:
            raise TypeError(f"expected {self.in_dtype.__name__} -> {self.out_dtype.__name__}")
        if inp.ndim != 1 or out.ndim != 1 or not inp.flags.c_contiguous or not out.flags.c_contiguous:
            raise ValueError("input and output must be 1-D contiguous arrays")
        ir, ow = C.c_size_t(0), C.c_size_t(0)
        st = lib().orion_b200_block_process(self._h, inp.ctypes.data if inp.size else None, inp.size,
                                            out.ctypes.data if out.size else None, out.size,
                                            C.byref(ir), C.byref(ow))
        _check(st, self._h)
        return WorkReport(ir.value, ow.value)

    process_into = process                                            # src/core.rs:19-21

    def plan(self, n_in: int, out_cap: int) -> WorkReport:
        wr = lib().orion_b200_block_plan(self._h, n_in, out_cap)
        return WorkReport(wr.in_read, wr.out_written)

    def run(self, inp: np.ndarray) -> np.ndarray:
        """Convenience: size the output like a careful caller would and return the written part."""
        inp = np.ascontiguousarray(inp, self.in_dtype)
        cap = self.plan(inp.size, 1 << 62).out_written
        out = np.zeros(cap, self.out_dtype)
        wr = self.process(inp, out)
        return out[:wr.out_written]

    def process_dev(self, d_in: int, n_in: int, d_out: int, out_cap: int) -> WorkReport:
        """Device-pointer variant: enqueues on the block's stream, does not synchronise."""
        ir, ow = C.c_size_t(0), C.c_size_t(0)
        st = lib().orion_b200_block_process_dev(self._h, C.c_void_p(d_in), n_in, C.c_void_p(d_out), out_cap,
                                                C.byref(ir), C.byref(ow))
        _check(st, self._h)
        return WorkReport(ir.value, ow.value)

    def synchronize(self):
        _check(lib().orion_b200_block_synchronize(self._h), self._h)

    def set_stream(self, cuda_stream: int):
        _check(lib().orion_b200_block_set_stream(self._h, C.c_void_p(cuda_stream)), self._h)

    def reset(self):
        _check(lib().orion_b200_block_reset(self._h), self._h)

    def set_option(self, option: int, value: float):
        _check(lib().orion_b200_block_set_option(self._h, option, float(value)), self._h)

    def state(self) -> np.ndarray:
        n = lib().orion_b200_block_get_state(self._h, None, 0)
        s = np.zeros(n, np.float32)
        lib().orion_b200_block_get_state(self._h, s.ctypes.data, n)
        return s

    def snapshot(self) -> bytes:
        """The block's complete streaming state (the reference's `Clone`)."""
        n = lib().orion_b200_block_snapshot_size(self._h)
        buf = C.create_string_buffer(n)
        _check(lib().orion_b200_block_snapshot(self._h, buf, n), self._h)
        return buf.raw

    def restore(self, blob: bytes) -> None:
        _check(lib().orion_b200_block_restore(self._h, blob, len(blob)), self._h)

    def set_trace(self, d_ptr: int):
        _check(lib().orion_b200_debug_set_trace(self._h, C.c_void_p(d_ptr)), self._h)

    @property
    def decimation(self) -> int:
        return int(lib().orion_b200_block_decimation(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().orion_b200_block_launch_count(self._h))

    @property
    def exact_host_ms(self) -> float:
        """Host time spent walking the oscillator recurrence (exact-replay mode), reported apart from kernel time."""
        return float(lib().orion_b200_block_exact_host_ms(self._h))

    def prepare_oscillator(self, n_in_per_call: int, n_calls: int = 1):
        """Walk the exact-mode oscillator ahead of the stream (enough for `n_calls` calls of `n_in_per_call` items)."""
        _check(lib().orion_b200_block_prepare_oscillator(self._h, n_in_per_call, n_calls), self._h)


def _check_create(status):
    if status != OK:
        L = lib()
        msg = L.orion_b200_status_string(status).decode()
        extra = L.orion_b200_last_create_error()
        if extra:
            msg += " -- " + extra.decode()
        raise OrionB200Error(status, msg)


def _mk(fn_name, *args) -> int:
    h = C.c_void_p(0)
    st = getattr(lib(), fn_name)(*args, C.byref(h))
    _check_create(st)
    return h.value


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


# ---- src/dsp --------------------------------------------------------------------------------------
class FirDecimator(Block):                                            # src/dsp/decim.rs:24-37
    def __init__(self, fs, m, cutoff_hz, trans_hz):
        super().__init__(_mk("orion_b200_fir_decimator_create", fs, int(m), cutoff_hz, trans_hz))

    @classmethod
    def from_taps(cls, taps, m):
        t = _f32(taps)
        self = cls.__new__(cls)
        Block.__init__(self, _mk("orion_b200_fir_decimator_create_taps", t.ctypes.data, t.size, int(m)))
        return self


class FirLowpassIq(Block):                                            # src/dsp/fir.rs:177-297
    def __init__(self, num_taps, cutoff_norm, stopband_db):
        super().__init__(_mk("orion_b200_fir_lowpass_iq_create", int(num_taps), cutoff_norm, stopband_db))
        self._ntaps = int(lib().orion_b200_kaiser_lowpass_taps(int(num_taps), cutoff_norm, stopband_db, None, 0))

    design = classmethod(lambda cls, num_taps, cutoff_norm, stopband_db: cls(num_taps, cutoff_norm, stopband_db))

    @classmethod
    def from_taps(cls, taps):
        t = _f32(taps)
        self = cls.__new__(cls)
        Block.__init__(self, _mk("orion_b200_fir_lowpass_iq_create_taps", t.ctypes.data if t.size else None, t.size))
        self._ntaps = max(int(t.size), 1)
        return self

    def group_delay(self) -> int:                                     # src/dsp/fir.rs:216-218
        return (self._ntaps - 1) // 2

    def filter_aligned(self, io: np.ndarray) -> None:                 # src/dsp/fir.rs:260-276 (in place)
        if io.dtype != np.complex64 or io.ndim != 1 or not io.flags.c_contiguous:
            raise TypeError("filter_aligned needs a contiguous complex64 vector")
        _check(lib().orion_b200_fir_lowpass_iq_filter_aligned(self._h, io.ctypes.data if io.size else None, io.size),
               self._h)


def half_cosine_mf_taps(sps) -> np.ndarray:                           # src/dsp/fir.rs:325-346
    n = lib().orion_b200_half_cosine_mf_taps(int(sps), None, 0)
    t = np.zeros(n, np.float32)
    lib().orion_b200_half_cosine_mf_taps(int(sps), t.ctypes.data, n)
    return t


class HalfCosineMf(Block):                                            # src/dsp/fir.rs:317-376, one push() per sample
    def __init__(self, sps):
        super().__init__(_mk("orion_b200_half_cosine_mf_create", int(sps)))


class Rotator(Block):                                                 # src/dsp/rotator.rs (rotate_block)
    def __init__(self, freq_hz, fs):
        super().__init__(_mk("orion_b200_rotator_create", freq_hz, fs))

    def set_freq(self, freq_hz, fs):
        _check(lib().orion_b200_oscillator_set_freq(self._h, freq_hz, fs), self._h)

    def reset_phase(self):
        _check(lib().orion_b200_oscillator_reset_phase(self._h), self._h)

    rotate_block = Block.process


class RotatorUsb(Block):                                              # Rotator::mix_usb_block, rotator.rs:88-94
    def __init__(self, freq_hz, fs):
        super().__init__(_mk("orion_b200_rotator_usb_create", freq_hz, fs))

    def set_freq(self, freq_hz, fs):
        _check(lib().orion_b200_oscillator_set_freq(self._h, freq_hz, fs), self._h)

    mix_usb_block = Block.process


class NcoMixer(Block):                                                # src/dsp/nco.rs (mix_with_nco per sample)
    def __init__(self, freq_hz, fs):
        super().__init__(_mk("orion_b200_nco_mixer_create", freq_hz, fs))
        self._fs = fs

    def set_freq(self, freq_hz):
        _check(lib().orion_b200_oscillator_set_freq(self._h, freq_hz, self._fs), self._h)


class Biquad(Block):                                                  # src/dsp/iir.rs:5-41
    def __init__(self, b0, b1, b2, a1, a2):
        super().__init__(_mk("orion_b200_biquad_create", b0, b1, b2, a1, a2))


class LpCascade(Block):                                               # src/dsp/iir.rs:44-84
    def __init__(self, fs, fc):
        super().__init__(_mk("orion_b200_lp_cascade_create", fs, fc))

    design = classmethod(lambda cls, fs, fc: cls(fs, fc))


class LpDcCascade(Block):                                             # src/dsp/iir.rs:90-187
    def __init__(self, fs, lp_fc, dc_cut_hz, map_sqrt=False):
        super().__init__(_mk("orion_b200_lp_dc_cascade_create", fs, lp_fc, dc_cut_hz, int(bool(map_sqrt))))

    design = classmethod(lambda cls, fs, lp_fc, dc_cut_hz: cls(fs, lp_fc, dc_cut_hz))


class DcBlocker(Block):                                               # src/dsp/dc.rs:8-59
    def __init__(self, fs, cut_hz):
        super().__init__(_mk("orion_b200_dc_blocker_create", fs, cut_hz))


class IirCascade(Block):                                              # N x Biquad::process
    def __init__(self, sos):
        s = _f32(sos).reshape(-1, 5)
        super().__init__(_mk("orion_b200_iir_cascade_create", s.ctypes.data, s.shape[0]))


# ---- src/demodulate -------------------------------------------------------------------------------
class FmQuadratureDemod(Block):                                       # src/demodulate/fm.rs:11-78
    def __init__(self, fs, dev_hz, audio_bw_hz):
        super().__init__(_mk("orion_b200_fm_demod_create", fs, dev_hz, audio_bw_hz))

    def with_translate(self, freq_hz):
        _check(lib().orion_b200_fm_demod_with_translate(self._h, freq_hz), self._h)
        return self


class PmQuadratureDemod(Block):                                       # src/demodulate/pm.rs:12-67
    def __init__(self, fs, k, audio_bw_hz):
        super().__init__(_mk("orion_b200_pm_demod_create", fs, k, audio_bw_hz))


class AmEnvelopeDemod(Block):                                         # src/demodulate/am.rs:10-130
    def __init__(self, fs, audio_bw_hz):
        super().__init__(_mk("orion_b200_am_demod_create", fs, audio_bw_hz))

    def with_abs_approx(self, k1, k2):
        _check(lib().orion_b200_am_demod_with_abs_approx(self._h, k1, k2), self._h)
        return self


class SsbProductDemod(Block):                                         # src/demodulate/ssb.rs:9-72
    def __init__(self, fs, bfo_hz, audio_bw_hz):
        super().__init__(_mk("orion_b200_ssb_demod_create", fs, bfo_hz, audio_bw_hz))


class CwEnvelopeDemod(Block):                                         # src/demodulate/cw.rs:8-47
    def __init__(self, sample_rate, tone_hz, env_bw_hz):
        super().__init__(_mk("orion_b200_cw_demod_create", sample_rate, tone_hz, env_bw_hz))

    def set_gain(self, g):
        _check(lib().orion_b200_cw_demod_set_gain(self._h, g), self._h)


# ---- src/dsp/agc.rs (next-row scope) ----------------------------------------------------------------
class AgcRms(Block):                                                  # src/dsp/agc.rs:8-75
    def __init__(self, fs, attack_ms, release_ms, target_rms):
        super().__init__(_mk("orion_b200_agc_rms_create", fs, attack_ms, release_ms, target_rms))

    @property
    def env(self) -> float:
        return float(lib().orion_b200_agc_env(self._h))


class AgcRmsIq(Block):                                                # src/dsp/agc.rs:81-150
    def __init__(self, fs, attack_ms, release_ms, target_rms):
        super().__init__(_mk("orion_b200_agc_rms_iq_create", fs, attack_ms, release_ms, target_rms))

    @property
    def env(self) -> float:
        return float(lib().orion_b200_agc_env(self._h))


# ---- src/modulate (next-row scope) ------------------------------------------------------------------
class AmDsbMod(Block):                                                # src/modulate/am.rs:10-120
    def __init__(self, fs, rf_hz, carrier_level, modulation_index):
        super().__init__(_mk("orion_b200_am_mod_create", fs, rf_hz, carrier_level, modulation_index))

    def set_clamp(self, on):
        _check(lib().orion_b200_am_mod_set_clamp(self._h, int(bool(on))), self._h)

    def set_gain(self, g):
        _check(lib().orion_b200_mod_set_gain(self._h, g), self._h)


class PmDirectPhaseMod(Block):                                        # src/modulate/pm.rs:10-47
    def __init__(self, sample_rate, kp_rad_per_unit, rf_hz):
        super().__init__(_mk("orion_b200_pm_mod_create", sample_rate, kp_rad_per_unit, rf_hz))

    def set_gain(self, g):
        _check(lib().orion_b200_mod_set_gain(self._h, g), self._h)


class FmPhaseAccumMod(Block):                                         # src/modulate/fm.rs:11-75
    def __init__(self, sample_rate, deviation_hz, rf_hz):
        super().__init__(_mk("orion_b200_fm_mod_create", sample_rate, deviation_hz, rf_hz))

    def set_deviation(self, deviation_hz):
        _check(lib().orion_b200_fm_mod_set_deviation(self._h, deviation_hz), self._h)

    def set_gain(self, g):
        _check(lib().orion_b200_mod_set_gain(self._h, g), self._h)


class CwKeyedMod(Block):                                              # src/modulate/cw.rs:10-102
    def __init__(self, sample_rate, tone_hz, rise_ms, fall_ms):
        super().__init__(_mk("orion_b200_cw_mod_create", sample_rate, tone_hz, rise_ms, fall_ms))

    def set_gain(self, g):
        _check(lib().orion_b200_mod_set_gain(self._h, g), self._h)


class SsbPhasingMod(Block):                                           # src/modulate/ssb.rs:11-114
    def __init__(self, fs, audio_bw_hz, audio_if_hz, rf_hz, usb):
        super().__init__(_mk("orion_b200_ssb_mod_create", fs, audio_bw_hz, audio_if_hz, rf_hz, int(bool(usb))))


# ---- src/demodulate/{bpsk,qpsk,qam}.rs (next-row scope): soft-symbol gain and hard-decision slicers -------------
class _SymbolGain(Block):
    def __init__(self, gain):
        super().__init__(_mk("orion_b200_symbol_gain_create", gain))

    def set_gain(self, g):
        _check(lib().orion_b200_symbol_gain_set(self._h, g), self._h)


class BpskDemod(_SymbolGain):                                         # src/demodulate/bpsk.rs:14-51
    pass


class QpskDemod(_SymbolGain):                                         # src/demodulate/qpsk.rs:13-50
    pass


class QamDemod(_SymbolGain):                                          # src/demodulate/qam.rs:44-82
    pass


class _Decider(Block):
    bits = 1

    def __init__(self):
        super().__init__(_mk("orion_b200_decider_create", self.bits))


class BpskDecider(_Decider):                                          # src/demodulate/bpsk.rs:54-88
    bits = 1


class QpskDecider(_Decider):                                          # src/demodulate/qpsk.rs:53-98
    bits = 2


class Qam16Decider(_Decider):                                         # src/demodulate/qam.rs:88-186
    bits = 4


class Qam64Decider(_Decider):
    bits = 6


class Qam256Decider(_Decider):
    bits = 8


def cfo_derotate(iq: np.ndarray, cfo_hz: float, fs: float) -> np.ndarray:
    """Rotator::new(-cfo_hz, fs).rotate_block(iq, out) on the GPU (sync/ofdm_sync.rs:527-528)."""
    iq = np.ascontiguousarray(iq, np.complex64)
    out = np.empty_like(iq)
    _check(lib().orion_b200_cfo_derotate(cfo_hz, fs, iq.ctypes.data if iq.size else None, out.ctypes.data if out.size else None, iq.size))
    return out


# ---- fused chain ----------------------------------------------------------------------------------
class Chain(Block):
    """[mixer] -> [FIR, decimate] -> [demod] -> [post biquads] in one kernel (orion_b200_chain_create)."""

    def __init__(self, *, mix=MIX_NONE, mix_freq_hz=0.0, mix_fs=1.0, fir=FIR_NONE, taps=None, decim=1,
                 demod=DEMOD_NONE, fs_demod=1.0, p0=0.0, p1=0.0, audio_bw_hz=0.0, translate_hz=None,
                 post_sos=None):
        sp = ChainSpec()
        sp.struct_size = C.sizeof(ChainSpec)
        sp.mix, sp.mix_freq_hz, sp.mix_fs = mix, mix_freq_hz, mix_fs
        sp.fir = fir
        self._taps = _f32(taps) if taps is not None else np.zeros(0, np.float32)
        sp.taps = self._taps.ctypes.data_as(C.POINTER(C.c_float)) if self._taps.size else None
        sp.ntaps, sp.decim = self._taps.size, int(decim)
        sp.demod, sp.fs_demod, sp.p0, sp.p1, sp.audio_bw_hz = demod, fs_demod, p0, p1, audio_bw_hz
        sp.translate = int(translate_hz is not None)
        sp.translate_hz = float(translate_hz or 0.0)
        self._sos = _f32(post_sos).reshape(-1, 5) if post_sos is not None else np.zeros((0, 5), np.float32)
        sp.post_sos = self._sos.ctypes.data_as(C.POINTER(C.c_float)) if self._sos.size else None
        sp.n_post = self._sos.shape[0]
        h = C.c_void_p(0)
        _check_create(lib().orion_b200_chain_create(C.byref(sp), C.byref(h)))
        super().__init__(h.value)


def _fill_spec(sp, keep, *, mix=MIX_NONE, mix_freq_hz=0.0, mix_fs=1.0, fir=FIR_NONE, taps=None, decim=1, demod=DEMOD_NONE,
               fs_demod=1.0, p0=0.0, p1=0.0, audio_bw_hz=0.0, translate_hz=None, post_sos=None):
    """Fill one orion_b200_chain_spec; `keep` collects the arrays its pointers refer to."""
    sp.struct_size = C.sizeof(ChainSpec)
    sp.mix, sp.mix_freq_hz, sp.mix_fs = mix, mix_freq_hz, mix_fs
    sp.fir = fir
    t = _f32(taps) if taps is not None else np.zeros(0, np.float32)
    sos = _f32(post_sos).reshape(-1, 5) if post_sos is not None else np.zeros((0, 5), np.float32)
    keep += [t, sos]
    sp.taps = t.ctypes.data_as(C.POINTER(C.c_float)) if t.size else None
    sp.ntaps, sp.decim = t.size, int(decim)
    sp.demod, sp.fs_demod, sp.p0, sp.p1, sp.audio_bw_hz = demod, fs_demod, p0, p1, audio_bw_hz
    sp.translate = int(translate_hz is not None)
    sp.translate_hz = float(translate_hz or 0.0)
    sp.post_sos = sos.ctypes.data_as(C.POINTER(C.c_float)) if sos.size else None
    sp.n_post = sos.shape[0]


# ---- channel bank (BASELINE config 5) and its sharding across processes ---------------------------------
def pin_host_to_device_numa_node(ordinal: int) -> dict:
    """Bind the calling thread (and the threads it starts later: the library's staging pool) to the CPUs of the NUMA node
    the GPU hangs off, so that host buffers allocated from here on are first-touched in that node's memory.  With one
    process per GPU this keeps every rank's host<->device copies off the inter-socket link (8 ranks on one node shared
    one socket's memory bandwidth: e2e scaled 3.4x from 1 to 8 GPUs).  Best effort: returns what it found and did, never
    raises (containers without /sys NUMA information are left alone)."""
    info = {"node": None, "cpus": 0, "pinned": False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(ordinal)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
            info["pinned"] = True
    except Exception as e:                                   # noqa: BLE001 -- best effort by design
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def shard_range(n_channels: int, rank: int, world: int) -> range:
    """Channels [g*C/G, (g+1)*C/G) of rank g: contiguous, covers every channel once."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return range(rank * n_channels // world, (rank + 1) * n_channels // world)


def gather_channels(local: np.ndarray, n_channels: int, rank: int, world: int, group=None, dst: int = 0):
    """Host-side gather of the per-rank [local_channels, n_out] outputs into [n_channels, n_out] on `dst`
    (None elsewhere).  torch.distributed carries host tensors only (gloo, or the CPU side of a mixed
    group): there is no device collective on the data path."""
    if world == 1:
        return np.ascontiguousarray(local)
    import torch
    import torch.distributed as dist
    mine = shard_range(n_channels, rank, world)
    assert local.shape[0] == len(mine), (local.shape, mine)
    n_out = local.shape[1] if local.ndim > 1 else 0
    biggest = max(len(shard_range(n_channels, r, world)) for r in range(world))
    buf = torch.zeros((biggest, n_out), dtype=torch.from_numpy(np.zeros(0, local.dtype)).dtype)
    buf[:len(mine)] = torch.from_numpy(np.ascontiguousarray(local))
    outs = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst, group=group)
    if rank != dst:
        return None
    full = np.zeros((n_channels, n_out), local.dtype)
    for r in range(world):
        rr = shard_range(n_channels, r, world)
        full[rr.start:rr.stop] = outs[r][:len(rr)].numpy()
    return full


class ChannelBank:
    """C independent narrowband chains on one wideband input (orion_b200_bank_*): channel c is the block
    `Chain(**specs[c])`.  `channels` selects a sub-range of the specs -- one bank per GPU / process."""

    def __init__(self, specs, channels=None):
        specs = list(specs)
        self.channel_ids = list(channels) if channels is not None else list(range(len(specs)))
        arr = (ChainSpec * len(self.channel_ids))()
        self._keep = []
        for i, c in enumerate(self.channel_ids):
            _fill_spec(arr[i], self._keep, **specs[c])
        h = C.c_void_p(0)
        st = lib().orion_b200_bank_create(arr, len(self.channel_ids), C.byref(h))
        _check(st)
        self._h = h
        self.decim = max(int(specs[self.channel_ids[0]].get("decim", 1)), 1) if specs[self.channel_ids[0]].get("fir", FIR_NONE) != FIR_NONE else 1
        self.out_dtype = np.complex64 if specs[self.channel_ids[0]].get("demod", DEMOD_NONE) == DEMOD_NONE else np.float32

    def _ck(self, st):
        if st != OK:
            raise OrionB200Error(st, lib().orion_b200_status_string(st).decode() + " -- " +
                                 (lib().orion_b200_bank_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().orion_b200_bank_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(lib().orion_b200_bank_channels(self._h))

    def out_items(self, n_in: int) -> int:
        return -(-n_in // self.decim)

    def process(self, x: np.ndarray) -> np.ndarray:
        """Host slices: one wideband input -> [channels, ceil(n/m)] outputs."""
        x = np.ascontiguousarray(x, np.complex64)
        n_out = self.out_items(x.size)
        out = np.zeros((len(self), n_out), self.out_dtype)
        ir, ow = C.c_size_t(0), C.c_size_t(0)
        self._ck(lib().orion_b200_bank_process(self._h, x.ctypes.data if x.size else None, x.size,
                                               out.ctypes.data if out.size else None, n_out, C.byref(ir), C.byref(ow)))
        assert ow.value == n_out or x.size == 0
        return out

    def process_dev(self, d_in: int, n_in: int, d_out: int, out_stride: int) -> WorkReport:
        ir, ow = C.c_size_t(0), C.c_size_t(0)
        self._ck(lib().orion_b200_bank_process_dev(self._h, C.c_void_p(d_in), n_in, C.c_void_p(d_out), out_stride,
                                                   C.byref(ir), C.byref(ow)))
        return WorkReport(ir.value, ow.value)

    def synchronize(self):
        self._ck(lib().orion_b200_bank_synchronize(self._h))

    def set_stream(self, cuda_stream: int):
        """Run the bank on a caller's CUDA stream (0: back to its own)."""
        self._ck(lib().orion_b200_bank_set_stream(self._h, C.c_void_p(cuda_stream)))

    def reset(self):
        self._ck(lib().orion_b200_bank_reset(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().orion_b200_bank_launch_count(self._h))


# ---- src/core.rs chain wrappers --------------------------------------------------------------------
class _ChainWrapper:
    """IqToIqChain / IqToAudioChain / AudioToIqChain (src/core.rs:25-109): one block, a grow-only
    scratch, and `process` returns `input.len()` items whatever the block reported."""

    def __init__(self, block: Block):
        self.block = block
        self._out = np.zeros(0, block.out_dtype)

    def process(self, inp) -> np.ndarray:
        return self.process_ref(np.ascontiguousarray(inp, self.block.in_dtype))

    def process_ref(self, inp: np.ndarray) -> np.ndarray:
        n = inp.size
        if self._out.size < n:                                        # src/core.rs:99-101
            grown = np.zeros(n, self.block.out_dtype)
            grown[:self._out.size] = self._out
            self._out = grown
        self.block.process_into(inp, self._out[:n])
        return self._out[:n].copy()                                   # src/core.rs:104 (`to_vec`)

    def process_into(self, inp: np.ndarray, out: np.ndarray) -> WorkReport:
        return self.block.process_into(inp, out)


IqToIqChain = _ChainWrapper
IqToAudioChain = _ChainWrapper
AudioToIqChain = _ChainWrapper
