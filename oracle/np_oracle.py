"""Second, independent restatement of the reference arithmetic in numpy float32.

TEST INFRASTRUCTURE ONLY.  Purpose: pin oracle/orion_oracle.c.  It is written from
the reference sources a second time with a different program structure (direct
convolution over a zero-padded history instead of circular buffers, whole-array
f32 ops instead of per-sample loops where the arithmetic allows), so an indexing or
FMA-placement mistake in either restatement shows up as a bit mismatch in
tests/test_oracle_crosscheck.py.  One-shot (from reset state) only; small sizes.

f32 semantics: every numpy op on float32 arrays/scalars is one correctly rounded
IEEE op.  fma32() is an exactly rounded fused multiply-add (product exact in f64,
TwoSum error term used to resolve f64->f32 double-rounding ties).  Design-time
transcendental functions call glibc's sinf/cosf/expf/powf through ctypes (what
Rust's f32::sin etc. resolve to on linux-gnu), not numpy's SIMD kernels.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util

import numpy as np

f32 = np.float32
_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("sinf", "cosf", "expf", "sqrtf"):
    getattr(_libm, _n).restype = C.c_float
    getattr(_libm, _n).argtypes = [C.c_float]
_libm.powf.restype = C.c_float
_libm.powf.argtypes = [C.c_float, C.c_float]

PI = f32(np.pi)
TAU = f32(2 * np.pi)
EPS = f32(1.1920929e-07)


def sinf(x): return f32(_libm.sinf(float(f32(x))))
def cosf(x): return f32(_libm.cosf(float(f32(x))))
def expf(x): return f32(_libm.expf(float(f32(x))))
def powf(x, y): return f32(_libm.powf(float(f32(x)), float(f32(y))))


def fma32(a, b, c):
    """Exactly rounded f32 fma for scalars or arrays of float32."""
    a64 = np.asarray(a, np.float64)
    b64 = np.asarray(b, np.float64)
    c64 = np.asarray(c, np.float64)
    p = a64 * b64                     # exact: 24+24 significant bits
    s = p + c64                       # one f64 rounding
    bb = s - p
    err = (p - (s - bb)) + (c64 - bb)  # TwoSum: s + err == p + c exactly
    r = s.astype(np.float32)
    # fix double rounding: s exactly halfway between two f32 values but true sum is not
    r64 = r.astype(np.float64)
    lo = np.nextafter(r, f32(-np.inf)).astype(np.float64)
    hi = np.nextafter(r, f32(np.inf)).astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        tie_lo = (s == (r64 + lo) * 0.5) & (err != 0)
        tie_hi = (s == (r64 + hi) * 0.5) & (err != 0)
    r = np.where(tie_lo & (err < 0), lo.astype(np.float32), r)
    r = np.where(tie_hi & (err > 0), hi.astype(np.float32), r)
    return r.astype(np.float32) if isinstance(r, np.ndarray) and r.ndim else f32(r)


# ---- design ---------------------------------------------------------------------

def fir_lowpass_taps(fs, pass_hz, trans_hz):
    """src/dsp/fir.rs:16-44"""
    fs, pass_hz, trans_hz = f32(fs), f32(pass_hz), f32(trans_hz)
    pass_hz = max(pass_hz, f32(10.0))
    trans_hz = max(trans_hz, f32(pass_hz * f32(0.2)))
    ntaps = max(int(np.ceil(f32(fs / trans_hz))), 31) | 1
    fc = f32(pass_hz / fs)
    m0 = ntaps // 2
    taps = np.zeros(ntaps, f32)
    for n in range(ntaps):
        m = n - m0
        if m == 0:
            sinc = f32(f32(2.0) * fc)
        else:
            x = f32(PI * f32(m))
            arg = f32(f32(f32(f32(2.0) * PI) * fc) * f32(m))
            sinc = f32(f32(f32(f32(2.0) * fc) * sinf(arg)) / x)
        warg = f32(f32(f32(f32(2.0) * PI) * f32(n)) / f32(f32(ntaps) - f32(1.0)))
        w = f32(f32(0.5) - f32(f32(0.5) * cosf(warg)))
        taps[n] = f32(sinc * w)
    s = f32(0.0)
    for t in taps:
        s = f32(s + t)
    return (taps / s).astype(f32)


def kaiser_beta(a_db):
    a_db = f32(a_db)
    if a_db > 50.0:
        return f32(f32(0.1102) * f32(a_db - f32(8.7)))
    if a_db >= 21.0:
        d = f32(a_db - f32(21.0))
        return f32(f32(f32(0.5842) * powf(d, 0.4)) + f32(f32(0.07886) * d))
    return f32(0.0)


def bessel_i0(x):
    half = f32(f32(0.5) * f32(x))
    term = f32(1.0)
    s = f32(1.0)
    for k in range(1, 41):
        term = f32(term * f32(half / f32(k)))
        t = f32(term * term)
        s = f32(s + t)
        if t < f32(f32(1e-12) * s):
            break
    return s


def kaiser_lowpass_taps(num_taps, cutoff_norm, stopband_db):
    """src/dsp/fir.rs:113-141"""
    m = max(int(num_taps), 3) | 1
    mid = f32(m // 2)
    fc = min(max(f32(cutoff_norm), f32(1e-4)), f32(0.4999))
    beta = kaiser_beta(stopband_db)
    i0b = bessel_i0(beta)
    taps = np.zeros(m, f32)
    for n in range(m):
        d = f32(f32(n) - mid)
        if d == 0.0:
            ideal = f32(f32(2.0) * fc)
        else:
            ideal = f32(sinf(f32(f32(TAU * fc) * d)) / f32(PI * d))
        r = f32(d / mid)
        inner = max(f32(f32(1.0) - f32(r * r)), f32(0.0))
        w = f32(bessel_i0(f32(beta * f32(np.sqrt(inner)))) / i0b)
        taps[n] = f32(ideal * w)
    s = f32(0.0)
    for t in taps:
        s = f32(s + t)
    if abs(s) > EPS:
        taps = (taps / s).astype(f32)
    return taps


def lp_biquad_coeffs(fs, fc):
    """src/dsp/iir.rs:49-71 -> b0,b1,b2,a1,a2"""
    w0 = f32(f32(TAU * f32(fc)) / f32(fs))
    sn, cs = sinf(w0), cosf(w0)
    alpha = f32(sn / f32(f32(2.0) * f32(np.sqrt(f32(0.5)))))
    omc = f32(f32(1.0) - cs)
    b0 = f32(omc * f32(0.5)); b1 = omc; b2 = b0
    a0 = f32(f32(1.0) + alpha); a1 = f32(f32(-2.0) * cs); a2 = f32(f32(1.0) - alpha)
    norm = f32(f32(1.0) / a0)
    return np.array([b0 * norm, b1 * norm, b2 * norm, a1 * norm, a2 * norm], f32)


def dc_pole(fs, cut_hz):
    r = f32(f32(1.0) - f32(f32(f32(2.0) * PI) * f32(max(f32(cut_hz), f32(0.1)) / f32(fs))))
    return min(max(r, f32(0.0)), f32(0.9999))


def cw_alpha(fs, bw):
    return expf(f32(f32(-TAU * max(f32(bw), f32(1.0))) / f32(fs)))


def rotator_w(freq_hz, fs):
    phi = f32(f32(TAU * f32(freq_hz)) / f32(fs))
    return cosf(phi), sinf(phi)


# ---- per-sample primitives --------------------------------------------------------

def atan2_approx(y, x):
    """src/util.rs:305-322, vectorised (each op one f32 rounding)."""
    y = np.asarray(y, f32); x = np.asarray(x, f32)
    ax, ay = np.abs(x), np.abs(y)
    swap = ax < ay
    mn = np.where(swap, ax, ay); mx = np.where(swap, ay, ax)
    r = (mn / (mx + EPS)).astype(f32)
    r2 = (r * r).astype(f32)
    inner = (f32(-0.2447) + (r2 * f32(0.0663)).astype(f32)).astype(f32)
    phi = (r * (f32(0.78539816) + (r2 * inner).astype(f32)).astype(f32)).astype(f32)
    phi = np.where(swap, (f32(1.5707964) - phi).astype(f32), phi)
    sgn = np.where(y < 0, f32(-1.0), f32(1.0))
    return np.where(x < 0, ((f32(3.1415927) - phi).astype(f32) * sgn).astype(f32),
                    (phi * sgn).astype(f32)).astype(f32)


def rotator_phasors(freq_hz, fs, n):
    """n successive Rotator::next() values from reset (src/dsp/rotator.rs:44-61)."""
    wr, wi = rotator_w(freq_hz, fs)
    zr, zi = f32(1.0), f32(0.0)
    out = np.zeros(n, np.complex64)
    for i in range(n):
        nzr = fma32(zr, wr, -f32(zi * wi))
        nzi = fma32(zi, wr, f32(zr * wi))
        zr, zi = nzr, nzi
        if ((i + 1) & 0x3FF) == 0:
            r2 = f32(f32(zr * zr) + f32(zi * zi))
            inv = f32(f32(1.0) / f32(np.sqrt(r2)))
            zr = f32(zr * inv); zi = f32(zi * inv)
        out[i] = complex(zr, zi)
    return out


def rotate_block(x, p):
    a = x.real.astype(f32); b = x.imag.astype(f32)
    pr = p.real.astype(f32); pi = p.imag.astype(f32)
    re = fma32(a, pr, -(b * pi).astype(f32))
    im = fma32(b, pr, (a * pi).astype(f32))
    return (re + 1j * im).astype(np.complex64)


def mix_usb(x, p):
    return fma32(x.real.astype(f32), p.real.astype(f32),
                 (x.imag.astype(f32) * p.imag.astype(f32)).astype(f32))


def nco_mix(x, p):
    a = x.real.astype(f32); b = x.imag.astype(f32)
    c = p.real.astype(f32); s = p.imag.astype(f32)
    re = ((a * c).astype(f32) - (b * s).astype(f32)).astype(f32)
    im = ((a * s).astype(f32) + (b * c).astype(f32)).astype(f32)
    return (re + 1j * im).astype(np.complex64)


# ---- FIR ----------------------------------------------------------------------------

def fir_lowpass_process(taps, x):
    """FirLowpass::process from reset (src/dsp/fir.rs:47-66) as a direct convolution:
    y[n] = sum_{t=0}^{L-2} taps[t]*x[n-1-t]  (+ taps[L-1]*x[n] last), unfused, ascending t."""
    taps = np.asarray(taps, f32); x = np.asarray(x, f32)
    L = taps.size; n = x.size
    xp = np.concatenate([np.zeros(L, f32), x])
    acc = np.zeros(n, f32)
    for t in range(L):
        lag = 0 if t == L - 1 else t + 1
        seg = xp[L - lag: L - lag + n]
        acc = (acc + (seg * taps[t]).astype(f32)).astype(f32)
    return acc


def fir_decimator_process(taps, m, x):
    """FirDecimator::process from reset (src/dsp/decim.rs:44-76)."""
    yi = fir_lowpass_process(taps, x.real.astype(f32))
    yq = fir_lowpass_process(taps, x.imag.astype(f32))
    return (yi[::m] + 1j * yq[::m]).astype(np.complex64)


def fir_iq_process(taps, x):
    """FirLowpassIq::push stream from reset (src/dsp/fir.rs:229-247):
    y[n] = sum_j taps[j]*x[n-j], fused, ascending j."""
    taps = np.asarray(taps, f32)
    L = taps.size; n = x.size
    xr = np.concatenate([np.zeros(L, f32), x.real.astype(f32)])
    xi = np.concatenate([np.zeros(L, f32), x.imag.astype(f32)])
    re = np.zeros(n, f32); im = np.zeros(n, f32)
    for j in range(L):
        re = fma32(xr[L - j: L - j + n], taps[j], re)
        im = fma32(xi[L - j: L - j + n], taps[j], im)
    return (re + 1j * im).astype(np.complex64)


# ---- IIR ----------------------------------------------------------------------------

def biquad_run(c, x, z=(0.0, 0.0)):
    b0, b1, b2, a1, a2 = (f32(v) for v in c)
    z1, z2 = f32(z[0]), f32(z[1])
    y = np.zeros(len(x), f32)
    for i, xi in enumerate(np.asarray(x, f32)):
        yi = fma32(xi, b0, z1)
        z1 = f32(fma32(xi, b1, z2) - f32(a1 * yi))
        z2 = f32(f32(xi * b2) - f32(a2 * yi))
        y[i] = yi
    return y


def dc_run(r, x):
    r = f32(r); x1 = f32(0.0); y1 = f32(0.0)
    y = np.zeros(len(x), f32)
    for i, xi in enumerate(np.asarray(x, f32)):
        yi = f32(f32(xi - x1) + f32(r * y1))
        x1 = xi; y1 = yi
        y[i] = yi
    return y


def lp_cascade_run(fs, fc, x):
    c = lp_biquad_coeffs(fs, fc)
    return biquad_run(c, biquad_run(c, x))


def lp_dc_run(fs, lp_fc, dc_cut, x, map_sqrt=False):
    y = lp_cascade_run(fs, lp_fc, x)
    if map_sqrt:
        with np.errstate(invalid="ignore"):
            y = np.sqrt(y).astype(f32)
    return dc_run(dc_pole(fs, dc_cut), y)


# ---- demodulators ----------------------------------------------------------------------

def _disc(z, prev0=1 + 0j):
    zr = z.real.astype(f32); zi = z.imag.astype(f32)
    pr = np.concatenate([[f32(prev0.real)], zr[:-1]]).astype(f32)
    pi = np.concatenate([[f32(prev0.imag)], zi[:-1]]).astype(f32)
    re = ((zr * pr).astype(f32) + (zi * pi).astype(f32)).astype(f32)
    im = ((zi * pr).astype(f32) - (zr * pi).astype(f32)).astype(f32)
    return atan2_approx(im, re)


def fm_demod(fs, dev_hz, audio_bw, x, translate_hz=None):
    """src/demodulate/fm.rs:22-77 from reset."""
    k = f32(f32(1.0) / max(f32(dev_hz), f32(1.0)))
    z = np.asarray(x, np.complex64)
    if translate_hz is not None:
        p = rotator_phasors(translate_hz, fs, z.size)
        a = z.real.astype(f32); b = z.imag.astype(f32)
        c = p.real.astype(f32); s = p.imag.astype(f32)
        re = ((a * c).astype(f32) + (b * s).astype(f32)).astype(f32)
        im = ((b * c).astype(f32) - (a * s).astype(f32)).astype(f32)
        z = (re + 1j * im).astype(np.complex64)
    d = (_disc(z) * k).astype(f32)
    return lp_cascade_run(fs, f32(f32(audio_bw) * f32(0.9)), d)


def pm_demod(fs, k, audio_bw, x):
    d = (f32(k) * _disc(np.asarray(x, np.complex64))).astype(f32)
    return lp_cascade_run(fs, f32(f32(audio_bw) * f32(0.9)), d)


def am_demod(fs, audio_bw, x, abs_approx=None):
    x = np.asarray(x, np.complex64)
    re = x.real.astype(f32); im = x.imag.astype(f32)
    fc = f32(f32(audio_bw) * f32(0.9))
    if abs_approx is None:
        p = fma32(re, re, (im * im).astype(f32))
        return lp_dc_run(fs, fc, 2.0, p, map_sqrt=True)
    k1, k2 = f32(abs_approx[0]), f32(abs_approx[1])
    e = fma32(k1, np.abs(re), (k2 * np.abs(im)).astype(f32))
    return lp_dc_run(fs, fc, 2.0, e)


def ssb_demod(fs, bfo_hz, audio_bw, x):
    x = np.asarray(x, np.complex64)
    p = rotator_phasors(bfo_hz, fs, x.size)
    y = mix_usb(x, p)
    return lp_dc_run(fs, f32(f32(audio_bw) * f32(0.9)), 2.0, y)


def cw_demod(fs, env_bw, x, gain=1.0):
    x = np.asarray(x, np.complex64)
    re = x.real.astype(f32); im = x.imag.astype(f32)
    mag = np.sqrt(((re * re).astype(f32) + (im * im).astype(f32)).astype(f32)).astype(f32)
    a = cw_alpha(fs, env_bw)
    oma = f32(f32(1.0) - a)
    y = f32(0.0)
    out = np.zeros(x.size, f32)
    for i in range(x.size):
        y = f32(f32(a * y) + f32(oma * mag[i]))
        out[i] = f32(y * f32(gain))
    return out


def agc_rms(fs, attack_ms, release_ms, target_rms, x, env0=0.0, iq=False):
    """AgcRms / AgcRmsIq (src/dsp/agc.rs:20-74, :93-149), numpy-f32 restatement; returns (out, env)."""
    fs = f32(fs)
    a_att = expf(f32(-1.0) / f32(fs * f32(max(f32(attack_ms), f32(1e-3)) / f32(1000.0))))
    a_rel = expf(f32(-1.0) / f32(fs * f32(max(f32(release_ms), f32(1e-3)) / f32(1000.0))))
    target = max(f32(target_rms), f32(1e-6))
    if iq:
        x = np.asarray(x, np.complex64)
        re = x.real.astype(f32); im = x.imag.astype(f32)
        x2 = ((re * re).astype(f32) + (im * im).astype(f32)).astype(f32)
        out = np.zeros(x.size, np.complex64)
    else:
        x = np.asarray(x, f32)
        x2 = (x * x).astype(f32)
        out = np.zeros(x.size, f32)
    env = f32(env0)
    if x.size and env == f32(0.0):
        env = max(x2[0], f32(1e-12))
    one = f32(1.0)
    for i in range(x.size):
        a = a_att if x2[i] > env else a_rel
        env = f32(f32(a * env) + f32(f32(one - a) * x2[i]))
        rms = max(f32(np.sqrt(env)), f32(1e-6))
        g = f32(target / rms)
        g = min(max(g, f32(0.05)), f32(20.0))
        if iq:
            out[i] = np.complex64(complex(f32(g * re[i]), f32(g * im[i])))
        else:
            out[i] = f32(g * x[i])
    return out, env


# ---- soft-symbol gain and hard-decision slicers (src/demodulate/{bpsk,qpsk,qam}.rs) ----------------------------
def symbol_gain(x, g):
    """BpskDemod / QpskDemod / QamDemod::process (bpsk.rs:31-50): out = (g*re, g*im)."""
    x = np.asarray(x, np.complex64)
    g = f32(g)
    out = np.empty(x.size, np.complex64)
    out.real = (g * x.real.astype(f32)).astype(f32)
    out.imag = (g * x.imag.astype(f32)).astype(f32)
    return out


def bpsk_decide(x, out_cap=None):
    """BpskDecider::process (bpsk.rs:67-88): one bit per symbol, re < 0."""
    x = np.asarray(x, np.complex64)
    n = x.size if out_cap is None else min(x.size, out_cap)
    return (x.real[:n] < 0).astype(np.uint8)


def qpsk_decide(x, out_cap=None):
    """QpskDecider::process (qpsk.rs:68-98): bits (re < 0, im < 0) per symbol; n_syms = min(len(in), len(out) / 2)."""
    x = np.asarray(x, np.complex64)
    n = x.size if out_cap is None else min(x.size, out_cap // 2)
    out = np.empty(2 * n, np.uint8)
    out[0::2] = x.real[:n] < 0
    out[1::2] = x.imag[:n] < 0
    return out


def qam_axis_scale(bits):
    """modulate/qam.rs:27-31: f64 average energy, f64 sqrt and reciprocal, cast to f32."""
    m = 1 << (bits // 2)
    return f32(1.0 / np.sqrt(2.0 * float(m * m - 1) / 3.0))


def qam_thresholds(bits):
    """qam.rs:20-31: midpoints -(M-2) + 2j, scaled, j < M-1."""
    m = 1 << (bits // 2)
    scale = qam_axis_scale(bits)
    return np.array([f32(f32(f32(2 * j) - f32(m - 2)) * scale) for j in range(m - 1)], f32)


def qam_decide(x, bits, out_cap=None):
    """QamDecider<BITS>::process (qam.rs:122-178): per axis, natural index = thresholds below v, Gray, MSB first."""
    assert bits in (4, 6, 8)
    x = np.asarray(x, np.complex64)
    n = x.size if out_cap is None else min(x.size, out_cap // bits)
    k = bits // 2
    th = qam_thresholds(bits)
    out = np.empty(n * bits, np.uint8)
    for axis, v in ((0, x.real[:n].astype(f32)), (1, x.imag[:n].astype(f32))):
        nat = (v[:, None] > th[None, :]).sum(axis=1)
        gray = nat ^ (nat >> 1)
        for b in range(k):
            out[axis * k + b::bits] = (gray >> (k - 1 - b)) & 1
    return out
