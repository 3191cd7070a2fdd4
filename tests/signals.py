"""Synthetic IQ generators shared by the parity tests, smoke() and bench.py.
Signals are built in f64 and cast to complex64; every generator is seeded."""
import numpy as np


def awgn(n, sigma, seed):
    r = np.random.default_rng(seed)
    return sigma * (r.standard_normal(n) + 1j * r.standard_normal(n))


def fm_iq(n, fs, f_c=100e3, dev=25e3, tones=((1e3, 0.5), (3.7e3, 0.25)), amp=0.5, sigma=1e-3, seed=0x0510, start=0):
    """C1: carrier offset f_c, two-tone message, peak deviation `dev`, plus AWGN."""
    t = (start + np.arange(n)) / fs
    msg_int = np.zeros(n)
    for f, a in tones:                       # integral of a*cos(2 pi f t)
        msg_int += a * np.sin(2 * np.pi * f * t) / (2 * np.pi * f)
    ph = 2 * np.pi * f_c * t + 2 * np.pi * dev * msg_int
    return (amp * np.exp(1j * ph) + awgn(n, sigma, seed)).astype(np.complex64)


def am_iq(n, fs, carrier=0.8, m=0.5, tones=(400.0, 1e3), sigma=1e-3, seed=0x0512, f_off=0.0, start=0):
    t = (start + np.arange(n)) / fs
    msg = sum(np.cos(2 * np.pi * f * t) for f in tones) / len(tones)
    env = carrier * (1.0 + m * msg)
    return (env * np.exp(2j * np.pi * f_off * t) + awgn(n, sigma, seed)).astype(np.complex64)


def ssb_iq(n, fs, f_bfo=250e3, tones=(300.0, 700.0, 1200.0, 1900.0, 2500.0), a=0.15, sigma=1e-3, seed=0x0511, start=0):
    t = (start + np.arange(n)) / fs
    x = sum(a * np.exp(2j * np.pi * (f_bfo + f) * t) for f in tones)
    return (x + awgn(n, sigma, seed)).astype(np.complex64)


def pm_iq(n, fs, k=0.8, tone=1e3, amp=0.5, sigma=1e-3, seed=0x0515):
    t = np.arange(n) / fs
    return (amp * np.exp(1j * k * np.sin(2 * np.pi * tone * t)) + awgn(n, sigma, seed)).astype(np.complex64)


def cw_iq(n, fs, tone=700.0, wpm_period=0.12, amp=0.6, sigma=1e-3, seed=0x0516):
    t = np.arange(n) / fs
    key = ((t / wpm_period) % 1.0) < 0.5
    return (amp * key * np.exp(2j * np.pi * tone * t) + awgn(n, sigma, seed)).astype(np.complex64)


def blockwise(gen, n, blk=2_400_000, seed=0, **kw):
    """A long stream built block by block (bounded host memory): `gen(n, ..., seed=, start=)` per block, the noise
    seed advancing with the block index so that the stream is reproducible whatever the block size of a caller."""
    out = np.empty(n, np.complex64)
    for i, s in enumerate(range(0, n, blk)):
        e = min(n, s + blk)
        out[s:e] = gen(e - s, seed=seed + i, start=s, **kw)
    return out


def wideband_noise_tones(n, fs, tones_hz=(1.1e5, -2.3e5, 3.7e5), sigma=0.25, amp=0.2, seed=0x0513, start=0):
    """C4: complex white noise plus a few in-band tones."""
    t = (start + np.arange(n)) / fs
    x = awgn(n, sigma, seed)
    for f in tones_hz:
        x = x + amp * np.exp(2j * np.pi * f * t)
    return x.astype(np.complex64)


def noise_c64(n, scale=0.5, seed=1):
    r = np.random.default_rng(seed)
    return (scale * (r.standard_normal(n) + 1j * r.standard_normal(n))).astype(np.complex64)


def noise_f32(n, scale=0.5, seed=2):
    return (scale * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


def parity(got, ref):
    """(max abs error / full scale, SNR in dB) with full scale := max|ref|."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0, np.inf
    fsv = float(np.max(np.abs(ref)))
    d = got.astype(np.complex128) - ref.astype(np.complex128)
    e = float(np.max(np.abs(d)))
    p_ref = float(np.sum(np.abs(ref.astype(np.complex128)) ** 2))
    p_err = float(np.sum(np.abs(d) ** 2))
    snr = np.inf if p_err == 0.0 else 10.0 * np.log10(max(p_ref, 1e-300) / p_err)
    return (e / fsv if fsv > 0 else e), snr


TOL = 1e-4       # max abs error, fraction of full scale (BASELINE.json north_star)
SNR_DB = 90.0    # demodulated-output SNR against the reference


def assert_parity(got, ref, tol=TOL, snr_db=SNR_DB, what=""):
    e, snr = parity(got, ref)
    assert e <= tol, f"{what}: max abs error {e:.3e} of full scale > {tol:.1e} (snr {snr:.1f} dB)"
    assert snr >= snr_db, f"{what}: SNR {snr:.1f} dB < {snr_db} dB (max err {e:.3e})"
    return e, snr


def bit_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    fa = a.view(np.float32)
    fb = b.view(np.float32)
    nan = np.isnan(fa)
    return bool(np.array_equal(nan, np.isnan(fb)) and np.array_equal(fa.view(np.uint32)[~nan], fb.view(np.uint32)[~nan]))


# ---- C5: many narrowband channels in one wideband stream ------------------------
def c5_channel_freqs(n_channels, spacing_hz):
    return [(c - n_channels // 2) * spacing_hz for c in range(n_channels)]


def c5_wideband(n, fs, n_channels, spacing_hz, amp=0.02, dev_hz=2.5e3, sigma=1e-4, seed=0x0514, only=None, start=0):
    """Channel c sits at (c - C/2) * spacing and carries FM (c even) or AM (c odd) with a per-channel tone.
    `only`: the channels that carry a signal (default all)."""
    t = (start + np.arange(n)) / fs
    x = awgn(n, sigma, seed)
    for c, fc in enumerate(c5_channel_freqs(n_channels, spacing_hz)):
        if only is not None and c not in only:
            continue
        tone = 300.0 + (c % 17) * 100.0
        if c % 2 == 0:
            ph = 2 * np.pi * fc * t + (dev_hz / tone) * np.sin(2 * np.pi * tone * t)
            x = x + amp * np.exp(1j * ph)
        else:
            x = x + amp * (1.0 + 0.5 * np.cos(2 * np.pi * tone * t)) * np.exp(2j * np.pi * fc * t)
    return x.astype(np.complex64)


def c5_specs(ob, fs, m, n_channels, spacing_hz, cutoff_hz, trans_hz, dev_hz=2.5e3, audio_bw_hz=3e3):
    """One chain spec per channel: Rotator(-f_c) -> FirDecimator(fs, m, cutoff, trans) -> FM (even) / AM (odd)."""
    taps = ob.fir_lowpass_design(fs, cutoff_hz, trans_hz)
    specs = []
    for c, fc in enumerate(c5_channel_freqs(n_channels, spacing_hz)):
        sp = dict(mix=ob.MIX_ROTATE, mix_freq_hz=-fc, mix_fs=fs, fir=ob.FIR_DECIM, taps=taps, decim=m, fs_demod=fs / m,
                  audio_bw_hz=audio_bw_hz)
        if c % 2 == 0:
            sp.update(demod=ob.DEMOD_FM, p0=dev_hz)
        else:
            sp.update(demod=ob.DEMOD_AM)
        specs.append(sp)
    return specs


def c5_oracle_channel(oracle, x, c, fs, m, n_channels, spacing_hz, cutoff_hz, trans_hz, dev_hz=2.5e3, audio_bw_hz=3e3, fast=False):
    """The reference composition for channel c: three blocks run back to back through intermediate vectors.
    fast: the FirDecimator's kept outputs are evaluated directly (oracle.fir_decim_kept, bit-identical)."""
    fc = c5_channel_freqs(n_channels, spacing_hz)[c]
    y = oracle.Rotator(-fc, fs).rotate_block(x)
    if fast:
        y = oracle.fir_decim_kept(oracle.fir_lowpass_taps(fs, cutoff_hz, trans_hz), m, y)
    else:
        y = oracle.FirDecimator(fs, m, cutoff_hz, trans_hz).run(y)
    dem = oracle.FmQuadratureDemod(fs / m, dev_hz, audio_bw_hz) if c % 2 == 0 else oracle.AmEnvelopeDemod(fs / m, audio_bw_hz)
    return dem.run(y)
