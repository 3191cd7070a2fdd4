"""GPU parity at the BASELINE configurations' OWN sizes and shapes.

Every test drives the C ABI with host buffers (orion_b200_block_process / orion_b200_bank_process) over the full
stream and compares ALL outputs with the reference composition run block by block on the CPU oracle.  The oracle's
input-rate FIRs are evaluated at the kept outputs only (oracle.fir_decim_kept / fir_iq_kept: bit-identical to the
loop-for-loop blocks, pinned by tests/test_oracle_crosscheck.py), which is what makes these sizes affordable.
Tolerance (BASELINE.json north_star): max abs error <= 1e-4 of full scale, SNR >= 90 dB, counts bit-exact.
"""
import numpy as np
import pytest

import oracle as O
import orion_b200 as ob
from signals import (am_iq, assert_parity, bit_equal, blockwise, c5_channel_freqs, c5_oracle_channel, c5_specs,
                     c5_wideband, fm_iq, ssb_iq, wideband_noise_tones)

pytestmark = pytest.mark.gpu


def test_c1_full_24m_samples_vs_oracle():
    """C1: FirDecimator(2.4e6, 8, 100e3, 38.4e3) -> FmQuadratureDemod(300e3, 25e3, 15e3).with_translate(100e3), 24 M samples."""
    fs, m, n = 2.4e6, 8, 24_000_000
    x = blockwise(fm_iq, n, fs=fs, seed=0x0510)
    taps = ob.fir_lowpass_design(fs, 100e3, 38400.0)
    assert taps.size == 63
    chain = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_FM, fs_demod=fs / m, p0=25e3,
                     audio_bw_hz=15e3, translate_hz=100e3)
    out = np.zeros(n // m, np.float32)
    wr = chain.process(x, out)
    assert wr == (n, n // m)
    ref = O.FmQuadratureDemod(fs / m, 25e3, 15e3).with_translate(100e3).run(O.fir_decim_kept(O.fir_lowpass_taps(fs, 100e3, 38400.0), m, x))
    assert_parity(out, ref, what="C1 at 24 M samples")
    # the tail alone (the last second of the stream): nothing accumulates over 3 M outputs
    assert_parity(out[-300_000:], ref[-300_000:], what="C1 last second")


def test_c2_full_12m_samples_ssb_vs_oracle():
    """C2: Rotator(-250e3, 1.2e6).rotate_block -> FirLowpassIq::design(201, 0.01, 60) -> keep every 25th ->
    SsbProductDemod(48e3, 0, 2800), 12 M samples.  The product detector sees the mixer's ABSOLUTE phase, so this is the
    configuration that needs the exact-replay oscillator (the closed form is 2.8e-3 rad off at 12 M samples)."""
    fs, m, n = 1.2e6, 25, 12_000_000
    x = blockwise(ssb_iq, n, fs=fs, seed=0x0511)
    taps = ob.kaiser_lowpass_taps(201, 0.01, 60.0)
    chain = ob.Chain(mix=ob.MIX_ROTATE, mix_freq_hz=-250e3, mix_fs=fs, fir=ob.FIR_IQ, taps=taps, decim=m,
                     demod=ob.DEMOD_SSB, fs_demod=fs / m, p0=0.0, audio_bw_hz=2800.0)
    out = np.zeros(n // m, np.float32)
    wr = chain.process(x, out)
    assert wr == (n, n // m)
    y = O.Rotator(-250e3, fs).rotate_block(x)
    y = O.fir_iq_kept(O.kaiser_lowpass_taps(201, 0.01, 60.0), m, y)
    ref = O.SsbProductDemod(fs / m, 0.0, 2800.0).run(y)
    assert_parity(out, ref, what="C2 at 12 M samples")
    assert_parity(out[-48_000:], ref[-48_000:], what="C2 last second")
    assert chain.exact_host_ms > 0.0                     # the oscillator recurrence was walked, and is accounted for


def test_c3_full_38m_samples_am_four_sections_vs_oracle():
    """C3: FirDecimator(384e3, 8, 10e3, 6144) -> AmEnvelopeDemod(48e3, 5e3) -> LpCascade(48e3, 3e3): two biquads + sqrt +
    2 Hz DC blocker + two more biquads over 4.8 M outputs (18 750 tiles: the slow-pole look-back chain at its real depth)."""
    fs, m, n = 384e3, 8, 38_400_000
    x = blockwise(am_iq, n, fs=fs, seed=0x0512)
    taps = ob.fir_lowpass_design(fs, 10e3, 6144.0)
    assert taps.size == 63
    sos = np.tile(ob.lp_biquad_design(fs / m, 3e3), (2, 1))
    for abs_approx in (False, True):
        chain = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=m, demod=ob.DEMOD_AM_ABS if abs_approx else ob.DEMOD_AM,
                         fs_demod=fs / m, p0=0.9482, p1=0.3920, audio_bw_hz=5e3, post_sos=sos)
        out = np.zeros(n // m, np.float32)
        wr = chain.process(x, out)
        assert wr == (n, n // m)
        y = O.fir_decim_kept(O.fir_lowpass_taps(fs, 10e3, 6144.0), m, x)
        am = O.AmEnvelopeDemod(fs / m, 5e3, abs_approx=abs_approx)
        ref = O.LpCascade(fs / m, 3e3).run(am.run(y))
        assert not np.isnan(ref).any()
        assert_parity(out, ref, what=f"C3 at 38.4 M samples (abs_approx={abs_approx})")
        assert_parity(out[-48_000:], ref[-48_000:], what="C3 last second")
        del chain


@pytest.mark.parametrize("design", ["reference_designer", "kaiser_fir_iq"])
def test_c4_1023_taps_decimate_32_vs_oracle(design):
    """C4 at 6.4 M samples, both tap designs of SURVEY 8d: (i) FirDecimator::new(100e6, 32, 450e3, 97.8e3) -- the
    reference designer's quirky 1023 taps with the FirLowpass pairing; (ii) kaiser_lowpass_taps(1023, 1/64, 80) through
    FirLowpassIq + keep every 32nd."""
    fs, m, n = 100e6, 32, 6_400_000
    x = blockwise(wideband_noise_tones, n, fs=fs, seed=0x0513)
    if design == "reference_designer":
        taps = ob.fir_lowpass_design(fs, 450e3, 97800.0)
        blk = ob.FirDecimator.from_taps(taps, m)
        ref = O.fir_decim_kept(O.fir_lowpass_taps(fs, 450e3, 97800.0), m, x)
    else:
        taps = ob.kaiser_lowpass_taps(1023, 1.0 / 64.0, 80.0)
        blk = ob.Chain(fir=ob.FIR_IQ, taps=taps, decim=m)
        ref = O.fir_iq_kept(O.kaiser_lowpass_taps(1023, 1.0 / 64.0, 80.0), m, x)
    assert taps.size == 1023
    out = np.zeros(n // m, np.complex64)
    wr = blk.process(x, out)
    assert wr == (n, n // m)
    assert_parity(out, ref, what=f"C4 ({design}) at 6.4 M samples")


def test_c5_true_plan_64_channels_one_million_samples_all_compared():
    """C5 with the real plan: fs 8.192 MS/s, FirDecimator(513 taps, /128), 64 of the 1024 channel slots (edges, centre,
    FM and AM alike), 1 M wideband samples, EVERY channel compared with the block-by-block oracle composition."""
    cfg = dict(fs=8.192e6, m=128, n_channels=1024, spacing_hz=8e3, cutoff_hz=3.5e3, trans_hz=16e3)
    n = 1_048_576
    ids = sorted(set(list(range(0, 16)) + list(range(504, 520)) + list(range(1008, 1024)) + list(range(100, 1000, 57))))[:64]
    assert len(ids) == 64
    x = blockwise(c5_wideband, n, blk=262_144, fs=cfg["fs"], n_channels=1024, spacing_hz=cfg["spacing_hz"], only=set(ids), seed=0x0514)
    specs = c5_specs(ob, **cfg)
    assert specs[0]["taps"].size == 513
    bank = ob.ChannelBank(specs, channels=ids)
    y = bank.process(x)
    assert y.shape == (64, n // 128)
    worst = (0.0, 1e9)
    for i, c in enumerate(ids):
        ref = c5_oracle_channel(O, x, c, fast=True, **cfg)
        e, snr = assert_parity(y[i], ref, what=f"C5 channel {c}")
        worst = (max(worst[0], e), min(worst[1], snr))
    print(f"C5 true plan, 64 channels: worst max-err {worst[0]:.2e} of full scale, worst SNR {worst[1]:.1f} dB")
    # a bank over a sub-range is bit-identical to the same channels of the larger bank (what sharding relies on)
    sub = ob.ChannelBank(specs, channels=ids[16:32])
    assert bit_equal(sub.process(x), y[16:32])


@pytest.mark.parametrize("f,fs,n", [(100e3, 2.4e6, 24_000_000), (1.5e3, 48e3, 4_800_000), (-250e3, 1.2e6, 12_000_000)])
def test_absolute_phase_blocks_are_bit_exact_on_long_streams(f, fs, n):
    """Rotator::rotate_block, mix_usb_block and mix_with_nco over BASELINE-length streams: the exact-replay oscillator
    reproduces the reference's f32 phasor recurrence (renormalised every 1024 steps), so the outputs are bit-identical,
    not merely within tolerance -- in one call and in ragged chunks alike."""
    r = np.random.default_rng(7)
    x = (0.5 * (r.standard_normal(n, np.float32) + 1j * r.standard_normal(n, np.float32))).astype(np.complex64)
    assert bit_equal(ob.Rotator(f, fs).run(x), O.Rotator(f, fs).rotate_block(x))
    assert bit_equal(ob.RotatorUsb(f, fs).run(x), O.Rotator(f, fs).mix_usb_block(x))
    assert bit_equal(ob.NcoMixer(f, fs).run(x), O.Nco(f, fs).mix(x))
    g = ob.Rotator(f, fs)
    cuts = [0, 1, 1000, 1023, 1024, 1025, 12_345, n // 3 + 5, n]
    got = np.concatenate([g.run(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert bit_equal(got, O.Rotator(f, fs).rotate_block(x))


def test_ssb_demod_rate1_long_stream():
    """SsbProductDemod alone over 4.8 M samples at 48 kS/s with a non-zero BFO (1.5 kHz): absolute BFO phase + 2 Hz DC
    blocker (the slow pole) over 600 tiles."""
    fs, n = 48e3, 4_800_000
    x = blockwise(ssb_iq, n, fs=fs, f_bfo=1.5e3, seed=0x0517)
    out = ob.SsbProductDemod(fs, 1.5e3, 2800.0).run(x)
    ref = O.SsbProductDemod(fs, 1.5e3, 2800.0).run(x)
    assert_parity(out, ref, what="SsbProductDemod at 4.8 M samples")
