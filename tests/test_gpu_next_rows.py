"""GPU parity for the remaining "next" rows of SURVEY.md section 8(f): the FM / CW / SSB modulators (row 1), the
soft-symbol gain blocks, hard-decision slicers and the CFO de-rotation call site (row 4).

Modulators are compared with the C oracle (loop-for-loop restatement of src/modulate/{fm,cw,ssb}.rs); the slicers, which
are byte work, with the numpy oracle and must be bit-exact."""
import numpy as np
import pytest

import oracle
import orion_b200 as ob
from oracle import np_oracle as npo
from signals import assert_parity, noise_c64, parity

pytestmark = pytest.mark.gpu


def _audio(n, fs, tones=((400.0, 0.5), (1100.0, 0.3)), seed=3):
    t = np.arange(n) / fs
    x = sum(a * np.sin(2 * np.pi * f * t) for f, a in tones)
    return (x + 1e-3 * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


def _stream(gpu, ref, x, chunks, out_dtype):
    og, orf, pos = [], [], 0
    for c in chunks:
        a, b = np.zeros(c, out_dtype), np.zeros(c, out_dtype)
        wg, wr = gpu.process(x[pos:pos + c], a), ref.process(x[pos:pos + c], b)
        assert tuple(wg) == tuple(wr)
        og.append(a); orf.append(b); pos += c
    return np.concatenate(og), np.concatenate(orf)


@pytest.mark.parametrize("fs,dev,rf,n", [(48e3, 2.5e3, 0.0, 100_000), (2.4e6, 25e3, 100e3, 262_144)])
def test_fm_modulator(fs, dev, rf, n):
    """The reference's running phasor is a prefix sum of the phase; its own f32 rounding walk bounds how long the two can
    agree (~1e6 samples after a reset, include/orion_b200.h), hence the lengths here."""
    x = _audio(n, fs)
    g, ref = ob.FmPhaseAccumMod(fs, dev, rf), oracle.FmPhaseAccumMod(fs, dev, rf)
    g.set_gain(0.7); ref.set_gain(0.7)
    og, orf = _stream(g, ref, x, [n // 3, 5, n - n // 3 - 5], np.complex64)
    assert_parity(og, orf, what=f"fm mod fs={fs}")


def test_fm_modulator_roundtrip_through_the_gpu_demodulator():          # roundtrip/fm.rs:11-27 with both ends on the GPU
    fs, n = 48e3, 48_000
    t = np.arange(n) / fs
    x = (0.5 * np.sin(2 * np.pi * 1e3 * t)).astype(np.float32)
    iq = ob.FmPhaseAccumMod(fs, 2.5e3, 0.0).run(x)
    y = ob.FmQuadratureDemod(fs, 2.5e3, 5e3).run(iq)
    yr = oracle.FmQuadratureDemod(fs, 2.5e3, 5e3).run(oracle.FmPhaseAccumMod(fs, 2.5e3, 0.0).run(x))
    assert_parity(y, yr, what="fm mod -> demod")


@pytest.mark.parametrize("fs,tone,rise,fall,n", [(48e3, 700.0, 5.0, 5.0, 200_000), (48e3, 0.0, 2.0, 8.0, 60_000), (2.4e6, 1e5, 0.5, 1.0, 1_500_000)])
def test_cw_modulator(fs, tone, rise, fall, n):
    key = (((np.arange(n) / fs) / 0.06) % 1.0 < 0.5).astype(np.float32) * 1.2 - 0.1      # overshoots [0, 1]: the clamp matters
    g, ref = ob.CwKeyedMod(fs, tone, rise, fall), oracle.CwKeyedMod(fs, tone, rise, fall)
    g.set_gain(0.8); ref.set_gain(0.8)
    og, orf = _stream(g, ref, key, [n // 2, 3, n - n // 2 - 3], np.complex64)
    assert_parity(og, orf, what=f"cw mod fs={fs}")


@pytest.mark.parametrize("usb", [True, False])
def test_ssb_modulator(usb):
    fs, n = 48e3, 150_000
    x = _audio(n, fs, tones=((700.0, 0.4), (1900.0, 0.3)))
    g, ref = ob.SsbPhasingMod(fs, 2.8e3, 1.5e3, 6e3, usb), oracle.SsbPhasingMod(fs, 2.8e3, 1.5e3, 6e3, usb)
    og, orf = _stream(g, ref, x, [n // 2, 7, n - n // 2 - 7], np.complex64)
    assert_parity(og, orf, what=f"ssb mod usb={usb}")


def test_symbol_gain_blocks_are_bit_exact():
    x = noise_c64(100_003, seed=11)
    for cls in (ob.BpskDemod, ob.QpskDemod, ob.QamDemod):
        g = cls(0.37)
        out = np.zeros(90_000, np.complex64)                       # shorter than the input: n = min(len(in), len(out))
        wr = g.process(x, out)
        assert tuple(wr) == (90_000, 90_000)
        assert np.array_equal(out.view(np.uint32), npo.symbol_gain(x[:90_000], 0.37).view(np.uint32))
        g.set_gain(2.0)
        assert np.array_equal(g.run(x).view(np.uint32), npo.symbol_gain(x, 2.0).view(np.uint32))


@pytest.mark.parametrize("cls,bits", [("BpskDecider", 1), ("QpskDecider", 2), ("Qam16Decider", 4), ("Qam64Decider", 6), ("Qam256Decider", 8)])
def test_deciders_are_bit_exact(cls, bits):
    r = np.random.default_rng(bits)
    n = 200_001
    x = (r.standard_normal(n) + 1j * r.standard_normal(n)).astype(np.complex64) * 0.7
    if bits >= 4:                                                  # symbols exactly on decision thresholds and on zero
        th = npo.qam_thresholds(bits)
        x[:th.size] = th + 1j * th[::-1]
    x[100] = 0.0
    x[101] = complex(-0.0, 0.0)
    ref_fn = {1: npo.bpsk_decide, 2: npo.qpsk_decide}.get(bits, lambda v, cap=None: npo.qam_decide(v, bits, cap))
    g = getattr(ob, cls)()
    out = g.run(x)
    assert out.dtype == np.uint8 and out.size == n * bits
    assert np.array_equal(out, ref_fn(x))
    cap = 1000 * bits + (bits - 1)                                 # a ragged output slice: whole symbols only
    short = np.full(cap, 255, np.uint8)
    wr = g.process(x, short)
    assert tuple(wr) == (1000, 1000 * bits)
    assert np.array_equal(short[:1000 * bits], ref_fn(x, cap)) and np.all(short[1000 * bits:] == 255)


@pytest.mark.parametrize("cls,bits", [("BpskDecider", 1), ("QpskDecider", 2), ("Qam16Decider", 4)])
def test_deciders_invert_the_reference_mapper_known_answers(cls, bits):
    """The constellation points the reference's own unit tests write down (tests/unit/bpsk.rs:9-19, qpsk.rs:9-18,
    qam.rs:9-41) through the GPU deciders: the bit patterns of those tests come back."""
    from test_oracle_behaviour import reference_mapper_vector
    want, syms = reference_mapper_vector(bits)
    assert np.array_equal(getattr(ob, cls)().run(syms), want)


def test_decider_rejects_unsupported_constellations():            # qam.rs:13-18 check_bits
    with pytest.raises(ob.OrionB200Error):
        ob._Decider.bits = 5
        try:
            ob._Decider()
        finally:
            ob._Decider.bits = 1


def test_cfo_derotation_call_site():                               # sync/ofdm_sync.rs:527-528
    fs, cfo, n = 2.4e6, 1234.5, 300_000
    x = noise_c64(n, seed=5)
    got = ob.cfo_derotate(x, cfo, fs)
    ref = oracle.Rotator(-cfo, fs).run(x)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))    # exact-replay oscillator: bit for bit
