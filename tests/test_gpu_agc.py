"""GPU parity: AgcRms / AgcRmsIq (src/dsp/agc.rs) through the C ABI against the oracle.

The envelope tracker is a data-dependent recurrence; the GPU evaluates it in chunks with a warm-up (agc_kernels.cu).
Bar: WorkReport bit-exact, values within 1e-4 of full scale and SNR >= 90 dB; the first chunk of every call starts
from the carried state and must be bit-identical."""
import os

import numpy as np
import pytest

import oracle
import orion_b200 as ob
from signals import assert_parity

pytestmark = pytest.mark.gpu


def _levels(n, period, lo=0.03, hi=0.9):
    return np.where((np.arange(n) // period) % 2 == 0, lo, hi)


@pytest.mark.parametrize("iq", [False, True])
@pytest.mark.parametrize("fs,att,rel,n", [(48e3, 5.0, 50.0, 400_000), (48e3, 0.2, 5.0, 100_000), (2.4e6, 0.05, 0.5, 3_000_000)])
def test_agc_matches_oracle(iq, fs, att, rel, n):
    r = np.random.default_rng(0xA6C1)
    lvl = _levels(n, max(n // 7, 1))
    if iq:
        x = (lvl * (r.standard_normal(n) + 1j * r.standard_normal(n))).astype(np.complex64)
        g, ref = ob.AgcRmsIq(fs, att, rel, 0.25), oracle.AgcRmsIq(fs, att, rel, 0.25)
    else:
        x = (lvl * r.standard_normal(n)).astype(np.float32)
        g, ref = ob.AgcRms(fs, att, rel, 0.25), oracle.AgcRms(fs, att, rel, 0.25)
    chunks = [n // 3, 1, n - n // 3 - 1]                      # streaming: the envelope is carried across calls
    og, orf, pos = [], [], 0
    for c in chunks:
        a = np.zeros(c, x.dtype); b = np.zeros(c, x.dtype)
        wg = g.process(x[pos:pos + c], a); wr = ref.process(x[pos:pos + c], b)
        assert tuple(wg) == tuple(wr)
        og.append(a); orf.append(b); pos += c
    og, orf = np.concatenate(og), np.concatenate(orf)
    assert_parity(og, orf, what=f"agc iq={iq} fs={fs}")
    assert abs(g.env - ref.env) <= 1e-6 * max(ref.env, 1e-12)


def test_agc_first_chunk_is_bit_exact_and_short_output_rule():
    r = np.random.default_rng(7)
    x = (0.3 * r.standard_normal(5000)).astype(np.float32)
    g, ref = ob.AgcRms(48e3, 5.0, 50.0, 0.3), oracle.AgcRms(48e3, 5.0, 50.0, 0.3)
    a = np.zeros(3000, np.float32); b = np.zeros(3000, np.float32)          # out shorter than in: n = min(len)
    wg, wr = g.process(x, a), ref.process(x, b)
    assert tuple(wg) == tuple(wr) == (3000, 3000)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))             # 3000 < one chunk: the reference recursion itself
    assert np.float32(g.env) == np.float32(ref.env)


def test_agc_reference_unit_test_on_gpu():                      # tests/unit/agc.rs:8-32
    fs, n = 48_000.0, 8_000
    x = np.where(np.arange(n) < n // 2, 0.02, 1.0).astype(np.float32).astype(np.complex64)
    out = ob.AgcRmsIq(fs, 0.2, 5.0, 0.2).run(x)
    rms_tail = np.sqrt(np.mean(np.abs(out[n - 1000:]).astype(np.float64) ** 2))
    assert abs(rms_tail - 0.2) < 0.03, rms_tail


def test_agc_zero_input_reseeds_like_the_reference():           # agc.rs:58-61: env == 0 -> seeded from x[0] of the next call
    g, ref = ob.AgcRms(48e3, 5.0, 50.0, 0.3), oracle.AgcRms(48e3, 5.0, 50.0, 0.3)
    z = np.zeros(256, np.float32)
    x = np.full(256, 0.5, np.float32)
    for blk in (z, x, z, x):
        a, b = g.run(blk), ref.run(blk)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
