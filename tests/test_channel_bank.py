"""Channel bank (BASELINE config 5) and its sharding across processes.

CPU half: the sharding arithmetic and the host-side gather over `gloo` with world_size 2 -- each rank
computes its channel range (with the CPU oracle standing in for the GPU bank, which this container
cannot run) and rank 0 must assemble exactly the single-process result.
GPU half: `ChannelBank` through the C ABI against the oracle's block-by-block composition per channel."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle
import orion_b200 as ob
from signals import assert_parity, c5_oracle_channel, c5_specs, c5_wideband

CFG = dict(fs=1.024e6, m=16, n_channels=12, spacing_hz=16e3, cutoff_hz=3.5e3, trans_hz=16e3)


def test_shard_ranges_partition_the_channels():
    for C in (1, 7, 12, 1024):
        for W in (1, 2, 3, 4, 8):
            got = [c for r in range(W) for c in ob.shard_range(C, r, W)]
            assert got == list(range(C))
            sizes = [len(ob.shard_range(C, r, W)) for r in range(W)]
            assert max(sizes) - min(sizes) <= 1
    assert list(ob.shard_range(1024, 3, 8)) == list(range(384, 512))
    with pytest.raises(ValueError):
        ob.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, n, path):
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here), os.path.join(os.path.dirname(here), "orion-sdr_b200", "python")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        x = c5_wideband(n, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])      # every rank regenerates it from the seed
        mine = ob.shard_range(CFG["n_channels"], rank, world)
        local = np.stack([c5_oracle_channel(oracle, x, c, **CFG) for c in mine])
        full = ob.gather_channels(local, CFG["n_channels"], rank, world)
        if rank == 0:
            np.save(path, full)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    n = 8192
    path = str(tmp_path / "gathered.npy")
    mp.spawn(_rank_main, args=(2, _free_port(), n, path), nprocs=2, join=True)
    got = np.load(path)
    x = c5_wideband(n, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])
    want = np.stack([c5_oracle_channel(oracle, x, c, **CFG) for c in range(CFG["n_channels"])])
    assert got.shape == want.shape == (CFG["n_channels"], n // CFG["m"])
    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_bank_matches_block_by_block_composition():
    n = 65_536
    x = c5_wideband(n, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])
    specs = c5_specs(ob, **CFG)
    bank = ob.ChannelBank(specs)
    assert len(bank) == CFG["n_channels"]
    out = bank.process(x)
    assert out.shape == (CFG["n_channels"], n // CFG["m"]) and out.dtype == np.float32
    for c in range(CFG["n_channels"]):
        assert_parity(out[c], c5_oracle_channel(oracle, x, c, **CFG), what=f"channel {c}")
    assert bank.launch_count == 3                # one shared front-end kernel + one batched launch per demodulator kind
    # streaming: a second call continues every channel's state
    x2 = c5_wideband(2 * n, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])
    bank.reset()
    a = np.concatenate([bank.process(x2[:n]), bank.process(x2[n:])], axis=1)
    b = ob.ChannelBank(specs).process(x2)
    for c in range(CFG["n_channels"]):
        assert_parity(a[c], b[c], what=f"channel {c} two calls vs one")


@pytest.mark.gpu
def test_bank_with_nco_mixers_matches_block_by_block_composition():
    """Channels cut out with the unfused mixer (Nco::mix, nco.rs:63-66) instead of Rotator::rotate_block: the shared front
    end's MIX_NCO instance (every product and sum rounded on its own) against NcoMixer -> FirDecimator -> FM per channel."""
    n = 32_768
    x = c5_wideband(n, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])
    specs = c5_specs(ob, **CFG)
    for sp in specs:
        sp["mix"] = ob.MIX_NCO
        sp["demod"], sp["p0"] = ob.DEMOD_FM, 2.5e3
    bank = ob.ChannelBank(specs)
    out = bank.process(x)
    taps = oracle.fir_lowpass_taps(CFG["fs"], CFG["cutoff_hz"], CFG["trans_hz"])
    for c, sp in enumerate(specs):
        y = oracle.Nco(sp["mix_freq_hz"], CFG["fs"]).mix(x)
        y = oracle.FirDecimator(taps=taps, m=CFG["m"]).run(y)
        ref = oracle.FmQuadratureDemod(CFG["fs"] / CFG["m"], 2.5e3, sp["audio_bw_hz"]).run(y)
        assert_parity(out[c], ref, what=f"NCO-mixed channel {c}")


@pytest.mark.gpu
def test_bank_general_path_agrees_with_the_batched_path(monkeypatch):
    """Banks the batched kernels do not cover (here: forced) run one block per channel; both paths meet the oracle."""
    n = 40_000
    x = c5_wideband(n, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])
    specs = c5_specs(ob, **CFG)
    fast = ob.ChannelBank(specs).process(x)
    monkeypatch.setenv("ORION_B200_BANK_GENERAL", "1")
    general_bank = ob.ChannelBank(specs)
    general = general_bank.process(x)
    assert general_bank.launch_count == CFG["n_channels"]
    for c in range(CFG["n_channels"]):
        ref = c5_oracle_channel(oracle, x, c, **CFG)
        assert_parity(fast[c], ref, what=f"batched, channel {c}")
        assert_parity(general[c], ref, what=f"general, channel {c}")


@pytest.mark.gpu
def test_bank_short_output_and_ragged_lengths():
    """decim.rs:66-75 per channel: all input consumed, min(ceil(n/m), out_stride) outputs written; odd call lengths."""
    import ctypes as C
    specs = c5_specs(ob, **CFG)
    bank = ob.ChannelBank(specs)
    x = c5_wideband(30_011, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])
    a = np.concatenate([bank.process(x[:10_007]), bank.process(x[10_007:10_008]), bank.process(x[10_008:])], axis=1)
    for c in (0, 5, 11):
        xs = [x[:10_007], x[10_007:10_008], x[10_008:]]
        rot = oracle.Rotator(-(c - 6) * CFG["spacing_hz"], CFG["fs"])
        dec = oracle.FirDecimator(CFG["fs"], CFG["m"], CFG["cutoff_hz"], CFG["trans_hz"])
        dem = oracle.FmQuadratureDemod(CFG["fs"] / CFG["m"], 2.5e3, 3e3) if c % 2 == 0 else oracle.AmEnvelopeDemod(CFG["fs"] / CFG["m"], 3e3)
        ref = np.concatenate([dem.run(dec.run(rot.rotate_block(seg))) for seg in xs])
        assert_parity(a[c], ref, what=f"ragged calls, channel {c}")


@pytest.mark.gpu
def test_sharded_banks_cover_the_full_bank():
    n = 32_768
    x = c5_wideband(n, CFG["fs"], CFG["n_channels"], CFG["spacing_hz"])
    specs = c5_specs(ob, **CFG)
    full = ob.ChannelBank(specs).process(x)
    parts = [ob.ChannelBank(specs, channels=ob.shard_range(CFG["n_channels"], r, 3)).process(x) for r in range(3)]
    np.testing.assert_array_equal(np.concatenate(parts, axis=0), full)        # same kernels, same inputs: bit-equal


@pytest.mark.gpu
def test_bank_runs_on_a_caller_stream():
    import torch
    import orion_b200 as ob
    from signals import c5_specs
    cfg = dict(fs=8.192e6, m=128, n_channels=64, spacing_hz=8e3, cutoff_hz=3.5e3, trans_hz=16e3)
    n = 262_144
    g = torch.Generator(device="cuda").manual_seed(3)
    x = 0.1 * torch.complex(torch.randn(n, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g))
    a, b = ob.ChannelBank(c5_specs(ob, **cfg)), ob.ChannelBank(c5_specs(ob, **cfg))
    n_out = n // cfg["m"]
    ya = torch.zeros((64, n_out), dtype=torch.float32, device="cuda")
    yb = torch.zeros_like(ya)
    torch.cuda.synchronize()
    a.process_dev(x.data_ptr(), n, ya.data_ptr(), n_out)
    a.synchronize()
    st = torch.cuda.Stream()
    b.set_stream(st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    b.process_dev(x.data_ptr(), n, yb.data_ptr(), n_out)
    e1.record(st)
    e1.synchronize()                                           # stream order alone: the outputs are complete
    assert e0.elapsed_time(e1) > 0.0
    assert torch.equal(ya, yb)
    b.set_stream(0)
    b.synchronize()
