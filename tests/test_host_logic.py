"""CPU-only tests of the host logic behind the C ABI: the library loads and exports every
symbol of include/orion_b200.h, the restated design math equals the oracle bit for bit, and
the launch plan (polyphase tap table, staged-tile geometry, scan tables) is emulated in numpy
with exactly the index arithmetic the kernels use and checked against the direct definitions."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import orion_b200 as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KTHREADS = 32


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "orion_b200.h")).read()
    declared = set(re.findall(r"\b(orion_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = C.CDLL(ob.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"missing exports: {missing}"
    assert declared == set(ob.EXPORTED_SYMBOLS)
    assert ob.lib().orion_b200_abi_version() == 1


def test_no_device_is_an_error_not_a_fallback():
    if ob.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(ob.OrionB200Error) as e:
        ob.FmQuadratureDemod(48e3, 2.5e3, 5e3)
    assert e.value.status == ob.ERR_NO_DEVICE


@pytest.mark.parametrize("fs,p,t", [(2.4e6, 100e3, 38400.0), (96e3, 10e3, 2400.0), (100e6, 450e3, 97800.0),
                                    (384e3, 10e3, 6144.0), (48e3, 5.0, 0.5)])
def test_fir_design_matches_oracle(fs, p, t):
    a, b = ob.fir_lowpass_design(fs, p, t), oracle.fir_lowpass_taps(fs, p, t)
    assert a.size == b.size and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("n,fc,a", [(81, 0.1, 60.0), (201, 0.01, 60.0), (1023, 1 / 64, 80.0), (4, 0.3, 30.0), (31, 0.25, 10.0)])
def test_kaiser_design_matches_oracle(n, fc, a):
    x, y = ob.kaiser_lowpass_taps(n, fc, a), oracle.kaiser_lowpass_taps(n, fc, a)
    assert x.size == y.size and np.array_equal(x.view(np.uint32), y.view(np.uint32))
    assert ob.kaiser_num_taps(0.01, a) == oracle.kaiser_num_taps(0.01, a)
    assert np.float32(ob.kaiser_transition_norm(n, a)) == np.float32(oracle.kaiser_transition_norm(n, a))


def test_scalar_design_helpers_match_oracle():
    for fs, fc in [(48e3, 4.5e3), (300e3, 13.5e3), (64e3, 2.7e3)]:
        assert np.array_equal(ob.lp_biquad_design(fs, fc).view(np.uint32),
                              np.asarray(oracle.lp_biquad_coeffs(fs, fc), np.float32).view(np.uint32))
        assert np.float32(ob.dc_pole(fs, 2.0)) == np.float32(oracle.dc_pole(fs, 2.0))
        assert np.float32(ob.cw_alpha(fs, 50.0)) == np.float32(oracle.cw_alpha(fs, 50.0))


# --------------------------------------------------------------------------------------------------
# staged polyphase FIR: numpy emulation of fir_staged()/stage_tile() indexing
# --------------------------------------------------------------------------------------------------
def _emulate_staged(plan, M, x_virtual, H, n_out):
    """x_virtual[s + H] = sample s of the virtual stream (history first).  Returns y[0:n_out]."""
    R, U, Mb, O, P_pad, HR = (plan[k] for k in ("R", "U", "Mb", "O", "P_pad", "HR"))
    rs, rows = plan["row_samples"], plan["rows"]
    tab = plan["table"][:, 0].astype(np.float64), plan["table"][:, 1].astype(np.float64)
    NPT = R * U
    ntiles = -(-n_out // (KTHREADS * NPT))
    y = np.zeros(ntiles * KTHREADS * NPT, np.complex128)

    def sample(s):
        i = s + H
        ok = (i >= 0) & (i < x_virtual.size)
        return np.where(ok, x_virtual[np.clip(i, 0, x_virtual.size - 1)], 0)

    for t in range(ntiles):
        G0 = t * KTHREADS - HR
        row_start = rs * (G0 + np.arange(rows)) + (O - Mb + 2)                 # row_start_sample()
        smem = sample(row_start[:, None] + np.arange(rs)[None, :])             # [rows, rs]
        tid = np.arange(KTHREADS)
        acc = np.zeros((U, R, KTHREADS), np.complex128)
        for q in range(Mb // 2):
            off = Mb - 2 - 2 * q
            w = [None] * R
            for k in range(R - 1):
                w[k] = (smem[tid, (k + 1) * Mb + off], smem[tid, (k + 1) * Mb + off + 1])
            for rr in range(HR):
                for kk in range(R):
                    col = kk * Mb + off
                    w[(kk + R - 1) % R] = (smem[tid + 1 + rr, col], smem[tid + 1 + rr, col + 1])
                    c = rr * R + kk
                    for u in range(U):
                        ti = (u * (Mb // 2) + q) * P_pad + c
                        for i in range(R):
                            w0, w1 = w[(kk + i) % R]
                            acc[u, i] += tab[0][ti] * w0 + tab[1][ti] * w1
        for i in range(R):
            for u in range(U):
                j = t * KTHREADS * NPT + tid * NPT + i * U + u
                y[j] = acc[u, i]
    return y[:n_out]


def _direct(g, M, x_virtual, H, n_out):
    y = np.zeros(n_out, np.complex128)
    for j in range(n_out):
        for t in range(len(g)):
            s = M * j - t + H
            if 0 <= s < x_virtual.size:
                y[j] += float(g[t]) * x_virtual[s]
    return y


@pytest.mark.parametrize("kind,L,M", [
    (ob.FIR_DECIM, 63, 8), (ob.FIR_DECIM, 31, 4), (ob.FIR_DECIM, 41, 2), (ob.FIR_DECIM, 33, 16),
    (ob.FIR_DECIM, 65, 32), (ob.FIR_IQ, 81, 1), (ob.FIR_IQ, 51, 25), (ob.FIR_IQ, 21, 3), (ob.FIR_DECIM, 31, 5),
    (ob.FIR_IQ, 1, 1), (ob.FIR_DECIM, 47, 12), (ob.FIR_DECIM, 95, 48),
])
def test_staged_fir_plan_reproduces_direct_fir(kind, L, M, rng):
    taps = rng.standard_normal(L).astype(np.float32)
    plan = ob.debug_fir_plan(kind, taps, M)
    assert plan["front"] == 1, plan
    g = plan["g"]
    if kind == ob.FIR_DECIM:                      # fir.rs:57-66 pairing
        assert g[0] == taps[-1] and np.array_equal(g[1:], taps[:-1])
    else:
        assert np.array_equal(g, taps)
    R, U, Mb = plan["R"], plan["U"], plan["Mb"]
    assert Mb == M * U and Mb % 2 == 0 and plan["P_pad"] % R == 0 and plan["HR"] * R == plan["P_pad"]
    assert plan["row_pitch"] % 16 == 0 and (plan["row_pitch"] // 16) % 2 == 1      # conflict-free lane stride
    assert plan["rows"] == KTHREADS + plan["HR"] <= 256 and plan["row_pitch"] <= 2048
    H = plan["H"]
    assert H % 2 == 0 and H >= L
    # the first staged sample of tile 0 must lie inside the history
    assert plan["row_samples"] * (-plan["HR"]) + plan["O"] - Mb + 2 >= -H
    NPT = R * U
    n_out = KTHREADS * NPT + 37                   # one full tile + a ragged one
    n_in = M * n_out
    xv = (rng.standard_normal(H + n_in) + 1j * rng.standard_normal(H + n_in))
    want = _direct(g, M, xv, H, n_out)
    got = _emulate_staged(plan, M, xv, H, n_out)
    assert np.max(np.abs(got - want)) < 1e-9 * max(1.0, np.max(np.abs(want)))


def test_large_shapes_fall_back_to_global_front(rng):
    plan = ob.debug_fir_plan(ob.FIR_DECIM, rng.standard_normal(513).astype(np.float32), 256)
    assert plan["front"] == 2 and plan["H"] >= 513


# --------------------------------------------------------------------------------------------------
# section groups: emulate pass 1 (impulse-response dot products) / warp scan / look-back / pass 2 in f64
# --------------------------------------------------------------------------------------------------
def _step(sec_type, c, x, s):
    if sec_type == 1:
        y = x * c[0] + s[0]
        return y, np.array([x * c[1] + s[1] - c[3] * y, x * c[2] - c[4] * y])
    if sec_type == 2:
        y = x - s[0] + c[0] * s[1]
        return y, np.array([x, y])
    y = c[0] * s[0] + c[1] * x
    return y, np.array([y, 0.0])


def _cascade(secs, x, st):
    """One input sample through the cascade; st is the (2*count) state vector."""
    st = st.copy()
    v = x
    for q, (t, c) in enumerate(secs):
        v, st[2 * q:2 * q + 2] = _step(t, c, v, st[2 * q:2 * q + 2])
    return v, st


def _f32(c):
    out = np.zeros(5)
    out[:len(c)] = np.asarray(c, np.float32).astype(np.float64)
    return out


_BQ = lambda: list(ob.lp_biquad_design(48e3, 4.5e3).astype(np.float64))   # noqa: E731
_GROUPS = {
    "lr4": lambda: [(1, _BQ()), (1, _BQ())],
    "bq_dc": lambda: [(1, _BQ()), (2, [0.99973822])],
    "dc": lambda: [(2, [0.99973822])],
    "dc_clamped": lambda: [(2, [0.9999])],
    "onepole": lambda: [(3, [0.98, 0.02])],
}


@pytest.mark.parametrize("name", list(_GROUPS))
@pytest.mark.parametrize("npt", [8, 16, 2])
def test_group_tables_stitch_chunks_exactly(name, npt, rng):
    secs = [(t, _f32(c)) for t, c in _GROUPS[name]()]
    T = ob.debug_group_tables([(t, list(c)) for t, c in secs], npt)
    D = T["D"]
    assert D == 2 * len(secs)
    ntiles, tile_items = 5, KTHREADS * npt
    x = rng.standard_normal(ntiles * tile_items)
    s_carry = rng.standard_normal(D) * 0.1
    for q, (t, _) in enumerate(secs):
        if t == 3:
            s_carry[2 * q + 1] = 0.0
    # sequential truth
    st = s_carry.copy()
    want = np.zeros_like(x)
    for n in range(x.size):
        want[n], st = _cascade(secs, x[n], st)
    # chunked, exactly as the kernel stitches it
    got = np.zeros_like(x)
    aggs = []
    prev_incl = None
    lv, lane_m, lb = T["lv"].astype(np.float64), T["lane"].astype(np.float64), T["lb"].astype(np.float64)
    for t in range(ntiles):
        xs = x[t * tile_items:(t + 1) * tile_items].reshape(KTHREADS, npt)
        e = np.zeros((KTHREADS, D))
        for th in range(KTHREADS):
            z = np.zeros(D)
            for i in range(npt):
                _, z = _cascade(secs, xs[th, i], z)
            e[th] = z
        e_dot = xs @ T["imp"].astype(np.float64)          # the kernel's pass 1
        assert np.allclose(e_dot, e, rtol=1e-5, atol=1e-6 * max(1.0, np.max(np.abs(e))))
        E = e_dot.copy()
        for l in range(5):                                 # Kogge-Stone across the warp
            d = 1 << l
            prev = E.copy()
            for ln in range(d, 32):
                E[ln] = prev[ln] + lv[l] @ prev[ln - d]
        X = np.zeros_like(E)
        X[1:] = E[:-1]
        aggs.append(E[31].copy())
        sin = np.zeros(D)
        for k in range(min(t + 1, T["depth"])):           # predecessors past `depth` weigh nothing
            pay = aggs[t - 1 - k] if t - 1 - k >= 0 else s_carry
            sin = sin + lb[k] @ pay
        incl = T["tile"].astype(np.float64) @ sin + aggs[-1]
        if prev_incl is not None:
            assert np.allclose(sin, prev_incl, rtol=1e-5, atol=1e-6)
        prev_incl = incl
        for th in range(KTHREADS):
            z = X[th] + lane_m[th] @ sin
            for i in range(npt):
                got[t * tile_items + th * npt + i], z = _cascade(secs, xs[th, i], z)
    scale = max(1.0, np.max(np.abs(want)))
    assert np.max(np.abs(got - want)) < 2e-5 * scale
    # look-back depth: fast poles need a handful of predecessor tiles, the DC pole hundreds
    if name == "lr4":
        assert 1 <= T["depth"] <= 8 and T["agg_only"]
    if name.startswith("dc") or name == "bq_dc":
        r = float(np.float32(secs[-1][1][0]))
        Tt = KTHREADS * npt
        assert T["depth"] > 32 and not T["agg_only"]
        assert r ** (Tt * T["depth"]) < 1e-29 and r ** (Tt * max(T["depth"] - 2, 0)) > 1e-31


def test_rust_ffi_declares_every_symbol():
    """ffi/orion-b200-sys/src/lib.rs (source only: no Rust toolchain here) must declare every non-debug
    entry point of the header, and nothing the header does not have."""
    hdr = open(os.path.join(ROOT, "include", "orion_b200.h")).read()
    declared = set(re.findall(r"\b(orion_b200_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "ffi", "orion-b200-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (orion_b200_[a-z0-9_]+)\s*\(", rs))
    assert bound <= declared, sorted(bound - declared)
    missing = {s for s in declared - bound if "_debug_" not in s}
    assert not missing, sorted(missing)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under the package, include/ or ffi/ may import, link or
    load it (a product path through the oracle, or any CPU fallback, would void every parity claim)."""
    bad = []
    for top in ("orion-sdr_b200", "include", "ffi"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            if os.sep + "build" in d or os.sep + "lib" in d[len(ROOT):] and d.endswith("lib") or "variants" in d or "__pycache__" in d:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".rs", ".toml", "Makefile")):
                    txt = open(os.path.join(d, f), errors="ignore").read()
                    if re.search(r"\boracle\b|orion_oracle|np_oracle", txt):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


@pytest.mark.parametrize("sps", [0, 1, 2, 32, 255, 1024])
def test_half_cosine_taps_match_oracle(sps):
    a, b = ob.half_cosine_mf_taps(sps), oracle.half_cosine_taps(sps)
    assert a.size == b.size == max(sps, 1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert abs(float(np.sum(a.astype(np.float64) ** 2)) - 1.0) < 1e-5


def test_plain_c_client_compiles_links_and_runs(tmp_path):
    """include/orion_b200.h is usable from C99 and the shared library links from a C program (tests/c_abi/host_only.c)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "host_only")
    libdir = os.path.dirname(ob.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "host_only.c"), "-o", exe,
                           "-L", libdir, "-lorion_b200", "-lm", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok" in out.stdout


def test_numa_pinning_helper_is_best_effort():
    """bench.py binds every rank to the NUMA node of its GPU before it allocates host buffers.  Without a GPU (or without
    /sys NUMA information) the helper must report that and leave the affinity alone -- never raise."""
    import os
    import orion_b200 as ob
    before = os.sched_getaffinity(0)
    info = ob.pin_host_to_device_numa_node(0)
    assert set(info) >= {"node", "cpus", "pinned"}
    if not info["pinned"]:
        assert os.sched_getaffinity(0) == before
    else:
        assert os.sched_getaffinity(0) <= before and info["cpus"] == len(os.sched_getaffinity(0))
        os.sched_setaffinity(0, before)
