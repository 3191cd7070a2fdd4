import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "orion-sdr_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import orion_b200 as ob
n = 24_000_000
taps = ob.fir_lowpass_design(2.4e6, 100e3, 38400.0)
which = sys.argv[1] if len(sys.argv) > 1 else "dec"
if which == "dec":
    blk, out_items = ob.FirDecimator(2.4e6, 8, 100e3, 38400.0), n // 8
    y = torch.empty(2 * out_items, device="cuda")
else:
    blk, out_items = ob.Chain(fir=ob.FIR_DECIM, taps=taps, decim=8, demod=ob.DEMOD_FM, fs_demod=3e5, p0=25e3, audio_bw_hz=15e3, translate_hz=100e3), n // 8
    y = torch.empty(out_items, device="cuda")
x = torch.randn(2 * n, device="cuda")
ntiles = -(-out_items // 256)
tr = torch.zeros(ntiles * 16 + 8, dtype=torch.int64, device="cuda")
tr[ntiles * 16] = 2 ** 62
st = torch.cuda.Stream(); blk.set_stream(st.cuda_stream)
for _ in range(3):
    blk.process_dev(x.data_ptr(), n, y.data_ptr(), out_items)
blk.synchronize()
blk.set_trace(tr.data_ptr())
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
h0 = time.perf_counter()
blk.process_dev(x.data_ptr(), n, y.data_ptr(), out_items)
h1 = time.perf_counter()
e1.record(st)
blk.synchronize()
raw = tr.cpu().numpy()
print("single launch: events %.1f us, host enqueue %.1f us, globaltimer span %.1f us" % (e0.elapsed_time(e1) * 1e3, (h1 - h0) * 1e6, (raw[ntiles * 16 + 1] - raw[ntiles * 16]) / 1e3))
blk.set_trace(0)
h0 = time.perf_counter()
for _ in range(20):
    blk.process_dev(x.data_ptr(), n, y.data_ptr(), out_items)
h1 = time.perf_counter()
blk.synchronize()
h2 = time.perf_counter()
print("20 launches: host enqueue %.1f us each, total wall %.1f us each" % ((h1 - h0) / 20 * 1e6, (h2 - h0) / 20 * 1e6))
t = raw[:ntiles * 16].reshape(-1, 16)
k0, k1 = raw[ntiles * 16], raw[ntiles * 16 + 1]
g = t[:, 5]
print("kernel span %.1f us; first consume at +%.1f us, last consume at +%.1f us (kernel end +%.1f us)" % ((k1 - k0) / 1e3, (g.min() - k0) / 1e3, (g.max() - k0) / 1e3, (k1 - k0) / 1e3))
qs = np.percentile(g - k0, [1, 10, 25, 50, 75, 90, 99]) / 1e3
print("consume-time percentiles (us):", [round(float(v), 1) for v in qs])
order_t = np.argsort(g)
print("tile index of first 10 consumed:", order_t[:10].tolist(), " last 10:", order_t[-10:].tolist())
spans = []
for s_ in np.unique(t[:, 6]):
    m_ = t[:, 6] == s_
    spans.append((g[m_].max() - g[m_].min()) / 1e3)
print("per-SM consume span us: min %.1f median %.1f max %.1f" % (min(spans), float(np.median(spans)), max(spans)))
ok = t[:, 0] > 0
t = t[ok]
sm = t[:, 6]
print("tiles traced", t.shape[0])
def stats(name, v):
    v = v[np.isfinite(v)]
    print(f"{name:28s} mean {v.mean():9.0f}  p10 {np.percentile(v,10):9.0f}  p50 {np.percentile(v,50):9.0f}  p90 {np.percentile(v,90):9.0f} cycles")
stats("wait for slot (ready-consume)", (t[:, 1] - t[:, 0]).astype(float))
stats("FIR (fir_done-ready)", (t[:, 2] - t[:, 1]).astype(float))
stats("front rest (front-fir_done)", (t[:, 3] - t[:, 2]).astype(float))
stats("  halo+refill (8-2)", (t[:, 8] - t[:, 2]).astype(float))
stats("  front_map (9-8)", (t[:, 9] - t[:, 8]).astype(float))
stats("  group_front (3-9)", (t[:, 3] - t[:, 9]).astype(float))
stats("    dot (11-9)", (t[:, 11] - t[:, 9]).astype(float))
stats("    scan (12-11)", (t[:, 12] - t[:, 11]).astype(float))
stats("    publish (13-12)", (t[:, 13] - t[:, 12]).astype(float))
stats("    park (3-13)", (t[:, 3] - t[:, 13]).astype(float))
if which != "dec":
    f = t[:, 4] > 0
    stats("issue->finish (pending)", (t[f, 4] - t[f, 3]).astype(float))
    f2 = f & (t[:, 10] > 0) & (t[:, 14] > 0)
    if f2.any():
        stats("  front done -> finish entered", (t[f2, 14] - t[f2, 3]).astype(float))
        stats("  look-back (entered -> done)", (t[f2, 10] - t[f2, 14]).astype(float))
        stats("  recursion + store (-> finish done)", (t[f2, 4] - t[f2, 10]).astype(float))
# per-warp iteration time: sort by (warp id), consecutive consume stamps
w = t[:, 7]
order = np.lexsort((t[:, 0], w))
tw, ww = t[order], w[order]
same = ww[1:] == ww[:-1]
stats("warp iteration (consume->next consume)", (tw[1:, 0] - tw[:-1, 0])[same].astype(float))
# kernel span on one SM
for s in np.unique(sm)[:2]:
    m = sm == s
    ts = t[m]
    t0 = ts[:, 0].min()
    print("SM", int(s), "tiles", int(m.sum()), "span cycles", int(ts[:, 3].max() - t0), flush=True)
    for wv in np.unique(ts[:, 7])[:3]:
        mm = ts[ts[:, 7] == wv]
        mm = mm[np.argsort(mm[:, 0])]
        print("  warp", int(wv), "tiles", mm.shape[0], "consume stamps (rel):", [int(v - t0) for v in mm[:, 0]][:14], flush=True)
        print("       ready-consume:", [int(v) for v in (mm[:, 1] - mm[:, 0])][:14], flush=True)
        print("       fir:", [int(v) for v in (mm[:, 2] - mm[:, 1])][:14], flush=True)
