"""Per-stage parity diagnostic (GPU box): CUDA chain vs the oracle on every tap, block by block."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
from oracle import port

rfm = conftest.load_package()
TAPS = ["demod_in", "baseband", "rawstereo", "mono_rs", "stereo_rs", "lp_stereo", "lp_mono", "rds_dec", "rds_lp", "rds_pll", "rds_mf"]
OT = {"lp_stereo": ("lp", 0), "lp_mono": ("lp", 1)}

def run(rate, nblocks, S=1):
    fs, ds, blk = conftest.RATES[rate]
    iq, groups = conftest.station(rate, nblocks)
    o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
    d = rfm.FmDecoderBatch(fs, -0.15 * fs, downsample=ds, n_streams=S, max_block_len=blk)
    c_o, c_d = o.constants(), d.constants()
    print(rate, "constants equal:", np.array_equal(c_o[:51], c_d[:51]), np.nonzero(c_o[:51] != c_d[:51])[0])
    for w in range(6):
        print("  table", w, np.array_equal(o.table(w).view(np.uint32), d.table(w).view(np.uint32)))
    worst = 0.0
    for b in range(nblocks):
        x = iq[b * blk:(b + 1) * blk]
        a_o = o.process_u8(x)
        a_d = d.process_u8(np.broadcast_to(x, (S,) + x.shape))
        line = []
        for t in TAPS:
            if t in OT:
                name, half = OT[t]
                full = o.tap(name); na = full.size // 2
                to = full[half * na:(half + 1) * na]
            else:
                to = o.tap(t)
            td = d.tap(t, S - 1)
            if to.shape != td.shape:
                line.append(f"{t}:SHAPE{to.shape}{td.shape}")
                continue
            ne = int(np.count_nonzero(to.view(np.uint32) != td.view(np.uint32)))
            line.append(f"{t}:{ne}" + (f"({np.max(np.abs(to - td)):.1e})" if ne else ""))
        ok = a_o.shape == a_d[S - 1].shape
        err = float(np.max(np.abs(a_o - a_d[S - 1]))) if ok and a_o.size else -1
        worst = max(worst, err)
        st_o, st_d = o.status(), d.status(S - 1)
        print(f" blk {b}: audio n={a_o.size} shape_ok={ok} maxerr={err:.2e} stereo={st_o['stereo']}/{st_d['stereo']} "
              f"status_eq={all(np.float32(st_o[k]) == np.float32(st_d[k]) for k in st_o)} | " + " ".join(line))
    g_o, g_d = o.take_groups(), d.take_groups(S - 1)
    b_o, b_d = o.take_bits(), d.take_bits(S - 1)
    print(f" groups oracle={len(g_o)} gpu={len(g_d)} equal={np.array_equal(g_o, g_d)}; bits {b_o.size}/{b_d.size} equal={np.array_equal(b_o, b_d)}; worst audio err {worst:.2e}")
    if len(g_o):
        print("  first groups:", g_o[:3].tolist(), " sent:", groups[:3].tolist())

if __name__ == "__main__":
    print(rfm.lib().rfm_version())
    for rate, nb in (("1.0M", 12), ("1.2M", 12), ("2.4M", 20), ("390k", 8)):
        t = time.time()
        run(rate, nb, S=1)
        print("  time", time.time() - t)
    run("1.0M", 3, S=70)
