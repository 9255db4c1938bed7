"""Single-stream latency probe (the add-on's own call: 1 stream, 65536 samples at 1.0 MS/s, blocking, host buffers).
   python tools/probe_single.py [calls]      -- prints the median ms per call; run under ncu for the per-kernel split."""
import importlib.util, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("radiofm_b200", os.path.join(ROOT, "pvr.rtl.radiofm_b200", "__init__.py"),
                                              submodule_search_locations=[os.path.join(ROOT, "pvr.rtl.radiofm_b200")])
rfm = importlib.util.module_from_spec(spec); sys.modules["radiofm_b200"] = rfm; spec.loader.exec_module(rfm)
import importlib
import numpy as np
synth = importlib.import_module("radiofm_b200.synth")
fs, ds, blk = 1.0e6, 4, 65536
ncall = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iq, _ = synth.make_station_u8(fs, 12 * blk, stream_id=0)
x = ((iq.astype(np.float64) / (255.0 / 2.0) - 1.0).astype(np.float32)).reshape(12, blk, 2)
d = rfm.FmDecoderBatch(fs, -0.15 * fs, downsample=ds, n_streams=1, max_block_len=blk, device=0)
ts = []
for c in range(ncall):
    t0 = time.perf_counter(); d.process_cf32(x[c % 12][None]); ts.append((time.perf_counter() - t0) * 1e3)
print("single stream ms per call: median %.3f min %.3f" % (statistics.median(ts[6:]), min(ts[6:])))
d.close()
