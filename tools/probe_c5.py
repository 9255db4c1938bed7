"""C5 timeline probe: the freqshift-mixer wideband step (100 stations of one 50 MS/s capture) with the decoder's kernel
timeline (experiments library: RFM_LIB_PATH=...libradiofm_b200_exp.so RFM_DEBUG_TIMELINE=1) and event brackets around
the front end of every call.   python tools/probe_c5.py [steps] [stations]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
rfm = load_package()
synth_device = importlib.import_module("radiofm_b200.synth_device")
wideband = importlib.import_module("radiofm_b200.wideband")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_st = int(sys.argv[2]) if len(sys.argv) > 2 else 100
FS5, BLK5, BPC5 = 50.0e6, 32000, 64
n_call = BLK5 * BPC5
freqs = [(k - n_st // 2) * 200000.0 for k in range(n_st)]
dev = torch.device("cuda", 0)
capture = synth_device.make_wideband_u8(torch, FS5, 3 * n_call, freqs, dev)
wb = wideband.WidebandReceiver(torch, freqs, FS5, BLK5, BPC5, mixer=os.environ.get("MIXER", "freqshift"), device=0,
                               lanes_sms=int(os.environ.get("LANES_SMS", "8")), n_slots=int(os.environ.get("SLOTS", "3")))
for i in range(3):
    wb.process_device(capture.data_ptr() + 2 * (i % 3) * n_call)
wb.wait(); torch.cuda.synchronize()
wb.dec.set_profiling(True)
base = torch.cuda.Event(enable_timing=True); base.record(); 
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    wb.process_device(capture.data_ptr() + 2 * (i % 3) * n_call)
wb.wait(); e1.record(); torch.cuda.synchronize()
print("ms_per_step %.4f" % (e0.elapsed_time(e1) / steps))
prof = wb.dec.profile()
print({k: round(v[0] / steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])})
wb.close()
