import torch, importlib, sys
sys.path.insert(0, ".")
from __graft_entry__ import load_package
rfm = load_package()
synth_device = importlib.import_module("radiofm_b200.synth_device"); wideband = importlib.import_module("radiofm_b200.wideband")
FS5, BLK5, BPC5 = 50.0e6, 32000, 64; n_call = BLK5 * BPC5; n_st = 100
freqs = [(k - n_st // 2) * 200000.0 for k in range(n_st)]
cap = synth_device.make_wideband_u8(torch, FS5, n_call, freqs, torch.device("cuda", 0))
wb = wideband.WidebandReceiver(torch, freqs, FS5, BLK5, BPC5, mixer="freqshift", device=0, lanes_sms=1)
for _ in range(3): wb.dc.process_device(1, cap.data_ptr(), n_call, wb.bb.data_ptr(), wb.n_bb, n_call)
torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(5): wb.dc.process_device(1, cap.data_ptr(), n_call, wb.bb.data_ptr(), wb.n_bb, n_call)
f1.record(); torch.cuda.synchronize(); print('front end alone ms %.4f' % (f0.elapsed_time(f1) / 5))
wb.close()
