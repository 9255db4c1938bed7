"""e2e leg of bench.py in isolation: submit path at several K, with the copy / kernel event profile (measurement aid)."""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
rfm = load_package()
import importlib
synth_device = importlib.import_module("radiofm_b200.synth_device")
FS, DS, BLK, S = 2.4e6, 11, 65472, 4096
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
iq = synth_device.make_batch_u8(torch, S, FS, 2 * BLK, dev)
dec = rfm.FmDecoderBatch(FS, -0.15 * FS, downsample=DS, n_streams=S, max_block_len=BLK, device=0, n_groups=G)
stride = dec.max_audio_floats(BLK)
h_iq = torch.empty((2, S, BLK, 2), dtype=torch.uint8).pin_memory()
h_iq.copy_(iq.view(S, 2, BLK, 2).permute(1, 0, 2, 3))
h_audio = torch.empty((2, S, stride), dtype=torch.float32).pin_memory()
lib = rfm.lib(); kout = C.c_uint32(0)
def submit(i):
    rc = lib.rfm_decoder_submit_u8(dec._h, C.cast(h_iq[i % 2].data_ptr(), C.POINTER(C.c_uint8)), BLK,
                                   C.cast(h_audio[i % 2].data_ptr(), C.POINTER(C.c_float)), stride, C.byref(kout))
    assert rc == 0, lib.rfm_last_error()
for i in range(2): submit(i)
dec.synchronize()
for K in (8, 8, 24):
    t0 = time.perf_counter()
    for i in range(K): submit(i)
    t1 = time.perf_counter()
    dec.synchronize()
    t2 = time.perf_counter()
    dec.take_groups(0)
    t3 = time.perf_counter()
    print(f"G={G} K={K}: {(t3-t0)/K*1e3:.2f} ms/step ({S*BLK*K/(t3-t0)/1e9:.2f} GS/s)  enqueue {(t1-t0)/K*1e3:.2f} ms/step, sync tail {(t2-t1)*1e3:.2f} ms, take_groups {(t3-t2)*1e3:.2f} ms")
dec.set_profiling(True)
for i in range(8): submit(i)
dec.synchronize()
p = dec.profile()
print({k: (round(v[0] / 8, 3), v[1]) for k, v in sorted(p.items(), key=lambda kv: -kv[1][0])[:8]})
