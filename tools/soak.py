"""Long parity soak on the GPU box (not a pytest): S distinct stations (clean, noisy, mono, noise-only, silence) x NB
blocks through the streaming submit path with stream groups, the blocking path and the device-pointer path, every
block's audio / bits / groups / UECP bytes / status compared bit for bit with the oracle restatement.
    python tools/soak.py [rate] [streams] [blocks]
"""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

rfm = load_package()
import importlib  # noqa: E402

synth = importlib.import_module("radiofm_b200.synth")
from oracle import port, uecp_port  # noqa: E402

RATES = {"1.0M": (1.0e6, 4, 65536), "1.2M": (1.2e6, 5, 65520), "2.4M": (2.4e6, 11, 65472)}
rate = sys.argv[1] if len(sys.argv) > 1 else "2.4M"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 48
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 24
fs, ds, blk = RATES[rate]
n = NB * blk
rng = np.random.default_rng(99)


def make(s):
    kind = s % 6
    if kind == 4:     # noise only
        return rng.integers(0, 256, (n, 2), dtype=np.uint8)
    if kind == 5:     # silence (constant mid-scale) then a station
        iq = synth.make_station_u8(fs, n, stream_id=s)[0]
        iq[: n // 3] = 127
        return iq
    snr = {0: None, 1: 30.0, 2: 15.0, 3: None}[kind]
    return synth.make_station_u8(fs, n, stream_id=s, snr_db=snr, stereo=(kind != 3), rds=(kind != 3))[0]


t0 = time.time()
with ThreadPoolExecutor(8) as ex:
    iqs = np.stack(list(ex.map(make, range(S))))
print(f"generated {S} stations x {NB} blocks @ {rate} in {time.time() - t0:.1f} s", flush=True)


def oracle_run(s):
    o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
    audio = [o.process_u8(iqs[s, b * blk:(b + 1) * blk]) for b in range(NB)]
    return audio, o.take_bits(), o.take_groups(), o.status()


t0 = time.time()
with ThreadPoolExecutor(16) as ex:
    want = list(ex.map(oracle_run, range(S)))
print(f"oracle done in {time.time() - t0:.1f} s", flush=True)


def check(tag, audio_blocks, d):
    bad = 0
    for s in range(S):
        wa, wb, wg, wst = want[s]
        for b in range(NB):
            a = audio_blocks[b][s]
            if a.shape != wa[b].shape or not np.array_equal(a.view(np.uint32), wa[b].view(np.uint32)):
                bad += 1
                print(f"  {tag}: audio differs stream {s} block {b}")
                break
        if not np.array_equal(d.take_bits(s), wb):
            bad += 1
            print(f"  {tag}: bits differ stream {s}")
        o = uecp_port.OracleGroupDecoder()
        wu = b"".join(uecp_port.stuff_frame(f) for f in o.decode(wg))[:16385 + 300]
        gu = d.take_uecp(s, 1 << 20)
        if not np.array_equal(d.take_groups(s), wg) or (len(wu) <= 16384 and gu != wu):
            bad += 1
            print(f"  {tag}: groups / uecp differ stream {s}")
        st = d.status(s)
        if any(np.float32(wst[k]) != np.float32(st[k]) for k in wst):
            bad += 1
            print(f"  {tag}: status differs stream {s}: {wst} {st}")
    print(f"{tag}: {'OK' if bad == 0 else 'FAILED'} ({S} streams x {NB} blocks, {sum(len(w[2]) for w in want)} groups)", flush=True)
    return bad


total = 0
# blocking host path, 4 stream groups
d = rfm.FmDecoderBatch(fs, -0.15 * fs, downsample=ds, n_streams=S, max_block_len=blk, n_groups=4)
out = [d.process_u8(iqs[:, b * blk:(b + 1) * blk]) for b in range(NB)]
total += check("blocking G=4", out, d)
d.close()

# streaming submit path, 8 groups, results read after the end
import torch  # noqa: E402

d = rfm.FmDecoderBatch(fs, -0.15 * fs, downsample=ds, n_streams=S, max_block_len=blk, n_groups=8)
stride = d.max_audio_floats(blk)
h_iq = torch.from_numpy(np.ascontiguousarray(iqs.reshape(S, NB, blk, 2).transpose(1, 0, 2, 3))).pin_memory()
h_audio = torch.zeros((NB, S, stride), dtype=torch.float32).pin_memory()
lib = rfm.lib()
k = C.c_uint32(0)
counts = []
for b in range(NB):
    rc = lib.rfm_decoder_submit_u8(d._h, C.cast(h_iq[b].data_ptr(), C.POINTER(C.c_uint8)), blk,
                                   C.cast(h_audio[b].data_ptr(), C.POINTER(C.c_float)), stride, C.byref(k))
    assert rc == 0, lib.rfm_last_error()
    counts.append(int(k.value))
d.synchronize()
out = [[h_audio[b, s, :counts[b]].numpy() for s in range(S)] for b in range(NB)]
total += check("submit G=8", out, d)
d.close()

# device-pointer path, one group, everything enqueued back to back
d = rfm.FmDecoderBatch(fs, -0.15 * fs, downsample=ds, n_streams=S, max_block_len=blk, n_groups=1)
d_iq = torch.from_numpy(iqs).cuda()
d_audio = torch.zeros((NB, S, stride), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
counts = []
with torch.cuda.stream(st):
    for b in range(NB):
        counts.append(d.process_u8_device(d_iq.data_ptr() + b * blk * 2, n, blk, d_audio[b].data_ptr(), stride, st.cuda_stream))
    d.wait(st.cuda_stream)
st.synchronize()
res = d_audio.cpu().numpy()
out = [[res[b, s, :counts[b]] for s in range(S)] for b in range(NB)]
total += check("device G=1", out, d)
d.close()
print("SOAK", "PASSED" if total == 0 else f"FAILED ({total})")
sys.exit(0 if total == 0 else 1)
