#!/usr/bin/env python
"""bench_wideband.py -- BASELINE.json configs[4]: one shared 50 MS/s u8 capture with 100 FM
stations on a 200 kHz raster; per station mixer + CRDSDownConvert (7 x HB51 -> 390 625 S/s) + cFmDecoder, stations
sharded over the ranks (no collective; every rank reads the same capture).  The headline bench stays bench.py (C4).

    python bench_wideband.py [--mixer osc|freqshift] [--stations 100] [--steps K] [--warmup W]
    torchrun --nproc-per-node N ... bench_wideband.py --gpus N ...

One step = one demodulator call = 64 front-end blocks of 32000 capture samples (2.048 M samples, 41 ms of signal) for
every station of the rank.  value = station-samples / s = stations x capture samples / time (whole job).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS, BLK, BPC = 50.0e6, 32000, 64
N_CALL = BLK * BPC


def station_freqs(n):
    return [(k - n // 2) * 200000.0 for k in range(n)]


def cpu_leg(capture_host: np.ndarray, freqs, mixer: str, seconds: float = 10.0):
    from oracle.wideband import OracleStation
    cores = len(os.sched_getaffinity(0))
    st = [OracleStation(freqs[k % len(freqs)], FS, BLK, BPC, mixer=mixer) for k in range(cores)]
    done = 0
    with cf.ThreadPoolExecutor(cores) as pool:
        list(pool.map(lambda o: o.process_u8(capture_host[:N_CALL]), st))
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            list(pool.map(lambda o: o.process_u8(capture_host[:N_CALL]), st))
            done += cores
        dt = time.perf_counter() - t0
    return {"value": done * N_CALL / dt / 1e6, "unit": "MS/s (station-samples)", "cores": cores, "kind": "port",
            "sample": f"{cores} threads, one station each ({mixer} mixer + 7 x HB51 + cFmDecoder), {done} demodulator "
                      f"calls of {N_CALL} capture samples in {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--stations", type=int, default=100)
    ap.add_argument("--mixer", default="osc", choices=["osc", "freqshift", "freqshift_unfused"])
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    rfm = load_package()
    synth_device = importlib.import_module("radiofm_b200.synth_device")
    wideband = importlib.import_module("radiofm_b200.wideband")
    shard = importlib.import_module("radiofm_b200.shard")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    freqs = station_freqs(args.stations)
    lo, hi = shard.shard_range(args.stations, rank, world)
    mine = freqs[lo:hi]
    K, W = args.steps, args.warmup
    ncalls = min(K + W, 6)
    capture = synth_device.make_wideband_u8(torch, FS, ncalls * N_CALL, freqs, dev)   # same on every rank
    torch.cuda.synchronize()
    wb = wideband.WidebandReceiver(torch, mine, FS, BLK, BPC, mixer=args.mixer, device=local)

    def step(i):
        return wb.process_device(capture.data_ptr() + 2 * (i % ncalls) * N_CALL)

    for i in range(W):
        step(i)
    wb.wait()
    barrier()
    l0 = rfm.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        nfl = step(W + i)
    wb.wait()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = rfm.launch_count() - l0
    value = args.stations * N_CALL * K / (ms * 1e-3) / 1e6

    # component split (rank 0, one extra step each): front end alone, demodulator alone
    parts = {}
    if rank == 0:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        if args.mixer in ("osc", "freqshift"):
            wb.dc.process_device(1, capture.data_ptr(), N_CALL, wb.bb.data_ptr(), wb.n_bb, N_CALL)
        else:
            for b in range(BPC):
                wb.shift.reset()
                wb.shift.process_device(1, capture.data_ptr() + 2 * b * BLK, BLK, wb.mixed.data_ptr() + 8 * b * BLK, N_CALL, BLK)
            wb.dc.process_device(0, wb.mixed.data_ptr(), N_CALL, wb.bb.data_ptr(), wb.n_bb, N_CALL)
        ev[1].record()
        wb.dec.process_cf32_device(wb.bb.data_ptr(), wb.n_bb, wb.n_bb, wb.audio.data_ptr(), wb.audio_stride)
        wb.dec.wait(0)
        ev[2].record()
        wb.dec.synchronize()
        torch.cuda.synchronize()
        parts = {"front_end_ms": ev[0].elapsed_time(ev[1]), "demodulator_ms": ev[1].elapsed_time(ev[2])}
    groups = [wb.dec.take_groups(s).shape[0] for s in range(min(len(mine), 4))]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_leg(capture[:N_CALL].cpu().numpy(), freqs, args.mixer.split("_")[0])
    if rank == 0:
        audio_bytes = args.stations / world * nfl * 4
        algo_bytes = 2.0 * N_CALL + audio_bytes          # per rank per step: one read of the capture + the audio
        print(json.dumps({
            "metric": "demodulated station-samples MS/s (wideband capture, stations sharded over the GPUs)",
            "value": value, "unit": "MS/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5: one shared 50 MS/s u8 capture, {args.stations} FM stations on a 200 kHz raster, "
                                   f"{args.mixer} mixer + CRDSDownConvert (7 x HB51) + cFmDecoder per station; one step = "
                                   f"{BPC} front-end blocks of {BLK} samples", "stations": args.stations,
                       "stations_per_gpu": len(mine), "mixer": args.mixer, "capture_rate": FS},
            "realtime_factor": (N_CALL / FS) / (ms / K * 1e-3),
            "hbm_algorithmic_gbs": algo_bytes / (ms / K * 1e-3) / 1e9,
            "parts": parts, "gpu_launches": launches, "audio_floats_per_station_per_step": nfl,
            "rds_groups_first_stations": groups, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
