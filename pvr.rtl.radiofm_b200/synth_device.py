"""Device-side synthetic batch generator (torch, float64 on the GPU) -- bench scaffolding only.

Same recipe as synth.py, vectorised over streams so that the C4 batch (4096 independent
2.4 MS/s stereo+RDS stations) can be created in HBM in seconds: every stream gets its own L/R tone
frequencies and start phases; the RDS basebands come from a small pool generated on the host by synth.py
(per-stream PI / PS for the first `rds_pool` streams, then reused with a per-stream delay).
"""
from __future__ import annotations

import numpy as np

from . import synth


def make_batch_u8(torch, n_streams: int, fs: float, n: int, device, *, first_stream: int = 0, rds_pool: int = 16,
                  chunk: int = 128):
    """-> uint8 tensor [n_streams, n, 2] on `device`."""
    pool = []
    for k in range(rds_pool):
        p = synth.stream_params(first_stream + k)
        groups = synth.rds_group_stream(p["pi"], p["ps"], int(np.ceil(n / fs * synth.RDS_BITRATE / 104.0)) + 3)
        pool.append(synth.rds_baseband(synth.rds_bits_from_groups(groups), fs, n + 4096))
    rds = torch.from_numpy(np.stack(pool)).to(device)  # [pool, n + 4096] float64
    out = torch.empty((n_streams, n, 2), dtype=torch.uint8, device=device)
    t = torch.arange(n, dtype=torch.float64, device=device) / fs
    wp = 2.0 * np.pi * synth.PILOT_HZ * t
    sin1, sin2, sin3 = torch.sin(wp), torch.sin(2.0 * wp), torch.sin(3.0 * wp)
    carrier = 2.0 * np.pi * (-0.15 * fs) * torch.arange(n, dtype=torch.float64, device=device) / fs
    rng = np.random.default_rng(4321 + first_stream)
    for s0 in range(0, n_streams, chunk):
        s1 = min(n_streams, s0 + chunk)
        m = s1 - s0
        fl = torch.from_numpy(rng.integers(3, 40, m) * 100.0).to(device)[:, None]
        fr = torch.from_numpy(rng.integers(3, 40, m) * 100.0).to(device)[:, None]
        ph = torch.from_numpy(rng.random((m, 2)) * 2.0 * np.pi).to(device)
        L = 0.2 * torch.sin(2.0 * np.pi * fl * t[None, :] + ph[:, :1])
        R = 0.2 * torch.sin(2.0 * np.pi * fr * t[None, :] + ph[:, 1:])
        idx = torch.arange(s0, s1, device=device) % rds_pool
        delay = torch.from_numpy(rng.integers(0, 4096, m)).to(device)
        gather = delay[:, None] + torch.arange(n, device=device)[None, :]
        r = torch.gather(rds[idx], 1, gather)
        mpx = 0.5 * (L + R) + 0.5 * (L - R) * sin2[None, :] + 0.1 * sin1[None, :] + 0.05 * r * sin3[None, :]
        del L, R, r, gather
        phi = 2.0 * np.pi * 75000.0 / fs * torch.cumsum(mpx, dim=1) + carrier[None, :]
        del mpx
        out[s0:s1, :, 0] = torch.clamp(torch.round(127.5 + 127.5 * 0.8 * torch.cos(phi)), 0, 255).to(torch.uint8)
        out[s0:s1, :, 1] = torch.clamp(torch.round(127.5 + 127.5 * 0.8 * torch.sin(phi)), 0, 255).to(torch.uint8)
        del phi
    return out


def make_wideband_u8(torch, fs: float, n: int, station_freqs, device, *, first_station: int = 0, rds_pool: int = 4,
                     rds_fs: float = 250000.0, amplitude: float = 0.8):
    """One shared wideband capture on the device: equal-power FM stations at
    station_freqs[k] Hz, stereo multiplex + RDS subcarrier each; the RDS basebands come from a small host-made pool at
    `rds_fs`, linearly interpolated to `fs`.  -> uint8 tensor [n, 2]."""
    t = torch.arange(n, dtype=torch.float64, device=device) / fs
    wp = 2.0 * np.pi * synth.PILOT_HZ * t
    sin1, sin2, sin3 = torch.sin(wp), torch.sin(2.0 * wp), torch.sin(3.0 * wp)
    n_lo = int(np.ceil(n / fs * rds_fs)) + 8
    pool = []
    for k in range(rds_pool):
        p = synth.stream_params(first_station + k)
        groups = synth.rds_group_stream(p["pi"], p["ps"], int(np.ceil(n / fs * synth.RDS_BITRATE / 104.0)) + 3)
        pool.append(synth.rds_baseband(synth.rds_bits_from_groups(groups), rds_fs, n_lo))
    rds = torch.from_numpy(np.stack(pool)).to(device)
    pos = t * rds_fs
    i0 = pos.floor().long()
    frac = pos - i0
    acc_i = torch.zeros(n, dtype=torch.float64, device=device)
    acc_q = torch.zeros(n, dtype=torch.float64, device=device)
    idx = torch.arange(n, dtype=torch.float64, device=device)
    for k, f in enumerate(station_freqs):
        p = synth.stream_params(first_station + k)
        L = p["left"][1] * torch.sin(2.0 * np.pi * p["left"][0] * t)
        R = p["right"][1] * torch.sin(2.0 * np.pi * p["right"][0] * t)
        r = rds[k % rds_pool]
        s = r[i0] * (1.0 - frac) + r[i0 + 1] * frac
        mpx = 0.5 * (L + R) + 0.5 * (L - R) * sin2 + 0.1 * sin1 + 0.05 * s * sin3
        phi = 2.0 * np.pi * 75000.0 / fs * torch.cumsum(mpx, dim=0) + (2.0 * np.pi * float(f) / fs) * idx
        acc_i += torch.cos(phi)
        acc_q += torch.sin(phi)
        del L, R, s, mpx, phi
    peak = float(torch.sqrt(acc_i * acc_i + acc_q * acc_q).max().item())
    g = amplitude / max(1.0, peak)
    out = torch.empty((n, 2), dtype=torch.uint8, device=device)
    out[:, 0] = torch.clamp(torch.round(127.5 + 127.5 * g * acc_i), 0, 255).to(torch.uint8)
    out[:, 1] = torch.clamp(torch.round(127.5 + 127.5 * g * acc_q), 0, 255).to(torch.uint8)
    return out
