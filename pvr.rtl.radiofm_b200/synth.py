"""Synthetic FM-broadcast IQ generator (u8 offset-binary, RTL-SDR style).

Test-signal infrastructure shared by tests/ and bench.py.  Recipe: SURVEY.md Appendix B
(the signal the reference chain was shown to decode):

    mpx(t) = 1/2 (L+R) + 1/2 (L-R) sin(2 wp t) + 0.1 sin(wp t) + 0.05 s(t) sin(3 wp t)
    phi[n] = 2 pi dev/Fs * cumsum(mpx)[n] + 2 pi f_off n / Fs
    I,Q    = clip(round(127.5 + 127.5 A cos/sin phi), 0, 255)

s(t) is the RDS baseband: differentially encoded bits, biphase impulse pairs shaped by the
IEC 62106 cosine filter, bit rate 57000/48 = 1187.5 bit/s, carrier locked to 3x pilot.
Block check words: CRC g(x)=0x5B9 XOR offset word (same as the reference's CRC_POLY,
/root/reference/src/RDSProcess.h:33).
"""
from __future__ import annotations

import numpy as np

PILOT_HZ = 19000.0
RDS_BITRATE = 57000.0 / 48.0
OFFSET_WORDS = {"A": 0x0FC, "B": 0x198, "C": 0x168, "Cp": 0x350, "D": 0x1B4}


def rds_checkword(info16: int, offset: int) -> int:
    """10-bit check word for a 16-bit information word (MSB first), XOR the offset word."""
    reg = 0
    for i in range(15, -1, -1):
        bit = (info16 >> i) & 1
        fb = ((reg >> 9) & 1) ^ bit
        reg = (reg << 1) & 0x3FF
        if fb:
            reg ^= 0x5B9 & 0x3FF
    return reg ^ offset


def rds_block_bits(info16: int, offset: int) -> list[int]:
    word = ((info16 & 0xFFFF) << 10) | rds_checkword(info16 & 0xFFFF, offset)
    return [(word >> i) & 1 for i in range(25, -1, -1)]


def rds_group_0a(pi: int, ps: str, segment: int, pty: int = 5, tp: int = 0, ms: int = 1) -> list[int]:
    """One type-0A group as 4 x u16 block words (A, B, C, D)."""
    ps = (ps + " " * 8)[:8]
    b = (0 << 12) | (0 << 11) | (tp << 10) | (pty << 5) | (ms << 3) | (segment & 3)
    c = 0xE0CD
    d = (ord(ps[2 * segment]) << 8) | ord(ps[2 * segment + 1])
    return [pi & 0xFFFF, b, c, d]


def rds_group_2a(pi: int, text: str, segment: int, pty: int = 5, tp: int = 0, ab: int = 0) -> list[int]:
    """One type-2A (RadioText) group; 4 characters per segment."""
    text = (text + "\r" + " " * 64)[:64]
    b = (2 << 12) | (0 << 11) | (tp << 10) | (pty << 5) | (ab << 4) | (segment & 15)
    c = (ord(text[4 * segment]) << 8) | ord(text[4 * segment + 1])
    d = (ord(text[4 * segment + 2]) << 8) | ord(text[4 * segment + 3])
    return [pi & 0xFFFF, b, c, d]


def rds_group_stream(pi: int = 0xD314, ps: str = "TESTB200", n_groups: int = 128,
                     radiotext: str | None = None) -> np.ndarray:
    """Known group stream: cyclic 0A groups (PS segments 0..3), optionally interleaved 2A."""
    groups = []
    seg2 = 0
    for g in range(n_groups):
        if radiotext is not None and g % 3 == 2:
            groups.append(rds_group_2a(pi, radiotext, seg2 % 16))
            seg2 += 1
        else:
            groups.append(rds_group_0a(pi, ps, g % 4))
    return np.asarray(groups, dtype=np.uint16)


def rds_bits_from_groups(groups: np.ndarray) -> np.ndarray:
    bits: list[int] = []
    for a, b, c, d in np.asarray(groups, dtype=np.uint16).tolist():
        version_b = (b >> 11) & 1
        bits += rds_block_bits(a, OFFSET_WORDS["A"])
        bits += rds_block_bits(b, OFFSET_WORDS["B"])
        bits += rds_block_bits(c, OFFSET_WORDS["Cp"] if version_b else OFFSET_WORDS["C"])
        bits += rds_block_bits(d, OFFSET_WORDS["D"])
    return np.asarray(bits, dtype=np.uint8)


def _rds_shape(x: np.ndarray) -> np.ndarray:
    """h(t) = cos(4 pi t/T) / (1 - (8 t/T)^2), x = t/T; removable singularity at |x| = 1/8."""
    x = np.asarray(x, dtype=np.float64)
    den = 1.0 - (8.0 * x) ** 2
    sing = np.abs(den) < 1e-9
    den = np.where(sing, 1.0, den)
    h = np.cos(4.0 * np.pi * x) / den
    h = np.where(sing, np.pi / 4.0, h)
    return np.where(np.abs(x) <= 1.0, h, 0.0)  # truncated to +-T (tails < 1.6 %)


def rds_baseband(bits: np.ndarray, fs: float, n: int, t0: float = 0.0) -> np.ndarray:
    """s(t) sampled at fs for n samples starting at time t0, normalised to peak 1."""
    T = 1.0 / RDS_BITRATE
    enc = np.cumsum(bits.astype(np.int64)) & 1  # differential encoding e[k] = e[k-1] ^ b[k]
    amp = 2.0 * enc - 1.0
    s = np.zeros(n, dtype=np.float64)
    span = int(np.ceil(1.6 * T * fs))
    for k, a in enumerate(amp):
        tk = k * T - t0
        c = int(round(tk * fs))
        lo, hi = max(0, c - span), min(n, c + span + int(T * fs))
        if hi <= 0 or lo >= n:
            if lo >= n:
                break
            continue
        t = (np.arange(lo, hi) / fs - tk) / T
        s[lo:hi] += a * (_rds_shape(t) - _rds_shape(t - 0.5))
    grid = np.linspace(-1.5, 1.5, 6001)
    peak = float(np.max(np.abs(_rds_shape(grid) - _rds_shape(grid - 0.5))))
    return s / peak


def mpx_signal(fs: float, n: int, *, left=(1000.0, 0.2), right=(3000.0, 0.2), pilot: float = 0.1,
               rds_level: float = 0.0, rds_bits: np.ndarray | None = None, mono_tone=None) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / fs
    wp = 2.0 * np.pi * PILOT_HZ * t
    if mono_tone is not None:
        f, a = mono_tone
        return a * np.sin(2.0 * np.pi * f * t)
    L = left[1] * np.sin(2.0 * np.pi * left[0] * t)
    R = right[1] * np.sin(2.0 * np.pi * right[0] * t)
    mpx = 0.5 * (L + R) + 0.5 * (L - R) * np.sin(2.0 * wp) + pilot * np.sin(wp)
    if rds_level > 0.0 and rds_bits is not None:
        mpx = mpx + rds_level * rds_baseband(rds_bits, fs, n) * np.sin(3.0 * wp)
    return mpx


def fm_modulate_u8(mpx: np.ndarray, fs: float, f_off: float, *, deviation: float = 75000.0,
                   amplitude: float = 0.8, snr_db: float | None = None, seed: int = 1234) -> np.ndarray:
    """FM-modulate mpx onto a carrier f_off Hz from the LO; returns u8 array [n, 2] (I, Q)."""
    n = mpx.shape[0]
    phi = 2.0 * np.pi * deviation / fs * np.cumsum(mpx) + 2.0 * np.pi * f_off * np.arange(n) / fs
    i = amplitude * np.cos(phi)
    q = amplitude * np.sin(phi)
    if snr_db is not None:
        rng = np.random.default_rng(seed)
        sigma = amplitude / np.sqrt(2.0) * 10.0 ** (-snr_db / 20.0)
        i = i + sigma * rng.standard_normal(n)
        q = q + sigma * rng.standard_normal(n)
    out = np.empty((n, 2), dtype=np.uint8)
    out[:, 0] = np.clip(np.rint(127.5 + 127.5 * i), 0, 255).astype(np.uint8)
    out[:, 1] = np.clip(np.rint(127.5 + 127.5 * q), 0, 255).astype(np.uint8)
    return out


def stream_params(stream_id: int) -> dict:
    """Per-stream content for the C4 batch: tones / PI / PS from the seed."""
    rng = np.random.default_rng(1234 + stream_id)
    fl = float(rng.integers(3, 40)) * 100.0
    fr = float(rng.integers(3, 40)) * 100.0
    pi = int(rng.integers(0x1000, 0xFFFF))
    ps = "S%07d" % (stream_id % 10_000_000)
    return {"left": (fl, 0.2), "right": (fr, 0.2), "pi": pi, "ps": ps}


def make_station_u8(fs: float, n: int, *, stream_id: int = 0, f_off: float | None = None,
                    stereo: bool = True, rds: bool = True, mono_tone=None,
                    left=None, right=None, pi: int | None = None, ps: str | None = None,
                    snr_db: float | None = None, n_groups: int | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Full synthetic station.  Returns (iq_u8[n,2], groups[n_groups,4]) -- groups transmitted."""
    p = stream_params(stream_id)
    if stream_id == 0:
        p.update({"left": (1000.0, 0.2), "right": (3000.0, 0.2), "pi": 0xD314, "ps": "TESTB200"})
    if left is not None:
        p["left"] = left
    if right is not None:
        p["right"] = right
    if pi is not None:
        p["pi"] = pi
    if ps is not None:
        p["ps"] = ps
    if f_off is None:
        f_off = -0.15 * fs  # reference convention: tuner LO = station + 0.15 Fs (RadioReceiver.cpp:237)
    groups = np.zeros((0, 4), dtype=np.uint16)
    bits = None
    if rds and mono_tone is None:
        if n_groups is None:
            n_groups = int(np.ceil(n / fs * RDS_BITRATE / 104.0)) + 2
        groups = rds_group_stream(p["pi"], p["ps"], n_groups)
        bits = rds_bits_from_groups(groups)
    mpx = mpx_signal(fs, n, left=p["left"], right=p["right"], pilot=0.1 if stereo else 0.0,
                     rds_level=0.05 if rds else 0.0, rds_bits=bits, mono_tone=mono_tone)
    iq = fm_modulate_u8(mpx, fs, f_off, snr_db=snr_db, seed=1234 + stream_id)
    return iq, groups


def make_wideband_u8(fs: float, n: int, station_freqs, *, first_stream: int = 0, rds: bool = True,
                     amplitude: float = 0.8) -> np.ndarray:
    """One shared wideband capture: FM stations at station_freqs[k] Hz from the LO, equal
    power, each with the C3 multiplex of stream `first_stream + k`; the sum is scaled to `amplitude` of full scale.
    Returns u8 [n, 2]."""
    acc = np.zeros(n, dtype=np.complex128)
    for k, f in enumerate(station_freqs):
        p = stream_params(first_stream + k)
        bits = None
        if rds:
            groups = rds_group_stream(p["pi"], p["ps"], int(np.ceil(n / fs * RDS_BITRATE / 104.0)) + 2)
            bits = rds_bits_from_groups(groups)
        mpx = mpx_signal(fs, n, left=p["left"], right=p["right"], rds_level=0.05 if rds else 0.0, rds_bits=bits)
        phi = 2.0 * np.pi * 75000.0 / fs * np.cumsum(mpx) + 2.0 * np.pi * f * np.arange(n) / fs
        acc += np.exp(1j * phi)
    acc *= amplitude / max(1.0, float(np.max(np.abs(acc))))
    out = np.empty((n, 2), dtype=np.uint8)
    out[:, 0] = np.clip(np.rint(127.5 + 127.5 * acc.real), 0, 255).astype(np.uint8)
    out[:, 1] = np.clip(np.rint(127.5 + 127.5 * acc.imag), 0, 255).astype(np.uint8)
    return out
