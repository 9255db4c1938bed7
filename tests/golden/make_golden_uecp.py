"""Generates tests/golden/uecp_kat.npz from the UNMODIFIED reference cRDSGroupDecoder
(oracle/_ref/libradiofm_ref_uecp.so, built by `make -C oracle uecp` where /root/reference exists).

    python tests/golden/make_golden_uecp.py

Scenarios (each: groups [n, 4] u16 -> raw UECP frames in order, u16-length-prefixed, + accepted PS names):
  scripted   every group type the decoder acts on (0A/0B, 1A/1B, 2A/2B, 3A with RT+ / TFC / unknown AID, 4A, 8A,
             10A, ODA-claimed 11A/12A/5A/8A), the ones it ignores (5A..9A unclaimed, 13A, 14A/B, 15A/B), PI / PTY
             changes, PS / RT / PTYN updates, A/B flag toggles
  fuzz       20 000 groups with random B / C / D words and an occasional PI change (4A dates kept >= 1900-03-01 so
             that the reference's int conversions stay defined)
  dialog     the scripted stream with the settings dialog open for its middle third (IsSettingActive)
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def grp(pi, gtype, version_b, pty, low5, c, d, tp=0):
    b = ((gtype & 0xF) << 12) | ((1 if version_b else 0) << 11) | ((tp & 1) << 10) | ((pty & 0x1F) << 5) | (low5 & 0x1F)
    return [pi & 0xFFFF, b, c & 0xFFFF, d & 0xFFFF]


def txt(s, n):
    return s.encode("latin-1")[:n].ljust(n, b" ")


def scripted(rng):
    g = []
    pi, pty = 0xD314, 10

    def ps_cycle(name, ta=0, ms=1, di=0b0101, tp=1, ver_b=False):
        t = txt(name, 8)
        for seg in range(4):
            dibit = (di >> (3 - seg)) & 1
            low5 = (ta << 4) | (ms << 3) | (dibit << 2) | seg
            g.append(grp(pi, 0, ver_b, pty, low5, 0xE0CD if not ver_b else pi, (t[2 * seg] << 8) | t[2 * seg + 1], tp))

    def rt(text, ab, ver_b=False, nseg=None):
        width = 2 if ver_b else 4
        t = txt(text, 16 * width)
        nseg = nseg if nseg is not None else 16
        for seg in range(nseg):
            low5 = (ab << 4) | seg
            if ver_b:
                g.append(grp(pi, 2, True, pty, low5, pi, (t[2 * seg] << 8) | t[2 * seg + 1]))
            else:
                g.append(grp(pi, 2, False, pty, low5, (t[4 * seg] << 8) | t[4 * seg + 1], (t[4 * seg + 2] << 8) | t[4 * seg + 3]))

    ps_cycle("TESTB200")
    ps_cycle("TESTB200")                      # unchanged: nothing published, flags stay set
    ps_cycle("RADIO  1", ta=1, ms=0, di=0b1010)
    g.append(grp(pi, 1, False, pty, 0, 0x80E0, 0x1234))   # 1A: PIN + slow labelling
    g.append(grp(pi, 1, False, pty, 0, 0x30AB, 0x1234))   # same PIN
    g.append(grp(pi, 1, True, pty, 0, pi, 0x4321))        # 1B: PIN only
    rt("Now playing: Blackwell Blues - The Tensor Cores", 0)
    rt("Now playing: Blackwell Blues - The Tensor Cores", 0)   # complete -> RT frame at the second segment 0
    g.append(grp(pi, 3, False, pty, (11 << 1) | 0, 0x1234, 0x4BD7))  # 3A: RT+ on 11A
    g.append(grp(pi, 3, False, pty, (12 << 1) | 0, 0x0000, 0xCD46))  # 3A: TFC on 12A
    g.append(grp(pi, 3, False, pty, (5 << 1) | 0, 0x0001, 0x1111))   # 3A: unknown AID on 5A
    g.append(grp(pi, 3, False, pty, (8 << 1) | 0, 0x0002, 0xCD46))   # 3A: TFC claims 8A
    g.append(grp(pi, 11, False, pty, 0x15, 0xAAAA, 0x5555))          # RT+ data, not ready
    rt("Now playing: Blackwell Blues - The Tensor Cores", 0, nseg=1)  # segment 0 again -> RT frame, RT+ ready
    g.append(grp(pi, 11, False, pty, 0x15, 0xAAAA, 0x5555))          # RT+ data forwarded
    g.append(grp(pi, 11, False, pty, 0x15, 0xBBBB, 0x6666))          # not ready any more
    g.append(grp(pi, 12, False, pty, 0x0A, 0x0102, 0x0304))          # TFC data
    g.append(grp(pi, 8, False, pty, 0x07, 0xC0DE, 0xFEFF))           # 8A claimed by TFC (bytes >= 0xFD in the frame)
    g.append(grp(pi, 3, False, pty, (8 << 1) | 0, 0x0002, 0x2222))   # 3A: 8A released
    g.append(grp(pi, 8, False, pty, 0x07, 0xFDFE, 0xFF00))           # 8A native TMC
    g.append(grp(pi, 5, False, pty, 0x03, 0x1111, 0x2222))           # 5A unclaimed: ignored
    g.append(grp(pi, 6, True, pty, 0x03, 0x1111, 0x2222))
    g.append(grp(pi, 7, False, pty, 0x03, 0x1111, 0x2222))
    g.append(grp(pi, 9, False, pty, 0x03, 0x1111, 0x2222))
    g.append(grp(pi, 13, False, pty, 0x03, 0x1111, 0x2222))
    g.append(grp(pi, 14, False, pty, 0x03, 0x1111, 0x2222))
    g.append(grp(pi, 14, True, pty, 0x03, 0x1111, 0x2222))
    g.append(grp(pi, 15, False, pty, 0x03, 0x1111, 0x2222))
    g.append(grp(pi, 15, True, pty, 0x03, 0x1111, 0x2222))
    # 4A: 2026-10-17 11:45 +02:00 (MJD 61330), and the K = 1 months (January / February)
    for mjd, hh, mm, off in ((61330, 11, 45, 4), (61060, 23, 59, 0x24), (58849, 0, 0, 0), (15079, 5, 6, 7)):
        low5 = (mjd >> 15) & 3
        c = ((mjd & 0x7FFF) << 1) | (hh >> 4)
        d = ((hh & 0xF) << 12) | (mm << 6) | off
        g.append(grp(pi, 4, False, pty, low5, c, d))
    t = txt("Football", 8)
    g.append(grp(pi, 10, False, pty, 0x00, (t[0] << 8) | t[1], (t[2] << 8) | t[3]))
    g.append(grp(pi, 10, False, pty, 0x01, (t[4] << 8) | t[5], (t[6] << 8) | t[7]))
    g.append(grp(pi, 10, False, pty, 0x10, 0x4142, 0x4344))          # A/B toggle
    pty = 3
    ps_cycle("RADIO  1", ta=1, ms=0, di=0b1010)                      # PTY change only
    rt("Short text", 1, nseg=4)                                      # A/B toggle, partial text
    rt("Short text", 1, nseg=4)                                      # 4 segments seen twice: count 8 -> not complete
    rt("B-version text 32 characters....", 0, ver_b=True)
    rt("B-version text 32 characters....", 0, ver_b=True, nseg=1)
    pi = 0x1A2B                                                      # PI change: Reset()
    ps_cycle("NEXT ONE", tp=0, ver_b=True)
    rt("after the PI change", 0)
    rt("after the PI change", 0, nseg=1)
    g.append(grp(pi, 11, False, pty, 0x15, 0xAAAA, 0x5555))          # ODA map was cleared by Reset
    ps_cycle("NEXT TWO", tp=0)
    g.append(grp(pi, 0, False, pty, (1 << 3) | 2, 0xE0CD, (ord("X") << 8) | ord("Y"), 0))  # one changed segment publishes
    return np.asarray(g, dtype=np.uint16)


def fuzz(rng, n=20000):
    g = np.zeros((n, 4), dtype=np.uint16)
    pi = 0x5000
    for i in range(n):
        if rng.random() < 0.002:
            pi = int(rng.integers(0, 1 << 16))
        b, c, d = (int(x) for x in rng.integers(0, 1 << 16, 3))
        if rng.random() < 0.7:
            b = (b & ~(0x1F << 5)) | (7 << 5)  # mostly stable PTY
        if (b >> 11) & 0x1F == 0x08:           # 4A: keep MJD >= 15079
            mjd = int(rng.integers(15079, 1 << 17))
            b = (b & ~3) | ((mjd >> 15) & 3)
            c = ((mjd & 0x7FFF) << 1) | (c & 1)
        if (b >> 11) & 0x1F == 0x06 and rng.random() < 0.6:   # 3A: known AIDs often
            d = 0x4BD7 if rng.random() < 0.5 else 0xCD46
        g[i] = (pi, b, c, d)
    return g


def pack(frames):
    out = bytearray()
    for f in frames:
        out += bytes([len(f) & 0xFF, len(f) >> 8]) + f
    return np.frombuffer(bytes(out), dtype=np.uint8)


def main():
    from oracle import ref_uecp
    if not ref_uecp.available():
        raise SystemExit("oracle/_ref/libradiofm_ref_uecp.so missing: make -C oracle uecp")
    rng = np.random.default_rng(20261017)
    out = {}
    s = scripted(rng)
    f = fuzz(rng)
    for name, groups in (("scripted", s), ("fuzz", f)):
        r = ref_uecp.RefGroupDecoder()
        out[f"{name}_groups"] = groups
        out[f"{name}_frames"] = pack(r.decode(groups))
        out[f"{name}_names"] = np.frombuffer(b"".join(r.take_names()), dtype=np.uint8)
    # dialog: middle third with IsSettingActive() true
    r = ref_uecp.RefGroupDecoder()
    n = s.shape[0]
    fr = r.decode(s[:n // 3])
    r.set_setting_active(True)
    fr += r.decode(s[n // 3:2 * n // 3])
    r.set_setting_active(False)
    fr += r.decode(s[2 * n // 3:])
    out["dialog_frames"] = pack(fr)
    out["dialog_names"] = np.frombuffer(b"".join(r.take_names()), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "uecp_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
