"""CPU restatement of cRDSGroupDecoder (RDSGroupDecoder.cpp:136-1001, default build macros) and of the
transport framing of cRadioReceiver::AddUECPDataFrame (RadioReceiver.cpp:387-414).

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  Pure Python (11 groups/s per stream: small cases only).
Pin: tests/test_uecp.py checks it frame for frame against the compiled reference (oracle/ref_uecp.py) where that
library exists, and against tests/golden/uecp_kat.npz (generated from the compiled reference by
tests/golden/make_golden_uecp.py) everywhere.  The framing (stuff_frame) is checked on hand-made
vectors and, through oracle/demux_port.py, against RadioReceiver.cpp compiled in place (tests/test_ref_addon.py).
Members the reference leaves uninitialised start at zero (the compiled reference is constructed in zeroed storage).
"""
from __future__ import annotations

AID_RTPLUS, AID_TFC = 0x4BD7, 0xCD46


def crc16(data: bytes) -> int:
    """RDSGroupDecoder.cpp:965-981."""
    crc = 0xFFFF
    for d in data:
        crc = ((crc >> 8) | (crc << 8)) & 0xFFFF
        crc ^= d
        crc ^= (crc & 0xFF) >> 4
        crc ^= (crc << 12) & 0xFFFF
        crc ^= ((crc & 0xFF) << 5) & 0xFFFF
    return (~crc) & 0xFFFF


def stuff_frame(frame: bytes) -> bytes:
    """RadioReceiver.cpp:397-411."""
    out = bytearray([0xFE])
    for v in frame:
        if v < 0xFD:
            out.append(v)
        else:
            out += bytes([0xFD, (v & 3) - 1])
    out.append(0xFF)
    return bytes(out)


def _c_int(x: float) -> int:
    """(int)double for in-range values: truncation toward zero."""
    return int(x)


class OracleGroupDecoder:
    def __init__(self, accept_name=True):
        self.setting_active = False
        self.accept_name = accept_name
        self.names: list[bytes] = []
        self.frames: list[bytes] = []
        # never initialised by the reference (zeroed storage)
        self.seq = 0
        self.pty = 0
        self.di_finished = 0
        self.rt_ab = 0
        self.ptyn_ab = 0
        self.ps_text = bytearray(9)      # function-static in the reference (:311)
        # zeroed storage for what Reset() will set on the first PI
        self.pi = 0
        self.rt_seg = 0
        self.rt_count = 0
        self.rt_first = False
        self.di = 0
        self.di_prev = 0
        self.ms = 0
        self.ms_prev = 0
        self.pin = 0
        self.ptyn_set = 0
        self.ps_name = bytearray(9)
        self.ps_set = 0
        self.ta_tp = 0
        self.rtplus_ready = False
        self.rt = bytearray(66)
        self.oda = [0] * 32
        self.ptyn = bytearray(9)

    def reset(self):
        """:140-164"""
        self.pi = 0
        self.rt_seg = 0
        self.rt_count = 0
        self.rt_first = False
        self.di = 0
        self.di_prev = 0xFF
        self.ms = 0
        self.ms_prev = 0xFF
        self.pin = 0xFFFF
        self.ptyn_set = 0
        self.ps_set = 0
        self.ta_tp = -1
        self.rtplus_ready = False
        self.rt = bytearray(66)
        self.oda = [0] * 32
        self.ptyn = bytearray(b"\x20" * 9)
        self.ps_name = bytearray(b"\x20" * 9)

    # ---- frame plumbing (:945-1001)
    def _send(self, payload):
        payload = [v & 0xFF for v in payload][:256]
        sqc = self.seq
        if self.setting_active:
            return
        self.seq = (self.seq + 1) & 0xFF
        body = bytes([0, 0, sqc, len(payload) & 0xFF] + payload)
        crc = crc16(body)
        self.frames.append(body + bytes([crc >> 8, crc & 0xFF]))

    def decode(self, groups) -> list[bytes]:
        for g in groups:
            self._decode([int(x) for x in g])
        out, self.frames = self.frames, []
        return out

    def _decode(self, b):
        """:166-272"""
        gtype = (b[1] >> 11) & 0x1F
        vb = bool(gtype & 1)
        if b[0] != self.pi:
            self.reset()
            self.pi = b[0]
            self._send([0x01, 0, 1, b[0] & 0xFF, b[0] >> 8])
        pty = (b[1] >> 5) & 0x1F
        if pty != self.pty:
            self.pty = pty
            self._send([0x07, 0, 1, pty])
        if gtype in (0x00, 0x01):
            self._type0(b)
        elif gtype in (0x02, 0x03):
            self._type1(b, vb)
        elif gtype in (0x04, 0x05):
            self._type2(b, vb)
        elif gtype == 0x06:
            self._type3a(b)
        elif gtype == 0x08:
            self._type4a(b)
        elif gtype == 0x14:
            self._type10a(b)
        elif gtype in (0x1C, 0x1D, 0x1E, 0x1F):
            pass
        elif self.oda[gtype] > 0:
            self._oda(b, self.oda[gtype])
        elif gtype == 0x10:
            self._send([0x30, 6, 0, b[1] & 0x1F, b[2] >> 8, b[2] & 0xFF, b[3] >> 8, b[3] & 0xFF])

    def _type0(self, b):
        """:309-408"""
        seg = b[1] & 3
        bit = {3: 1, 2: 2, 1: 4, 0: 8}[seg]
        if b[1] & 0x04:
            self.di |= bit
        elif self.di & bit:
            self.di ^= bit
        self.di_finished += 1
        ta_tp = (1 if b[1] & 0x10 else 0) | (2 if b[1] & 0x400 else 0)
        if self.ta_tp != ta_tp:
            self.ta_tp = ta_tp
            self._send([0x03, 0, 1, ta_tp])
        if self.di_finished >= 4 and self.di_prev != self.di:
            self.di_finished = 0
            self.di_prev = self.di
            self._send([0x04, 0, 1, self.di & 0xF])
        self.ms = 1 if b[1] & 0x08 else 0
        if self.ms_prev != self.ms:
            self.ms_prev = self.ms
            self._send([0x05, 0, 1, self.ms])
        self.ps_text[seg * 2] = b[3] >> 8
        self.ps_text[seg * 2 + 1] = b[3] & 0xFF
        self.ps_set |= 1 << seg
        if self.ps_set == 0x0F:
            if self.setting_active or self.ps_name[:8] != self.ps_text[:8]:
                self.names.append(bytes(self.ps_text[:8]).split(b"\0")[0].ljust(8, b"\0"))
                if self.accept_name and not self.setting_active:
                    self._send([0x02, 0, 1] + list(self.ps_text[:8]))
                    self.ps_name[:8] = self.ps_text[:8]
                self.ps_set = 0

    def _type1(self, b, vb):
        """:556-587"""
        if self.pin != b[3]:
            self.pin = b[3]
            self._send([0x06, 0, 1, b[3] >> 8, b[3] & 0xFF])
        if not vb:
            self._send([0x1A, 0, (b[2] >> 8) & 0x7F, b[2] & 0xFF])

    def _type2(self, b, vb):
        """:592-658"""
        seg = b[1] & 0x0F
        self.rtplus_ready = False
        if seg == 0 and self.rt_first and self.rt_count > 1:
            ready = True
            for i in range(self.rt_count):
                if not (self.rt_seg >> i) & 1:
                    ready = False
                    self.rt_seg = 0
                    self.rt_count = 0
                    break
            if ready:
                self._send([0x0A, 0, 1, 65, self.rt_ab] + list(self.rt[:64]))
                self.rtplus_ready = True
        ab = (b[1] >> 4) & 1
        if self.rt_ab != ab:
            self.rt = bytearray(b"\x20" * 66)
            self.rt_ab = ab
            self.rt_first = False
            self.rt_seg = 0
            self.rt_count = 0
        if not vb:
            self.rt[seg * 4:seg * 4 + 4] = bytes([b[2] >> 8, b[2] & 0xFF, b[3] >> 8, b[3] & 0xFF])
        else:
            self.rt[seg * 2:seg * 2 + 2] = bytes([b[3] >> 8, b[3] & 0xFF])
        self.rt_seg |= 1 << seg
        self.rt_count += 1
        if not self.rt_first and seg == 0:
            self.rt_first = True

    def _type3a(self, b):
        """:663-706"""
        self._send([0x40, b[1] & 0x1F, b[3] >> 8, b[3] & 0xFF, 0, b[2] >> 8, b[2] & 0xFF, 0])
        self.oda[b[1] & 0x1F] = b[3] if b[3] in (AID_RTPLUS, AID_TFC) else 0

    def _type4a(self, b):
        """:711-741 (C arithmetic: double MJD, (int) truncations, unsigned wrap)"""
        mjd = float(((b[1] & 3) << 15) | ((b[2] >> 1) & 0x7FFF))
        hours = ((b[2] & 1) << 4) | ((b[3] >> 12) & 0xF)
        minutes = (b[3] >> 6) & 0x3F
        offset = b[3] & 0x3F
        year = _c_int((mjd - 15078.2) / 365.25) & 0xFFFFFFFF
        month = _c_int((mjd - 14956.1 - _c_int(year * 365.25)) / 30.6001) & 0xFFFFFFFF
        day = _c_int(mjd - 14956 - _c_int(year * 365.25) - _c_int(month * 30.6001)) & 0xFFFFFFFF
        k = 1 if month in (14, 15) else 0
        year = (year + k + 1900) & 0xFFFFFFFF
        month = (month - (1 + k * 12)) & 0xFFFFFFFF
        self._send([0x0D, year % 100, month, day, hours, minutes, 0, 0, offset])

    def _type10a(self, b):
        """:821-853"""
        seg = b[1] & 1
        ab = (b[1] >> 4) & 1
        if self.ptyn_ab != ab:
            self.ptyn[:8] = b"\x20" * 8
            self.ptyn_ab = ab
            self.ptyn_set = 0
        self.ptyn[seg * 4:seg * 4 + 4] = bytes([b[2] >> 8, b[2] & 0xFF, b[3] >> 8, b[3] & 0xFF])
        self.ptyn_set |= 1 << seg
        if self.ptyn_set & 3:
            self._send([0x3A, 0, 1] + list(self.ptyn[:8]))

    def _oda(self, b, aid):
        """:922-960"""
        if aid == AID_RTPLUS:
            if self.rtplus_ready:
                self._send([0x46, 8, 0x4B, 0xD7, b[1] >> 8, b[1] & 0xFF, b[2] >> 8, b[2] & 0xFF, b[3] >> 8, b[3] & 0xFF])
                self.rtplus_ready = False
        elif aid == AID_TFC:
            self._send([0x46, 7, 0xCD, 0x46, b[1] & 0xFF, b[2] >> 8, b[2] & 0xFF, b[3] >> 8, b[3] & 0xFF])
