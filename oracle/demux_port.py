"""CPU restatement of the caller of the hot path: cRadioReceiver::DemuxRead with its IQ block queue
(RadioReceiver.cpp:420-542, audio level :528-529 / :584-598), composed from the oracle's own cFmDecoder restatement
(oracle/port.py) and group decoder restatement (oracle/uecp_port.py).

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  Parity pinned: tests/test_ref_addon.py runs this restatement
against the UNMODIFIED RadioReceiver.cpp / RTL_SDR_Source.cpp compiled in place (oracle/ref_addon_harness.cpp) --
packet order, ids, pts / duration, audio bits, UECP bytes, audio level.
"""
from __future__ import annotations

import numpy as np

from . import port, uecp_port

STREAM_TIME_BASE = 1000000.0
STREAMCHANGE = -11


class OracleDemux:
    def __init__(self, fs_if, tuning_offset, fs_pcm=48000.0, bw_pcm=15000.0, downsample=1, usver=False):
        self.dec = port.OracleFmDecoder(fs_if, tuning_offset, fs_pcm, bw_pcm, downsample, usver)
        self.groups = uecp_port.OracleGroupDecoder()
        self.fs_pcm = fs_pcm
        self.queue: list[np.ndarray] = []
        self.stream_change = True          # :345
        self.pts_next = STREAM_TIME_BASE   # :347
        self.audio_level = np.float32(0.0)
        self.uecp = bytearray()

    def write_u8(self, iq_u8):
        self.queue.append(np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1, 2))

    def read(self):
        """One DemuxRead call; None when the queue is empty (end marked)."""
        if self.stream_change:
            self.stream_change = False
            return STREAMCHANGE, 0.0, 0.0, None
        if self.uecp:
            data, self.uecp = bytes(self.uecp), bytearray()
            return 2, self.pts_next, 0.0, data
        if not self.queue:
            return None
        audio = self.dec.process_u8(self.queue.pop(0))
        for f in self.groups.decode(self.dec.take_groups()):        # DecodeRDS runs inside ProcessStream
            if len(self.uecp) <= 16384:                              # AddUECPDataFrame, :389-390
                self.uecp += uecp_port.stuff_frame(f)
        vsum = np.float32(0.0)
        vsumsq = np.float32(0.0)
        sq = (audio * audio).astype(np.float32)
        for v in sq:                                                 # float accumulation in sample order, :586-594
            vsumsq = np.float32(vsumsq + v)
        rms = float(np.sqrt(np.float64(np.float32(vsumsq / np.float32(audio.size)))))
        self.audio_level = np.float32(0.95 * float(self.audio_level) + 0.05 * rms)
        duration = float(audio.size) * STREAM_TIME_BASE / 2 / self.fs_pcm
        pkt = (1, self.pts_next, duration, audio)
        self.pts_next = self.pts_next + duration
        return pkt
