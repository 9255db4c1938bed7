"""Hash-only fixture for the long-run parity tests (BASELINE.json configs 2 / 3: 60 s of 1.2 MS/s stereo + RDS; the
10 s case is its first 153 blocks), generated from the UNMODIFIED reference (oracle/_ref/libradiofm_ref.so).

    python tests/golden/make_golden_long.py        (build container: needs `make -C oracle ref`)

Input: the 153-block (10.02 s) synthetic station of radiofm_b200.synth, stream 0, repeated to 1098 blocks (59.95 s):
block b of the run is block b mod 153 of the segment.  The seams (every 10.02 s) are phase jumps of carrier, pilot and
RDS clock -- the demodulator PLL, the pilot PLL (stereo flag drops and re-locks) and the RDS block sync all have to
re-acquire, which makes the run a harsher carried-state test than an unbroken capture.  The input itself is not stored
(20 MB): the test regenerates it and checks its SHA-256.

Stored (< 10 KB): audio SHA-256 after every 100 blocks and after block 153, audio floats and stereo flag per block,
every decoded group, a digest of the raw RDS bits.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402
from oracle import ref  # noqa: E402

SEG, NBLK = 153, 1098


def checkpoints():
    return sorted(set(list(range(100, NBLK + 1, 100)) + [SEG, NBLK]))


def main():
    fs, ds, blk = conftest.RATES["1.2M"]
    iq, _ = conftest.station("1.2M", SEG)
    d = ref.RefFmDecoder(fs, -0.15 * fs, downsample=ds)
    h = hashlib.sha256()
    cps, lens, stereo = {}, [], []
    for b in range(NBLK):
        k = b % SEG
        a, _ = d.process_staged(ref.u8_to_cf32(iq[k * blk:(k + 1) * blk]), want=())
        h.update(a.tobytes())
        lens.append(a.size)
        stereo.append(bool(d.status()["stereo"]))
        if b + 1 in checkpoints():
            cps[b + 1] = h.copy().hexdigest()
    groups, bits = d.take_groups(1 << 16), d.take_bits(1 << 22)
    np.savez_compressed(os.path.join(HERE, "long_1.2M.npz"), fs=fs, ds=ds, blk=blk, seg=SEG, nblk=NBLK,
                        iq_sha256=hashlib.sha256(iq.tobytes()).hexdigest(),
                        cp_blocks=np.array(sorted(cps)), cp_sha256=np.array([cps[k] for k in sorted(cps)]),
                        audio_len=np.array(lens, dtype=np.uint16), stereo=np.packbits(np.array(stereo)),
                        groups=groups, n_bits=bits.size, bits_sha256=hashlib.sha256(bits.tobytes()).hexdigest())
    print("blocks", NBLK, "groups", len(groups), "bits", bits.size, "stereo blocks", int(np.sum(stereo)),
          "first stereo", int(np.argmax(stereo)))


if __name__ == "__main__":
    main()
