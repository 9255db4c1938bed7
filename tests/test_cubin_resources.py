"""Static checks on the compiled sm_100a objects (no GPU needed): the hot kernels of the chain keep their state in
registers (no stack frame, no local memory), and the register counts stay inside the occupancy the launch geometry
assumes.  Skipped when the objects or cuobjdump are not there (build() makes them)."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "pvr.rtl.radiofm_b200", "csrc", "build")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

# kernel-name fragment -> (max registers per thread, note)
HOT = {
    "k_front_tmaILi11E": (64, "persistent, 128 threads, 3 outputs per thread, five CTAs per SM (TMA-staged window)"),
    "k_front_tmaILi5E": (64, ""),
    "k_front_tiledILb1ELi11E": (64, "128 threads, 4 outputs per thread: >= 8 CTAs per SM (rows a tensor map cannot describe)"),
    "k_front_tiledILb1ELi4E": (64, ""),
    "k_front_tiledILb1ELi5E": (64, ""),
    "k_bb_lanesILb0ELb0ELb0E": (128, "64 threads; pilot state + 13 double constants in registers"),
    "k_bb_lanesILb0ELb1ELb0E": (128, "immediate-barrier form (SM partition: up to 12 CTAs per SM)"),
    "k_demod_spec": (96, "32-thread CTAs, ~17 per SM"),
    "k_resample_tiledILi6E": (64, "192 threads, two CTAs per SM"),
    "k_resample_tiledILi4E": (64, ""),
    "k_rds_front3ILi15ELi23ELi43E": (64, "128 threads, eight CTAs per SM"),
    "k_rds_front3ILi15ELi19ELi35E": (64, ""),
    "k_rotfir_lanesILi0E": (64, ""),
    "k_rotfir_lanesILi1E": (64, ""),
    "k_rotfir_lanesILi2E": (64, ""),
    "k_rds_front": (64, ""),
    "k_rds_pll": (96, ""),
    "k_rds_slice": (96, ""),
    "k_audio_tail": (128, ""),
    "k_dc_chain_uniform": (64, "256 threads, two CTAs per SM (shared memory)"),
    "k_dc_oscE": (64, ""),
}


def _usage():
    objs = sorted(glob.glob(os.path.join(BUILD, "*.o")))
    if not objs or not os.path.exists(CUOBJDUMP):
        pytest.skip("no compiled objects / cuobjdump (run __graft_entry__.build())")
    out = {}
    for o in objs:
        r = subprocess.run([CUOBJDUMP, "--dump-resource-usage", o], capture_output=True, text=True)
        if r.returncode != 0:
            continue  # host-only object
        assert "sm_100a" in r.stdout, o
        name = None
        for line in r.stdout.splitlines():
            m = re.match(r"\s*Function\s+(\S+):", line)
            if m:
                name = m.group(1)
                continue
            m = re.search(r"REG:(\d+)\s+STACK:(\d+)\s+SHARED:(\d+)\s+LOCAL:(\d+)", line)
            if m and name:
                out[name] = tuple(int(v) for v in m.groups())
    return out


def test_hot_kernels_have_no_stack_or_local_memory():
    usage = _usage()
    for frag, (max_regs, _) in HOT.items():
        hits = {n: u for n, u in usage.items() if frag in n}
        assert hits, f"kernel {frag} not found in the compiled objects"
        for n, (regs, stack, _shared, local) in hits.items():
            assert stack == 0 and local == 0, (n, stack, local)
            assert regs <= max_regs, (n, regs, max_regs)


def test_every_kernel_is_sm_100a_and_fits_shared_memory():
    usage = _usage()
    assert len(usage) >= 40
    for n, (regs, _stack, shared, local) in usage.items():
        assert regs <= 128 and local == 0, (n, regs, local)
        assert shared <= 48 * 1024, (n, shared)  # static part; the dynamic part is opted in per launch (EnsureDynSmem)


def test_product_library_reads_no_environment_knobs():
    """The measurement knobs (RFM_DEBUG_* / RFM_LANES_* / RFM_RES_* / RFM_ROTFIR_*) exist only in the experiments
    build (RFM_EXPERIMENTS=1 sh build.sh -> libradiofm_b200_exp.so); the product library contains none of the names,
    and its lanes kernel has 5 named barriers (immediate ids), not 16."""
    lib = os.path.join(ROOT, "pvr.rtl.radiofm_b200", "libradiofm_b200.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    blob = open(lib, "rb").read()
    for frag in (b"RFM_DEBUG_", b"RFM_LANES_", b"RFM_RES_", b"RFM_ROTFIR_", b"FAKE_SINCOS"):
        assert frag not in blob, frag
    assert b"k_bb_lanesILb1E" not in blob  # the fake-sincos timing kernel is not even instantiated


def test_tma_is_really_used():
    """north_star: the front end's window is staged by the TMA unit -- the SASS of its kernel holds UTMALDG
    (cp.async.bulk.tensor) and the mbarrier wait (SYNCS), and none of the LDG-based staging of the fallback kernel."""
    lib = os.path.join(ROOT, "pvr.rtl.radiofm_b200", "libradiofm_b200.so")
    if not os.path.exists(lib) or not os.path.exists(CUOBJDUMP):
        pytest.skip("library / cuobjdump not available")
    r = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True)
    assert r.returncode == 0
    cur, ops = None, {}
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "k_front_tma" in cur:
            for op in ("UTMALDG", "SYNCS", "LDG.E.128"):
                if op in line:
                    ops.setdefault(cur, set()).add(op)
    assert len(ops) == 6, sorted(ops)   # ds = 1, 5, 11, exact and fused
    for k, v in ops.items():
        assert "UTMALDG" in v and "SYNCS" in v and "LDG.E.128" not in v, (k, v)


def test_packed_pairs_are_really_used_and_never_contracted():
    """The half-band chains (wideband front end, RDS front) do their taps on packed (re, im) pairs: their SASS holds
    FFMA2 and no scalar FMUL/FADD tap loop.  The exact product / exact sum are fma(a, b, -0) and fma(p, 1, c) with the
    two constants passed as kernel arguments -- ptxas contracts the literal forms into ONE FFMA2 per tap (not the
    reference's arithmetic): every such kernel must therefore hold an EVEN number of FFMA2, two per tap."""
    lib = os.path.join(ROOT, "pvr.rtl.radiofm_b200", "libradiofm_b200.so")
    if not os.path.exists(lib) or not os.path.exists(CUOBJDUMP):
        pytest.skip("library / cuobjdump not available")
    r = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True)
    assert r.returncode == 0
    cur, n2, nscalar = None, {}, {}
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and ("k_dc_chain_uniform" in cur or "k_rds_front3" in cur):
            if re.search(r"\bFFMA2\b", line):
                n2[cur] = n2.get(cur, 0) + 1
            elif re.search(r"\b(FMUL|FADD)\b", line):
                nscalar[cur] = nscalar.get(cur, 0) + 1
    assert len(n2) >= 5, sorted(n2)     # k_dc_chain_uniform<0..2, 51> and both k_rds_front3 plans
    for k, v in n2.items():
        # taps per output = (L + 1) / 2 even taps + the centre tap; outputs per thread 5 / 3 / 1 by stage
        if "k_dc_chain_uniform" in k:
            per_out = 2 * 27                          # HB51; the compiler may clone a stage body: whole outputs only
            assert v >= 9 * per_out and v % per_out == 0, (k, v)
        else:
            want = 5 * 9 + 3 * 13 + 23 if "ILi15ELi23ELi43" in k else 5 * 9 + 3 * 11 + 19
            assert v == 2 * want, (k, v, want)        # two FFMA2 per tap: nothing was contracted
        assert nscalar.get(k, 0) < v, (k, nscalar.get(k), v)   # what is left is the mixer in front of the taps
