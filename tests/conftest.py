"""Shared test plumbing.

* registers the ``gpu`` marker (tests that need a real B200; everything else runs on CPU);
* ``rfm``      -- the product's ctypes binding (pvr.rtl.radiofm_b200/__init__.py), loaded by path because
                  the package directory name is not a valid dotted module name;
* ``port``/``ref`` -- the checker: oracle/port.py (plain-C restatement) and oracle/ref.py (the
                  unmodified reference compiled by oracle/Makefile; present wherever oracle/_ref/ was built);
* cached synthetic stations.
"""
from __future__ import annotations

import functools
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_package():
    name = "radiofm_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "pvr.rtl.radiofm_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def rfm():
    return load_package()


@pytest.fixture(scope="session")
def synth():
    load_package()
    import importlib
    return importlib.import_module("radiofm_b200.synth")


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libradiofm_ref.so not built (needs /root/reference; see oracle/Makefile)")
    r.lib()
    return r


# (fs, downsample, block length) of the BASELINE.json configurations, SURVEY.md section 8 table
RATES = {
    "1.0M": (1.0e6, 4, 65536),
    "1.2M": (1.2e6, 5, 65520),
    "2.4M": (2.4e6, 11, 65472),
    "390k": (390625.0, 1, 16000),
}


@functools.lru_cache(maxsize=16)
def station(rate: str, n_blocks: int, stream_id: int = 0, stereo: bool = True, rds: bool = True, mono: bool = False):
    """Synthetic u8 IQ [n_blocks * block, 2] and the transmitted RDS groups."""
    load_package()
    import importlib
    synth = importlib.import_module("radiofm_b200.synth")
    fs, ds, blk = RATES[rate]
    n = n_blocks * blk
    iq, groups = synth.make_station_u8(fs, n, stream_id=stream_id, stereo=stereo, rds=rds,
                                       mono_tone=(1000.0, 0.5) if mono else None)
    return iq, groups


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
