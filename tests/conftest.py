import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "reference_icp_golden.npz"))


@pytest.fixture(scope="session")
def raw_scans():
    """The 1,831 polar scans of Scan_data_1, bit-exact (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "scan_data_1_packed.npz"))
    off = z["offsets"]
    rows = np.stack([z["quality"].astype(np.float64), z["angle64"].astype(np.float64) / 64.0,
                     z["dist4"].astype(np.float64) / 4.0], axis=1)
    return [rows[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.fixture(scope="session")
def cart_scans(raw_scans):
    """Cartesian (N,2) float64 scans through the oracle's process.py:38-52 restatement."""
    from oracle import icp_oracle as orc
    return [np.ascontiguousarray(orc.polar_to_cartesian(r)[:, :2]) for r in raw_scans]


@pytest.fixture(scope="session")
def oracle_pairs(cart_scans):
    """Extended-oracle results for all 1,830 consecutive pairs (k+1 -> k), 30 its, tol 1e-5."""
    from oracle import icp_oracle as orc
    return [orc.icp_extended(cart_scans[p + 1], cart_scans[p], 30, 1e-5) for p in range(len(cart_scans) - 1)]


def _unpack(name):
    z = np.load(os.path.join(GOLDEN, name))
    off = z["offsets"]
    rows = np.stack([z["quality"].astype(np.float64), z["angle64"].astype(np.float64) / 64.0,
                     z["dist4"].astype(np.float64) / 4.0], axis=1)
    return [rows[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.fixture(scope="session")
def cart_scans3():
    """The second bundled recording, scan_data_3/ (2,043 scans), Cartesian after the canonical filter
    (tests/golden/make_golden_scan3.py: lossless repack)."""
    from oracle import icp_oracle as orc
    return [np.ascontiguousarray(orc.polar_to_cartesian(r)[:, :2]) for r in _unpack("scan_data_3_packed.npz")]


@pytest.fixture(scope="session")
def golden3():
    return np.load(os.path.join(GOLDEN, "reference_icp_golden_scan3.npz"))


@pytest.fixture(scope="session")
def oracle_pairs3(cart_scans3):
    from oracle import icp_oracle as orc
    return [orc.icp_extended(cart_scans3[p + 1], cart_scans3[p], 30, 1e-5) for p in range(len(cart_scans3) - 1)]
