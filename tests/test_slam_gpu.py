"""The offline SLAM loop (icp_slam-yolo_b200/slam.py, the composition of
duc/ICP_LIDAR/slam_offline.py:318-455 out of the device primitives) against the same composition
of the CPU oracles, on the first frames of the bundled recording.  Every step is pinned on its
own; this test checks that they compose: same accept / reject decisions, same poses, same map,
same occupancy grid."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import icp_oracle as orc            # noqa: E402
from oracle import occupancy_oracle as occ      # noqa: E402,F401
from oracle.slam_oracle import OracleSlam        # noqa: E402


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import icp_slam_yolo_b200 as m
    m.lib()
    return m


def test_offline_slam_loop_matches_the_oracle_composition(pkg, cart_scans, tmp_path):
    from icp_slam_yolo_b200.slam import OfflineSlam, SlamConfig
    scans = [np.c_[s, np.zeros(len(s))] for s in cart_scans[2:132]]     # scans 3 .. 132 of Scan_data_1
    dev, ora = OfflineSlam(SlamConfig()), OracleSlam(SlamConfig())
    dev.first_scan(scans[0])
    ora.first(scans[0])
    accepted = 0
    for k, s in enumerate(scans[1:]):
        r = dev.step(s)
        o = ora.step(s)
        assert (r is None) == (o is None), f"frame {k}"
        if r is None:
            continue
        assert r.accepted == o[0], f"frame {k}: rmse {r.rmse} vs {o[1]}"
        assert abs(r.rmse - o[1]) < 1e-6 or (np.isinf(r.rmse) and np.isinf(o[1])), f"frame {k}"
        assert np.allclose(r.pose, ora.pose, atol=1e-6), f"frame {k}"
        got = dev.global_map.cpu().numpy()
        assert got.shape == ora.map.shape and np.allclose(got, ora.map, atol=1e-6), f"frame {k}"
        accepted += int(r.accepted)
    assert accepted >= 100 and len(ora.map) > 1000                       # the loop tracks; the map was re-sampled
    probs = dev.grid.probs_numpy()
    assert np.array_equal(probs.view(np.uint32), ora.occ.view(np.uint32))
    assert np.array_equal(dev.grid.image_numpy(), ora.img)
    assert probs.max() >= 0.65 and probs.min() < 0.2
    pcd, png = str(tmp_path / "global_map_offline.pcd"), str(tmp_path / "realtime_occupancy_map.png")
    dev.save(pcd, png)
    assert np.array_equal(pkg.map_io.read_png(png), ora.img)
    want = orc.voxel_down_sample_2d(ora.map, 25.0).astype(np.float32)
    back = pkg.map_io.read_pcd(pcd)
    assert back.shape == (len(want), 3) and np.allclose(back[:, :2], want, atol=1e-3)
