#!/usr/bin/env python
"""Second real recording: scan_data_3/ (2,043 scans scan_{0..2042}.npy) of the reference tree.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_scan3.py

Outputs
  scan_data_3_packed.npz      lossless (u8, u16, u16) repack, asserted bit for bit like Scan_data_1
  reference_icp_golden_scan3.npz   the UNMODIFIED reference icp(A, B, 30, 1e-5) (labels_segmentation/
                              icp.py:28-53) on every consecutive pair (k+1 -> k) after the canonical
                              polar filter (process.py:38-52): last increment R, t, cumulative pose
                              implied by src, CRC32 of the src bytes, point counts
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import icp_oracle as orc          # noqa: E402
from oracle import ref_loader                  # noqa: E402


def main():
    assert os.path.isdir(ref_loader.scan_dir("scan_data_3")), "needs /root/reference"
    ref = ref_loader.load_reference_icp()
    sdir = ref_loader.scan_dir("scan_data_3")
    q, a, d, off, scans = [], [], [], [0], []
    for k in range(0, 2043):
        raw = np.load(os.path.join(sdir, f"scan_{k}.npy"))
        assert raw.dtype == np.float64 and raw.ndim == 2 and raw.shape[1] == 3, (k, raw.shape)
        qq = raw[:, 0].astype(np.uint8)
        aa = np.round(raw[:, 1] * 64.0).astype(np.uint16)
        dd = np.round(raw[:, 2] * 4.0).astype(np.uint16)
        back = np.stack([qq.astype(np.float64), aa.astype(np.float64) / 64.0, dd.astype(np.float64) / 4.0], axis=1)
        assert np.array_equal(back.view(np.uint64), raw.view(np.uint64)), f"lossy repack at scan {k}"
        q.append(qq); a.append(aa); d.append(dd); off.append(off[-1] + len(raw))
        scans.append(orc.polar_to_cartesian_loop(raw)[:, :2].copy())
    np.savez_compressed(os.path.join(HERE, "scan_data_3_packed.npz"), quality=np.concatenate(q),
                        angle64=np.concatenate(a), dist4=np.concatenate(d),
                        offsets=np.asarray(off, dtype=np.int32), first_index=np.int32(0))
    n = len(scans) - 1
    R_last = np.zeros((n, 2, 2)); t_last = np.zeros((n, 2))
    theta_tot = np.zeros(n); t_tot = np.zeros((n, 2)); src_crc = np.zeros(n, dtype=np.uint32)
    n_src = np.zeros(n, dtype=np.int32); n_tgt = np.zeros(n, dtype=np.int32)
    for p in range(n):
        A, B = scans[p + 1], scans[p]
        n_src[p], n_tgt[p] = len(A), len(B)
        if len(A) == 0 or len(B) == 0:                # the reference raises inside SciPy on empty input
            continue
        src, R, t = ref.icp(A, B, 30, 1e-5)
        R_last[p], t_last[p] = R, t
        Rc, tc = ref.best_fit_transform(A, src)
        theta_tot[p] = np.arctan2(Rc[1, 0], Rc[0, 0]); t_tot[p] = tc
        src_crc[p] = zlib.crc32(np.ascontiguousarray(src).tobytes())
    np.savez_compressed(os.path.join(HERE, "reference_icp_golden_scan3.npz"), pair_R_last=R_last, pair_t_last=t_last,
                        pair_theta_tot=theta_tot, pair_t_tot=t_tot, pair_src_crc32=src_crc, pair_n_src=n_src,
                        pair_n_tgt=n_tgt)
    print("pairs", n, "points per scan", int(n_src.min()), "..", int(n_src.max()), "empty pairs", int(np.sum((n_src == 0) | (n_tgt == 0))))


if __name__ == "__main__":
    main()
