#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs
  scan_data_1_packed.npz   lossless repack of Scan_data_1/ (1,831 scans).  The
                           recorder quantises quality to integers, angle to
                           1/64 deg and distance to 1/4 mm (read_lidar.py:72),
                           so (u8, u16, u16) round-trips every float64 bit;
                           this script asserts that before writing.
  reference_icp_golden.npz outputs of the reference's own icp()/best_fit_transform()
                           (icp.py:5-53, imported unmodified): the import-time
                           circle demo (icp.py:55-67) and icp(A,B,30,1e-5) on
                           every consecutive Scan_data_1 pair (k+1 -> k), full
                           ``src`` arrays for a spot list, summary for all.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import icp_oracle as orc          # noqa: E402
from oracle import ref_loader                  # noqa: E402

SPOT_PAIRS = [1, 3, 50, 100, 185, 251, 350, 500, 673, 800, 1007, 1063, 1075, 1300, 1600, 1830]


def pack_scans(directory, first=1, last=1831):
    q, a, d, off = [], [], [], [0]
    for k in range(first, last + 1):
        raw = np.load(orc.scan_path(directory, k))
        assert raw.dtype == np.float64 and raw.ndim == 2 and raw.shape[1] == 3, (k, raw.shape)
        qq = raw[:, 0].astype(np.uint8)
        aa = np.round(raw[:, 1] * 64.0).astype(np.uint16)
        dd = np.round(raw[:, 2] * 4.0).astype(np.uint16)
        back = np.stack([qq.astype(np.float64), aa.astype(np.float64) / 64.0,
                         dd.astype(np.float64) / 4.0], axis=1)
        assert np.array_equal(back.view(np.uint64), raw.view(np.uint64)), f"lossy repack at scan {k}"
        q.append(qq); a.append(aa); d.append(dd); off.append(off[-1] + len(raw))
    return dict(quality=np.concatenate(q), angle64=np.concatenate(a), dist4=np.concatenate(d),
                offsets=np.asarray(off, dtype=np.int32), first_index=np.int32(first))


def main():
    assert ref_loader.reference_available(), "needs /root/reference"
    ref = ref_loader.load_reference_icp()
    sdir = ref_loader.scan_dir("Scan_data_1")

    packed = pack_scans(sdir)
    np.savez_compressed(os.path.join(HERE, "scan_data_1_packed.npz"), **packed)

    out = {}
    # ---- circle demo (icp.py:55-67), values produced at import time
    out["demo_A"] = ref.A
    out["demo_B"] = ref.B
    out["demo_A_aligned"] = ref.A_aligned
    out["demo_R_est"] = ref.R_est
    out["demo_t_est"] = ref.t_est

    # ---- Scan_data_1 consecutive pairs through the reference icp()
    scans = []
    for k in range(1, 1832):
        pts = orc.polar_to_cartesian_loop(np.load(orc.scan_path(sdir, k)))
        scans.append(pts[:, :2].copy())
    n = len(scans) - 1
    R_last = np.zeros((n, 2, 2)); t_last = np.zeros((n, 2))
    theta_tot = np.zeros(n); t_tot = np.zeros((n, 2)); src_crc = np.zeros(n, dtype=np.uint32)
    n_src = np.zeros(n, dtype=np.int32); n_tgt = np.zeros(n, dtype=np.int32)
    for p in range(n):            # pair p: scan p+2 -> scan p+1 (1-based file numbers)
        A, B = scans[p + 1], scans[p]
        src, R, t = ref.icp(A, B, 30, 1e-5)
        R_last[p], t_last[p] = R, t
        Rc, tc = ref.best_fit_transform(A, src)      # cumulative pose implied by src (quirk Q1)
        theta_tot[p] = np.arctan2(Rc[1, 0], Rc[0, 0]); t_tot[p] = tc
        src_crc[p] = zlib.crc32(np.ascontiguousarray(src).tobytes())
        n_src[p], n_tgt[p] = len(A), len(B)
        if (p + 1) in SPOT_PAIRS:
            out[f"spot_src_{p + 1}"] = src
    out.update(pair_R_last=R_last, pair_t_last=t_last, pair_theta_tot=theta_tot, pair_t_tot=t_tot,
               pair_src_crc32=src_crc, pair_n_src=n_src, pair_n_tgt=n_tgt,
               spot_pairs=np.asarray(SPOT_PAIRS, dtype=np.int32))
    np.savez_compressed(os.path.join(HERE, "reference_icp_golden.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
