"""Design study (CPU, NumPy): how often does an FP32 expanded-form search
(e = |t|^2 - 2 s.t, 2 FFMA per pair-eval) leave the float64 winner outside the best group of G
targets, given the guaranteed error bound?  Prints ambiguous-source fractions per group size."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import icp_oracle as orc

U = 2.0 ** -24


def f32(x):
    return np.asarray(x, dtype=np.float32)


def study(src64, tgt64, G, center=True):
    o = tgt64.mean(axis=0) if center else np.zeros(2)
    sc = f32(src64 - o); tc = f32(tgt64 - o)
    tt = f32(tc[:, 0].astype(np.float64) ** 2 + tc[:, 1].astype(np.float64) ** 2)
    a = (-2.0 * sc[:, 0]).astype(np.float64); b = (-2.0 * sc[:, 1]).astype(np.float64)
    inner = f32(b[:, None] * tc[None, :, 1].astype(np.float64) + tt[None, :].astype(np.float64))
    e = f32(a[:, None] * tc[None, :, 0].astype(np.float64) + inner.astype(np.float64))
    n, m = e.shape
    mp = (m + G - 1) // G * G
    ep = np.full((n, mp), np.inf, dtype=np.float32); ep[:, :m] = e
    gm = ep.reshape(n, mp // G, G).min(axis=2)
    order = np.sort(gm, axis=1)
    best, second = order[:, 0], (order[:, 1] if gm.shape[1] > 1 else np.full(n, np.inf))
    bg = gm.argmin(axis=1)
    Cs = np.abs(sc).max(axis=1).astype(np.float64); Ct = float(np.abs(tc).max())
    E = 8.0 * U * Ct * (Cs + Ct)                      # >= 6u Ct (Cs+Ct) proven bound
    rho = 2.83 * max(Ct, Cs.max()) * U
    ss = (sc.astype(np.float64) ** 2).sum(axis=1)
    dB = np.sqrt(np.maximum(best + ss, 0) + E)
    thr = best + 2 * E + 4 * dB * rho + 2 * rho * rho
    amb = second <= thr
    # truth check: the f64 winner must be in the best group whenever not ambiguous
    d64 = ((src64[:, None, :] - tgt64[None, :, :]) ** 2).sum(axis=2)
    win = d64.argmin(axis=1)
    bad = (~amb) & (win // G != bg)
    return amb.mean(), bad.sum(), n


def main():
    rows = []
    for G in (8, 16, 32):
        tot_amb = tot = bad = 0
        for p in range(0, 40):
            s, t, _, _ = orc.synth_room_pair(p)
            r = orc.icp_extended(s, t, 8, -1.0)
            src = s.astype(np.float64)
            # replay the oracle's src states: iteration states via cumulative application
            states = [src]
            cur = src
            for it in range(7):
                R, tt_ = orc.best_fit_transform(cur, t.astype(np.float64)[r.indices[it]])
                cur = (R @ cur.T).T + tt_
                states.append(cur)
            for st in states:
                a, b, n = study(st, t.astype(np.float64), G)
                tot_amb += a * n; tot += n; bad += b
        rows.append((G, tot_amb / tot, bad))
    print("synthetic rooms (360x360):", rows)
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scan_data_1_packed.npz"))
    off = z["offsets"]
    raw = np.stack([z["quality"].astype(float), z["angle64"] / 64.0, z["dist4"] / 4.0], 1)
    scans = [orc.polar_to_cartesian(raw[off[i]:off[i + 1]])[:, :2] for i in range(0, 600)]
    rows = []
    for G in (8, 16, 32):
        tot_amb = tot = bad = 0
        for p in range(2, 599, 7):
            a, b, n = study(scans[p + 1], scans[p], G)
            tot_amb += a * n; tot += n; bad += b
        rows.append((G, tot_amb / tot, bad))
    print("Scan_data_1 first-iteration:", rows)


if __name__ == "__main__":
    main()
