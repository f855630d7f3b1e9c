/*
 * icp_oracle.c -- plain-C restatement of the reference's 2D ICP loop.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ and by bench.py's CPU
 * legs as the checker for sizes where the NumPy/SciPy oracle is too slow (scan-to-map,
 * large batches).  Never linked into or called from the product.
 *
 * Follows /root/reference/labels_segmentation/icp.py:
 *   :37-38  KDTree(B).query(src)   -> nn_bruteforce(): float64 argmin, lowest index on ties
 *                                     (identical to the KD-tree on all 1.86 M real queries,
 *                                     SURVEY.md 7.1-1c; checked again in tests/test_oracle.py)
 *   :10-16  centroids, centred H    -> fit_closed_form()
 *   :17-25  SVD -> R, reflection fix, t -> closed form theta = atan2(H01-H10, H00+H11)
 *                                     (equal to the SVD route to ~1e-16 rad)
 *   :45     src = R src + t
 *   :48-51  lagged mean distance, |prev - mean| < tolerance, prev_error starts at 0
 * Parity: pinned through oracle/icp_oracle.py (tests compare the two on real scans).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void nn_bruteforce(const double* src, int n, const double* tgt, int m, int32_t* idx, double* d2) {
  for (int i = 0; i < n; ++i) {
    const double sx = src[2 * i], sy = src[2 * i + 1];
    double best = INFINITY;
    int bj = -1;
    for (int j = 0; j < m; ++j) {
      const double dx = sx - tgt[2 * j], dy = sy - tgt[2 * j + 1];
      const double d = dx * dx + dy * dy;
      if (d < best) { best = d; bj = j; }
    }
    idx[i] = bj;
    d2[i] = best;
  }
}

/* best_fit_transform (icp.py:5-26) on matched rows P -> Q, closed form. */
static void fit_closed_form(const double* P, const double* Q, int n, double* cs, double* sn,
                            double* tx, double* ty) {
  double cpx = 0, cpy = 0, cqx = 0, cqy = 0;
  for (int i = 0; i < n; ++i) { cpx += P[2*i]; cpy += P[2*i+1]; cqx += Q[2*i]; cqy += Q[2*i+1]; }
  cpx /= n; cpy /= n; cqx /= n; cqy /= n;
  double h00 = 0, h01 = 0, h10 = 0, h11 = 0;
  for (int i = 0; i < n; ++i) {
    const double ax = P[2*i] - cpx, ay = P[2*i+1] - cpy, bx = Q[2*i] - cqx, by = Q[2*i+1] - cqy;
    h00 += ax * bx; h01 += ax * by; h10 += ay * bx; h11 += ay * by;
  }
  const double num = h01 - h10, den = h00 + h11, hyp = hypot(num, den);
  *cs = hyp > 0 ? den / hyp : 1.0;
  *sn = hyp > 0 ? num / hyp : 0.0;
  *tx = cqx - (*cs * cpx - *sn * cpy);
  *ty = cqy - (*sn * cpx + *cs * cpy);
}

/*
 * icp (icp.py:28-53) with the bookkeeping of oracle.icp_oracle.icp_extended.
 * pose_total / pose_last: [R00 R01 R10 R11 tx ty].  max_corr_dist <= 0: no gate.
 * idx_history (nullable): [max_iterations][n].  src_out (nullable): [n][2].
 * Returns the iteration count.
 */
int icp_oracle(const double* A, int n, const double* B, int m, int max_iterations, double tolerance,
               const double* init_pose, double max_corr_dist, double* pose_total, double* pose_last,
               double* error, double* rmse, int32_t* inliers, int32_t* idx_history, double* src_out) {
  double R00 = 1, R01 = 0, R10 = 0, R11 = 1, T0 = 0, T1 = 0;
  double* src = (double*)malloc(sizeof(double) * 2 * (n > 0 ? n : 1));
  double* P = (double*)malloc(sizeof(double) * 2 * (n > 0 ? n : 1));
  double* Q = (double*)malloc(sizeof(double) * 2 * (n > 0 ? n : 1));
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
  double* d2 = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
  if (init_pose) { R00 = init_pose[0]; R01 = init_pose[1]; R10 = init_pose[2]; R11 = init_pose[3]; T0 = init_pose[4]; T1 = init_pose[5]; }
  for (int i = 0; i < n; ++i) {
    const double x = A[2*i], y = A[2*i+1];
    src[2*i] = init_pose ? R00 * x + R01 * y + T0 : x;
    src[2*i+1] = init_pose ? R10 * x + R11 * y + T1 : y;
  }
  pose_last[0] = 1; pose_last[1] = 0; pose_last[2] = 0; pose_last[3] = 1; pose_last[4] = 0; pose_last[5] = 0;
  *error = INFINITY; *rmse = INFINITY; *inliers = 0;
  int iters = 0;
  double prev = 0.0;
  for (int it = 0; it < max_iterations && n > 0 && m > 0; ++it) {
    nn_bruteforce(src, n, B, m, idx, d2);
    int k = 0;
    double sd = 0, sd2 = 0;
    for (int i = 0; i < n; ++i) {
      const double dist = sqrt(d2[i]);
      if (max_corr_dist > 0 && !(dist < max_corr_dist)) continue;
      P[2*k] = src[2*i]; P[2*k+1] = src[2*i+1];
      Q[2*k] = B[2*idx[i]]; Q[2*k+1] = B[2*idx[i]+1];
      sd += dist; sd2 += d2[i]; ++k;
    }
    if (k == 0) { *error = INFINITY; *rmse = INFINITY; *inliers = 0; break; }
    if (idx_history) for (int i = 0; i < n; ++i) idx_history[(size_t)it * n + i] = idx[i];
    double cs, sn, tx, ty;
    fit_closed_form(P, Q, k, &cs, &sn, &tx, &ty);
    for (int i = 0; i < n; ++i) {
      const double x = src[2*i], y = src[2*i+1];
      src[2*i] = cs * x - sn * y + tx;
      src[2*i+1] = sn * x + cs * y + ty;
    }
    const double n00 = cs * R00 - sn * R10, n01 = cs * R01 - sn * R11;
    const double n10 = sn * R00 + cs * R10, n11 = sn * R01 + cs * R11;
    const double nt0 = cs * T0 - sn * T1 + tx, nt1 = sn * T0 + cs * T1 + ty;
    R00 = n00; R01 = n01; R10 = n10; R11 = n11; T0 = nt0; T1 = nt1;
    pose_last[0] = cs; pose_last[1] = -sn; pose_last[2] = sn; pose_last[3] = cs; pose_last[4] = tx; pose_last[5] = ty;
    const double mean = sd / k;
    *error = mean; *rmse = sqrt(sd2 / k); *inliers = k;
    iters = it + 1;
    if (fabs(prev - mean) < tolerance) break;
    prev = mean;
  }
  pose_total[0] = R00; pose_total[1] = R01; pose_total[2] = R10; pose_total[3] = R11; pose_total[4] = T0; pose_total[5] = T1;
  if (src_out) for (int i = 0; i < 2 * n; ++i) src_out[i] = src[i];
  free(src); free(P); free(Q); free(idx); free(d2);
  return iters;
}
