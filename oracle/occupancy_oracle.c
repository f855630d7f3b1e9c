/*
 * occupancy_oracle.c -- plain-C restatement of the reference's occupancy-grid update.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for long replays where the
 * pure-Python restatement (oracle/occupancy_oracle.py) is too slow.  Never linked into or
 * called from the product.
 *
 * Follows /root/reference/duc/ICP_LIDAR/process.py:
 *   :86-112   bresenham_line (float error term dx/2.0; here doubled integers, same decisions)
 *   :114-177  update_occupancy_map: window of +-area cells around the robot, per point a ray
 *             from the robot cell; free cells *= p_free_dec until a cell >= 0.65 stops the ray,
 *             the end cell += p_occ_inc (clamped to 1); window re-rendered to grey levels.
 * Arithmetic: float32 products/sums with float32 constants (NumPy 2 scalar rules), float64 for
 * the cell coordinates, int() = truncation toward zero.
 * Parity: pinned through oracle/occupancy_oracle.py (tests compare the two) and directly
 * against tests/golden/reference_occupancy_golden.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static long long py_int(double v) { return (long long)v; }   /* int(): toward zero */

/* len(range(*slice(a, b).indices(n))) for a >= 0 */
static int slice_len(long long a, long long b, int n) {
  if (b < 0) { b += n; if (b < 0) b = 0; }
  if (b > n) b = n;
  if (a > n) a = n;
  return b > a ? (int)(b - a) : 0;
}

void occ_update(float* occ, uint8_t* image, int h, int w, const double* pts, int n,
                double robot_x, double robot_y, double cx, double cy, double res,
                float p_occ_inc, float p_free_dec, int area) {
  if (n == 0) return;
  const float thr_up = 0.65f;
  const long long rxp = py_int(cx + robot_x / res);
  const long long ryp = py_int(cy - robot_y / res);
  const long long x1 = rxp - area > 0 ? rxp - area : 0;
  const long long y1 = ryp - area > 0 ? ryp - area : 0;
  const long long x2 = rxp + area < w ? rxp + area : w;
  const long long y2 = ryp + area < h ? ryp + area : h;
  const int width = slice_len(x1, x2, w), height = slice_len(y1, y2, h);
  const long long rx = rxp - x1, ry = ryp - y1;
  for (int p = 0; p < n; ++p) {
    const long long px = py_int(cx + pts[2 * p] / res - (double)x1);
    const long long py = py_int(cy - pts[2 * p + 1] / res - (double)y1);
    if (!(0 <= px && px < width && 0 <= py && py < height)) continue;
    const long long dx = llabs(px - rx), dy = llabs(py - ry);
    const long long sx = rx > px ? -1 : 1, sy = ry > py ? -1 : 1;
    long long x = rx, y = ry;
    int stopped = 0;
    if (dx > dy) {
      long long e2 = dx;                       /* 2 * err */
      while (x != px) {
        if (0 <= x && x < width && 0 <= y && y < height) {
          float* c = &occ[(y1 + y) * w + (x1 + x)];
          if (*c >= thr_up) { stopped = 1; break; }
          const float v = *c * p_free_dec;
          *c = v > 0.0f ? v : 0.0f;            /* max(0.0, v): Python returns 0.0 unless v > 0.0 */
        }
        e2 -= 2 * dy;
        if (e2 < 0) { y += sy; e2 += 2 * dx; }
        x += sx;
      }
    } else {
      long long e2 = dy;
      while (y != py) {
        if (0 <= x && x < width && 0 <= y && y < height) {
          float* c = &occ[(y1 + y) * w + (x1 + x)];
          if (*c >= thr_up) { stopped = 1; break; }
          const float v = *c * p_free_dec;
          *c = v > 0.0f ? v : 0.0f;
        }
        e2 -= 2 * dx;
        if (e2 < 0) { x += sx; e2 += 2 * dy; }
        y += sy;
      }
    }
    if (!stopped) {                            /* the end cell, already known to be in the window */
      float* c = &occ[(y1 + py) * w + (x1 + px)];
      const float v = *c + p_occ_inc;
      *c = v < 1.0f ? v : 1.0f;                /* min(1.0, v) */
    }
  }
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const float g = (1.0f - occ[(y1 + y) * w + (x1 + x)]) * 255.0f;
      const uint8_t u = (uint8_t)g;
      uint8_t* q = &image[((y1 + y) * (long long)w + (x1 + x)) * 3];
      q[0] = u; q[1] = u; q[2] = u;
    }
}
