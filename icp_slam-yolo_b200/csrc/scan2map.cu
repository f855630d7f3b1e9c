// scan2map.cu -- scan-to-map ICP against a large map sharded across GPUs (sm_100a).
//
// Same reference loop as b200icp.cu (labels_segmentation/icp.py:28-53) with a target set of
// millions of points: the map shard streams through shared memory in 1,024-point chunks
// (TMA bulk copies, cp.async.bulk + mbarrier, double buffered) while every lane sweeps its
// source points over the chunk in FP32, and the exactness scheme of DESIGN.md 4.1 is applied
// with *per-chunk* origins so the FP32 error band stays ~1e-5 mm even for maps that span
// tens of metres:
//   s2m_prepare_kernel   chunk centroids, chunk-centred float32 SoA copy of the shard
//   s2m_bound / s2m_cull per tile of 512 scan points, the ordered list of chunks that can matter
//   s2m_sweep_kernel     per (source, 8 listed chunks): upper bound of the best distance,
//                        lower bounds of the best and of the runner-up group, best group
//   s2m_resolve_kernel   merge segments, FP32 in-group argmin, exact float64 distance of the
//                        winner; sources whose runner-up may beat the winner go to a list
//   s2m_exact_*_kernel   float64 brute force over the whole shard for the listed sources
//   s2m_update_kernel    after the records of all ranks are gathered: per point the global
//                        winner (smaller distance, then lower global index), the centred sums,
//                        closed-form pose, apply, convergence -- identical on every rank
// Replaces: KDTree(B).query(src) (icp.py:37-38), best_fit_transform (icp.py:5-26), the apply /
// convergence steps (icp.py:45-51) for the scan-to-local-map call shape of
// duc/ICP_LIDAR/mainn.py:297-318.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "b200icp.h"

void b200icp_set_error_str(const char* msg);   // defined in b200icp.cu

namespace {

constexpr int kChunk = 1024;          // targets per staged chunk
constexpr int kGroup = 8;             // targets per tracked group
constexpr int kGroupsPerChunk = kChunk / kGroup;   // 128
constexpr int kSweepThreads = 128;
constexpr int kSweepS = 4;            // source points per lane
constexpr int kSrcPerCta = kSweepThreads * kSweepS;   // 512
constexpr unsigned kFull = 0xffffffffu;

struct Partial {        // per (segment, source)
  float ub;             // upper bound of the distance to the best candidate
  float lb1;            // lower bound of the distance to the best group's minimum
  float lb2;            // lower bound over every other group seen
  uint32_t where;       // chunk * 128 + group of the best candidate (shard-local)
};

__device__ __forceinline__ double2 load_point(const void* base, int dtype, int64_t i) {
  if (dtype == B200ICP_F64) return __ldg(reinterpret_cast<const double2*>(base) + i);
  const float2 v = __ldg(reinterpret_cast<const float2*>(base) + i);
  return make_double2((double)v.x, (double)v.y);
}

// float64 squared distance in NumPy's operation order (no contraction)
__device__ __forceinline__ double dist2_f64(double sx, double sy, double2 t) {
  const double dx = __dsub_rn(sx, t.x), dy = __dsub_rn(sy, t.y);
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

__device__ __forceinline__ void merge_partial(float& g_ub, float& g_lb1, float& g_lb2,
                                              uint32_t& g_where, float ub, float lb1, float lb2,
                                              uint32_t where) {
  if (ub < g_ub) {
    g_lb2 = fminf(g_lb2, fminf(g_lb1, lb2));
    g_ub = ub; g_lb1 = lb1; g_where = where;
  } else {
    g_lb2 = fminf(g_lb2, lb1);
  }
}

// ---- mbarrier / TMA bulk copy (PTX; shared::cluster == shared::cta for a 1-CTA cluster) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------
// prepare: chunk origins + chunk-centred float32 SoA
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) s2m_prepare_kernel(const void* points, int dtype, int64_t m,
                                                          float* cx, float* cy, double* origin,
                                                          float* radius) {
  __shared__ double sred[8][2];
  __shared__ float fred[8];
  const int64_t c = blockIdx.x;
  const int64_t j0 = c * kChunk;
  const int cnt = (int)min((int64_t)kChunk, m - j0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double sx = 0.0, sy = 0.0;
  for (int j = tid; j < cnt; j += blockDim.x) {
    const double2 q = load_point(points, dtype, j0 + j);
    sx += q.x; sy += q.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(kFull, sx, o);
    sy += __shfl_xor_sync(kFull, sy, o);
  }
  if (lane == 0) { sred[warp][0] = sx; sred[warp][1] = sy; }
  __syncthreads();
  double tx = 0.0, ty = 0.0;
  for (int w = 0; w < 8; ++w) { tx += sred[w][0]; ty += sred[w][1]; }
  const double ox = tx / (double)cnt, oy = ty / (double)cnt;
  float amax = 0.f;
  for (int j = tid; j < kChunk; j += blockDim.x) {
    float fx = CUDART_INF_F, fy = CUDART_INF_F;      // sentinels: infinitely far
    if (j < cnt) {
      const double2 q = load_point(points, dtype, j0 + j);
      fx = (float)(q.x - ox); fy = (float)(q.y - oy);
      amax = fmaxf(amax, fmaxf(fabsf(fx), fabsf(fy)));
    }
    cx[j0 + j] = fx; cy[j0 + j] = fy;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(kFull, amax, o));
  if (lane == 0) fred[warp] = amax;
  __syncthreads();
  if (tid == 0) {
    float r = 0.f;
    for (int w = 0; w < 8; ++w) r = fmaxf(r, fred[w]);
    origin[2 * c] = ox; origin[2 * c + 1] = oy;
    radius[c] = r;
  }
}

// ------------------------------------------------------------------------------------------
// culling: which chunks can matter for a tile of 512 consecutive scan points?
//   bound: ub_i = min_c (|s_i - o_c| + r_c) is an upper bound of point i's NN distance (some map
//          point of chunk c lies within r_c of its origin o_c); one cheap pass over the chunk
//          table (n x n_chunks centre distances, ~0.1 % of the full sweep).
//   cull : chunk c is kept for tile T iff dist(o_c, box(T)) - r_c <= max_{i in T} ub_i (+ slack).
//          Every point of a culled chunk is farther from every point of the tile than that
//          point's nearest neighbour, so it can be neither the winner nor a tie: results are
//          identical to the full sweep.  Kept chunks are written as an ordered list per tile.
// ------------------------------------------------------------------------------------------
constexpr int kItemsPerTile = 256;    // sweep CTAs per tile of 512 scan points (grid.x)
// listed chunks per sweep CTA: the tile's kept chunks are spread evenly over kItemsPerTile CTAs
__host__ __device__ inline int seg_chunks_for(int kept) {
  const int s = (kept + kItemsPerTile - 1) / kItemsPerTile;
  return s < 1 ? 1 : s;
}
constexpr float kSqrt2Up = 1.4142137f;

constexpr int kBoundChunks = 256;     // chunk origins staged per CTA of the bound kernel

// grid (source blocks, chunk parts): every CTA stages kBoundChunks chunk origins in shared memory
// and folds them into ub[] with an atomic min on the (non-negative) float bit patterns.
// ub[] must be pre-set to a huge value (0x7f7f7f7f bytes).  Rounded up by the reader.
__global__ void __launch_bounds__(128) s2m_bound_kernel(const double* __restrict__ origin,
                                                        const float* __restrict__ radius, int n_chunks,
                                                        const double* __restrict__ src64, int n,
                                                        float* __restrict__ ub,
                                                        const b200icp_s2m_state* __restrict__ state) {
  __shared__ double sox[kBoundChunks], soy[kBoundChunks];
  __shared__ float srad[kBoundChunks];
  if (state->done) return;
  const int c0 = blockIdx.y * kBoundChunks, cnt = min(kBoundChunks, n_chunks - c0);
  for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
    sox[k] = origin[2 * (c0 + k)]; soy[k] = origin[2 * (c0 + k) + 1];
    srad[k] = radius[c0 + k] * kSqrt2Up;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sx = src64[2 * i], sy = src64[2 * i + 1];
  float best = CUDART_INF_F;
#pragma unroll 4
  for (int k = 0; k < cnt; ++k) {
    const float dx = (float)(sx - sox[k]), dy = (float)(sy - soy[k]);
    best = fminf(best, sqrtf(fmaf(dx, dx, dy * dy)) + srad[k]);
  }
  atomicMin(reinterpret_cast<unsigned*>(ub) + i, __float_as_uint(best));
}

__global__ void __launch_bounds__(256) s2m_cull_kernel(const double* __restrict__ origin,
                                                       const float* __restrict__ radius, int n_chunks,
                                                       const double* __restrict__ src64, int n,
                                                       const float* __restrict__ ub,
                                                       int32_t* __restrict__ tile_count,
                                                       int32_t* __restrict__ tile_list,
                                                       const b200icp_s2m_state* __restrict__ state) {
  __shared__ double sbox[8][4];
  __shared__ float sreach[8];
  __shared__ int wcount[8];
  __shared__ int base;
  if (state->done) return;
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double x0 = CUDART_INF, x1 = -CUDART_INF, y0 = CUDART_INF, y1 = -CUDART_INF;
  float reach = 0.f;
  for (int k = tid; k < kSrcPerCta; k += blockDim.x) {
    const int i = tile * kSrcPerCta + k;
    if (i < n) {
      const double sx = src64[2 * i], sy = src64[2 * i + 1];
      x0 = fmin(x0, sx); x1 = fmax(x1, sx); y0 = fmin(y0, sy); y1 = fmax(y1, sy);
      reach = fmaxf(reach, ub[i] * 1.000002f + 1e-6f);     // ub holds the un-rounded minimum
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = fmin(x0, __shfl_xor_sync(kFull, x0, o)); x1 = fmax(x1, __shfl_xor_sync(kFull, x1, o));
    y0 = fmin(y0, __shfl_xor_sync(kFull, y0, o)); y1 = fmax(y1, __shfl_xor_sync(kFull, y1, o));
    reach = fmaxf(reach, __shfl_xor_sync(kFull, reach, o));
  }
  if (lane == 0) { sbox[warp][0] = x0; sbox[warp][1] = x1; sbox[warp][2] = y0; sbox[warp][3] = y1; sreach[warp] = reach; }
  if (tid == 0) base = 0;
  __syncthreads();
  for (int w = 0; w < 8; ++w) {
    x0 = fmin(x0, sbox[w][0]); x1 = fmax(x1, sbox[w][1]);
    y0 = fmin(y0, sbox[w][2]); y1 = fmax(y1, sbox[w][3]);
    reach = fmaxf(reach, sreach[w]);
  }
  const double lim = (double)reach * 1.000001 + 1e-3;        // slack: cull arithmetic is float64
  int32_t* list = tile_list + (int64_t)tile * n_chunks;
  for (int c0 = 0; c0 < n_chunks; c0 += blockDim.x) {
    const int c = c0 + tid;
    bool keep = false;
    if (c < n_chunks) {
      const double2 o = __ldg(reinterpret_cast<const double2*>(origin) + c);
      const double dx = fmax(fmax(x0 - o.x, o.x - x1), 0.0), dy = fmax(fmax(y0 - o.y, o.y - y1), 0.0);
      keep = sqrt(dx * dx + dy * dy) - (double)(radius[c] * kSqrt2Up) <= lim;
    }
    const unsigned ballot = __ballot_sync(kFull, keep);
    if (lane == 0) wcount[warp] = __popc(ballot);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += wcount[w];
    if (keep) list[off + __popc(ballot & ((1u << lane) - 1u))] = c;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += wcount[w];
      base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) tile_count[tile] = base;
}

// ------------------------------------------------------------------------------------------
// sweep: FP32 direct-difference search of a segment of chunks for 512 source points
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSweepThreads) s2m_sweep_kernel(
    const float* __restrict__ cx, const float* __restrict__ cy, const double* __restrict__ origin,
    const float* __restrict__ radius, int n_chunks, const int32_t* __restrict__ tile_count,
    const int32_t* __restrict__ tile_list, const double* __restrict__ src64, int n,
    Partial* __restrict__ partials, const b200icp_s2m_state* __restrict__ state) {
  __shared__ __align__(128) float buf[2][2][kChunk];     // [stage][x|y][point]
  __shared__ __align__(8) uint64_t bars[2];
  if (state->done) return;
  const int tid = threadIdx.x;
  const int item = blockIdx.x, tile = blockIdx.y;
  const int kept = tile_count[tile];
  const int seg = seg_chunks_for(kept);
  const int c0 = item * seg, c1 = min(kept, c0 + seg);                 // positions in the tile's list
  if (c0 >= c1) return;
  const int32_t* __restrict__ list = tile_list + (int64_t)tile * n_chunks;

  double sx[kSweepS], sy[kSweepS];
  float g_ub[kSweepS], g_lb1[kSweepS], g_lb2[kSweepS];
  uint32_t g_where[kSweepS];
#pragma unroll
  for (int k = 0; k < kSweepS; ++k) {
    const int i = tile * kSrcPerCta + k * kSweepThreads + tid;
    const int ii = min(i, n - 1);
    sx[k] = src64[2 * ii]; sy[k] = src64[2 * ii + 1];
    g_ub[k] = CUDART_INF_F; g_lb1[k] = CUDART_INF_F; g_lb2[k] = CUDART_INF_F; g_where[k] = 0;
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr uint32_t kHalf = kChunk * sizeof(float);
  if (tid == 0) {
    const int64_t cc = list[c0];
    mbar_expect_tx(&bars[0], 2 * kHalf);
    bulk_g2s(&buf[0][0][0], cx + cc * kChunk, kHalf, &bars[0]);
    bulk_g2s(&buf[0][1][0], cy + cc * kChunk, kHalf, &bars[0]);
  }
  for (int pos = c0; pos < c1; ++pos) {
    const int c = list[pos];                  // the chunk swept in this trip
    const int st = (pos - c0) & 1;
    const uint32_t phase = ((pos - c0) >> 1) & 1;
    if (tid == 0 && pos + 1 < c1) {          // stage st^1 was released by the barrier below
      const int64_t cn = list[pos + 1];
      mbar_expect_tx(&bars[st ^ 1], 2 * kHalf);
      bulk_g2s(&buf[st ^ 1][0][0], cx + cn * kChunk, kHalf, &bars[st ^ 1]);
      bulk_g2s(&buf[st ^ 1][1][0], cy + cn * kChunk, kHalf, &bars[st ^ 1]);
    }
    const double ox = origin[2 * c], oy = origin[2 * c + 1];
    const float rc = radius[c];
    float nfx[kSweepS], nfy[kSweepS], best[kSweepS], second[kSweepS];
    int grp[kSweepS];
#pragma unroll
    for (int k = 0; k < kSweepS; ++k) {
      float vx = -(float)(sx[k] - ox), vy = -(float)(sy[k] - oy);
      asm volatile("" : "+f"(vx), "+f"(vy));
      nfx[k] = vx; nfy[k] = vy;
      best[k] = CUDART_INF_F; second[k] = CUDART_INF_F; grp[k] = 0;
    }
    mbar_wait(&bars[st], phase);
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(&buf[st][0][0]);
    const float4* __restrict__ y4 = reinterpret_cast<const float4*>(&buf[st][1][0]);
#pragma unroll 1
    for (int g = 0; g < kGroupsPerChunk; ++g) {
      const float4 xa = x4[2 * g], xb = x4[2 * g + 1];
      const float4 ya = y4[2 * g], yb = y4[2 * g + 1];
#pragma unroll
      for (int k = 0; k < kSweepS; ++k) {
        const float2 ax = make_float2(nfx[k], nfx[k]), ay = make_float2(nfy[k], nfy[k]);
        const float2 u0 = __fadd2_rn(ax, make_float2(xa.x, xa.y)), v0 = __fadd2_rn(ay, make_float2(ya.x, ya.y));
        const float2 u1 = __fadd2_rn(ax, make_float2(xa.z, xa.w)), v1 = __fadd2_rn(ay, make_float2(ya.z, ya.w));
        const float2 u2 = __fadd2_rn(ax, make_float2(xb.x, xb.y)), v2 = __fadd2_rn(ay, make_float2(yb.x, yb.y));
        const float2 u3 = __fadd2_rn(ax, make_float2(xb.z, xb.w)), v3 = __fadd2_rn(ay, make_float2(yb.z, yb.w));
        const float2 d0 = __ffma2_rn(v0, v0, __fmul2_rn(u0, u0)), d1 = __ffma2_rn(v1, v1, __fmul2_rn(u1, u1));
        const float2 d2 = __ffma2_rn(v2, v2, __fmul2_rn(u2, u2)), d3 = __ffma2_rn(v3, v3, __fmul2_rn(u3, u3));
        const float mm = fminf(fminf(fminf(d0.x, d0.y), fminf(d1.x, d1.y)),
                               fminf(fminf(d2.x, d2.y), fminf(d3.x, d3.y)));
        const float old = best[k];
        second[k] = fminf(second[k], fmaxf(old, mm));
        grp[k] = (mm < old) ? g : grp[k];
        best[k] = fminf(old, mm);
      }
    }
    // distance bounds of this chunk's candidates (per-chunk origin):
    // |sqrt(d32) - true| <= sqrt(d32) * 2^-22 + 2.9 * (cs + rc) * 2^-23
#pragma unroll
    for (int k = 0; k < kSweepS; ++k) {
      const float cs = fmaxf(fabsf(nfx[k]), fabsf(nfy[k]));
      const float mu = (cs + rc) * 3.8146973e-7f;                 // 3.2 * 2^-23
      const float sb = sqrtf(best[k]), ss = sqrtf(second[k]);
      const float ub = fmaf(sb, 1.0000005f, mu);
      const float lb1 = fmaf(sb, 0.9999995f, -mu);
      const float lb2 = fmaf(ss, 0.9999995f, -mu);
      merge_partial(g_ub[k], g_lb1[k], g_lb2[k], g_where[k], ub, lb1, lb2,
                    (uint32_t)c * kGroupsPerChunk + (uint32_t)grp[k]);
    }
    __syncthreads();     // every thread is done with stage st before it is refilled
  }
#pragma unroll
  for (int k = 0; k < kSweepS; ++k) {
    const int i = tile * kSrcPerCta + k * kSweepThreads + tid;
    if (i < n) {
      Partial p;
      p.ub = g_ub[k]; p.lb1 = g_lb1[k]; p.lb2 = g_lb2[k]; p.where = g_where[k];
      partials[(int64_t)item * n + i] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------
// resolve: merge segments, decide inside the best group, exact float64 for the winner
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) s2m_resolve_kernel(
    const void* points, int dtype, int64_t m, int64_t global_offset, const float* __restrict__ cx,
    const float* __restrict__ cy, const double* __restrict__ origin, const float* __restrict__ radius,
    const double* __restrict__ src64, int n, const Partial* __restrict__ partials,
    const int32_t* __restrict__ tile_count, b200icp_s2m_record* __restrict__ records, int32_t* __restrict__ amb_list,
    int32_t* __restrict__ amb_count, const b200icp_s2m_state* __restrict__ state) {
  if (state->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g_ub = CUDART_INF_F, g_lb1 = CUDART_INF_F, g_lb2 = CUDART_INF_F;
  uint32_t where = 0;
  const int kept = tile_count[i / kSrcPerCta];
  const int seg_chunks = seg_chunks_for(kept);
  const int n_items = (kept + seg_chunks - 1) / seg_chunks;
  if (n_items == 0) {            // every chunk of this shard is out of reach: another rank holds the NN
    b200icp_s2m_record none;
    none.d2 = CUDART_INF; none.gidx = 0x7fffffffffffffffLL; none.bx = 0.0; none.by = 0.0;
    records[i] = none;
    return;
  }
  int s = 0;
  for (; s + 4 <= n_items; s += 4) {          // four independent loads in flight per trip
    const Partial p0 = partials[(int64_t)s * n + i], p1 = partials[(int64_t)(s + 1) * n + i];
    const Partial p2 = partials[(int64_t)(s + 2) * n + i], p3 = partials[(int64_t)(s + 3) * n + i];
    merge_partial(g_ub, g_lb1, g_lb2, where, p0.ub, p0.lb1, p0.lb2, p0.where);
    merge_partial(g_ub, g_lb1, g_lb2, where, p1.ub, p1.lb1, p1.lb2, p1.where);
    merge_partial(g_ub, g_lb1, g_lb2, where, p2.ub, p2.lb1, p2.lb2, p2.where);
    merge_partial(g_ub, g_lb1, g_lb2, where, p3.ub, p3.lb1, p3.lb2, p3.where);
  }
  for (; s < n_items; ++s) {
    const Partial p = partials[(int64_t)s * n + i];
    merge_partial(g_ub, g_lb1, g_lb2, where, p.ub, p.lb1, p.lb2, p.where);
  }
  bool ambiguous = g_lb2 <= g_ub;
  const int64_t c = where / kGroupsPerChunk;
  const int64_t j0 = (int64_t)where * kGroup;                  // shard-local index of the group
  const double sx = src64[2 * i], sy = src64[2 * i + 1];
  const float fx = (float)(sx - origin[2 * c]), fy = (float)(sy - origin[2 * c + 1]);
  unsigned best = 0x7f800000u, second = 0x7f800000u;
#pragma unroll
  for (int u = 0; u < kGroup; ++u) {
    const float dx = fx - cx[j0 + u], dy = fy - cy[j0 + u];
    const float d = fmaf(dy, dy, dx * dx);                    // +inf for sentinel slots
    const unsigned key = (__float_as_uint(d) & ~7u) | (unsigned)u;
    second = min(second, max(best, key));
    best = min(best, key);
  }
  {
    const float bd = __uint_as_float(best & ~7u), sd = __uint_as_float(second & ~7u);
    const float cs = fmaxf(fabsf(fx), fabsf(fy));
    const float guard = (cs + radius[c]) * 4.76837158e-7f;     // 2^-21
    const float r = sqrtf(bd) * 1.000004f + guard;
    ambiguous |= sd <= r * r * 1.000001f;
  }
  const int64_t j = j0 + (best & 7u);
  const double2 b = load_point(points, dtype, min(j, m - 1));
  b200icp_s2m_record rec;
  rec.d2 = dist2_f64(sx, sy, b);
  rec.gidx = global_offset + j;
  rec.bx = b.x; rec.by = b.y;
  records[i] = rec;
  if (ambiguous) amb_list[atomicAdd(amb_count, 1)] = i;
}

// ------------------------------------------------------------------------------------------
// exact: float64 brute force over the shard for the listed sources (lowest index on ties).
// The shard is cut into kExactParts slices; CTA (p, y) scans slice p for the listed sources
// e = y, y + gridDim.y, ... with four independent loads in flight per thread, so even a single
// listed source is spread over kExactParts SMs instead of streaming the shard through one.
// ------------------------------------------------------------------------------------------
constexpr int kExactParts = 64;
constexpr int kExactRows = 8;

struct ExactPartial {
  double d2;
  long long j;
};

__global__ void __launch_bounds__(256) s2m_exact_scan_kernel(
    const void* points, int dtype, int64_t m, const double* __restrict__ src64,
    const int32_t* __restrict__ amb_list, const int32_t* __restrict__ amb_count,
    ExactPartial* __restrict__ exact_partials, const b200icp_s2m_state* __restrict__ state) {
  __shared__ double sd[8];
  __shared__ long long sj[8];
  if (state->done) return;
  const int count = *amb_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t slice = (m + kExactParts - 1) / kExactParts;
  const int64_t j_begin = (int64_t)blockIdx.x * slice, j_end = min(m, j_begin + slice);
  for (int e = blockIdx.y; e < count; e += gridDim.y) {
    const int i = amb_list[e];
    const double sx = src64[2 * i], sy = src64[2 * i + 1];
    double bd = CUDART_INF;
    long long bj = 0x7fffffffffffffffLL;
    int64_t j = j_begin + tid;
    for (; j + 3 * 256 < j_end; j += 4 * 256) {          // ascending j per thread
      const double2 q0 = load_point(points, dtype, j), q1 = load_point(points, dtype, j + 256);
      const double2 q2 = load_point(points, dtype, j + 512), q3 = load_point(points, dtype, j + 768);
      const double d0 = dist2_f64(sx, sy, q0), d1 = dist2_f64(sx, sy, q1);
      const double d2 = dist2_f64(sx, sy, q2), d3 = dist2_f64(sx, sy, q3);
      if (d0 < bd) { bd = d0; bj = j; }
      if (d1 < bd) { bd = d1; bj = j + 256; }
      if (d2 < bd) { bd = d2; bj = j + 512; }
      if (d3 < bd) { bd = d3; bj = j + 768; }
    }
    for (; j < j_end; j += 256) {
      const double d = dist2_f64(sx, sy, load_point(points, dtype, j));
      if (d < bd) { bd = d; bj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(kFull, bd, o);
      const long long oj = __shfl_xor_sync(kFull, bj, o);
      if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
    }
    if (lane == 0) { sd[warp] = bd; sj[warp] = bj; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (sd[w] < bd || (sd[w] == bd && sj[w] < bj)) { bd = sd[w]; bj = sj[w]; }
      ExactPartial r;
      r.d2 = bd; r.j = bj;
      exact_partials[(int64_t)e * kExactParts + blockIdx.x] = r;
    }
    __syncthreads();
  }
}

// one warp per listed source: lexicographic (distance, index) minimum over the slices
__global__ void __launch_bounds__(256) s2m_exact_reduce_kernel(
    const void* points, int dtype, int64_t global_offset, const int32_t* __restrict__ amb_list,
    const int32_t* __restrict__ amb_count, const ExactPartial* __restrict__ exact_partials,
    b200icp_s2m_record* __restrict__ records, const b200icp_s2m_state* __restrict__ state) {
  if (state->done) return;
  const int count = *amb_count;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < count; e += warps) {
    double bd = CUDART_INF;
    long long bj = 0x7fffffffffffffffLL;
    for (int p = lane; p < kExactParts; p += 32) {
      const ExactPartial r = exact_partials[(int64_t)e * kExactParts + p];
      if (r.d2 < bd || (r.d2 == bd && r.j < bj)) { bd = r.d2; bj = r.j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(kFull, bd, o);
      const long long oj = __shfl_xor_sync(kFull, bj, o);
      if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
    }
    if (lane == 0) {
      const double2 b = load_point(points, dtype, bj);
      b200icp_s2m_record rec;
      rec.d2 = bd; rec.gidx = global_offset + bj; rec.bx = b.x; rec.by = b.y;
      records[amb_list[e]] = rec;
    }
  }
}

// ------------------------------------------------------------------------------------------
// init / update
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) s2m_init_kernel(const void* scan, int dtype, int n,
                                                       const double* init_pose, double* src64,
                                                       b200icp_s2m_state* state) {
  double R00 = 1, R01 = 0, R10 = 0, R11 = 1, T0 = 0, T1 = 0;
  if (init_pose) { R00 = init_pose[0]; R01 = init_pose[1]; R10 = init_pose[2]; R11 = init_pose[3]; T0 = init_pose[4]; T1 = init_pose[5]; }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double2 q = load_point(scan, dtype, i);
    src64[2 * i] = init_pose ? R00 * q.x + R01 * q.y + T0 : q.x;
    src64[2 * i + 1] = init_pose ? R10 * q.x + R11 * q.y + T1 : q.y;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state->pose_total[0] = R00; state->pose_total[1] = R01; state->pose_total[2] = R10;
    state->pose_total[3] = R11; state->pose_total[4] = T0; state->pose_total[5] = T1;
    state->pose_last[0] = 1; state->pose_last[1] = 0; state->pose_last[2] = 0;
    state->pose_last[3] = 1; state->pose_last[4] = 0; state->pose_last[5] = 0;
    state->error = CUDART_INF; state->mean_d2 = CUDART_INF; state->prev_error = 0.0;   // icp.py:33
    state->iterations = 0; state->inliers = 0; state->done = (n <= 0) ? 1 : 0; state->reserved = 0;
  }
}

// One CTA of 1024 threads; every rank runs it on the same gathered records and therefore
// reaches bit-identical poses and the same `done` decision without another collective.
__global__ void __launch_bounds__(1024) s2m_update_kernel(
    const b200icp_s2m_record* __restrict__ records_all, int n_ranks, double* src64, int n,
    int max_iterations, double tolerance, double max_corr_dist, int32_t* idx_out,
    b200icp_s2m_state* state) {
  __shared__ double red[32][12];
  __shared__ double tot[12];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (state->done) return;
  const bool use_gate = max_corr_dist > 0.0 && isfinite(max_corr_dist);
  const double ox = src64[0], oy = src64[1];             // any common origin keeps the sums small
  double r[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = tid; i < n; i += blockDim.x) {
    b200icp_s2m_record w = records_all[i];
    for (int g = 1; g < n_ranks; ++g) {
      const b200icp_s2m_record c = records_all[(int64_t)g * n + i];
      if (c.d2 < w.d2 || (c.d2 == w.d2 && c.gidx < w.gidx)) w = c;
    }
    if (idx_out) idx_out[i] = (int32_t)w.gidx;
    const double dist = sqrt(w.d2);
    if (!use_gate || dist < max_corr_dist) {
      const double ax = src64[2 * i] - ox, ay = src64[2 * i + 1] - oy;
      const double qx = w.bx - ox, qy = w.by - oy;
      r[0] += ax; r[1] += ay; r[2] += qx; r[3] += qy;
      r[4] = fma(ax, qx, r[4]); r[5] = fma(ax, qy, r[5]);
      r[6] = fma(ay, qx, r[6]); r[7] = fma(ay, qy, r[7]);
      r[8] += dist; r[9] += w.d2; r[10] += 1.0;
    }
  }
#pragma unroll
  for (int q = 0; q < 11; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r[q] += __shfl_xor_sync(kFull, r[q], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 11; ++q) red[warp][q] = r[q];
  }
  __syncthreads();
  if (tid < 11) {
    double acc = 0.0;
    for (int w = 0; w < 32; ++w) acc += red[w][tid];
    tot[tid] = acc;
  }
  __syncthreads();
  const double cnt = tot[10];
  if (cnt < 0.5) {               // every correspondence gated out: stop, search not counted
    if (tid == 0) { state->error = CUDART_INF; state->mean_d2 = CUDART_INF; state->inliers = 0; state->done = 1; }
    return;
  }
  const double inv = 1.0 / cnt;
  const double max_ = tot[0] * inv, may_ = tot[1] * inv, mbx = tot[2] * inv, mby = tot[3] * inv;
  const double mean_error = tot[8] * inv;                          // icp.py:48
  const double h00 = fma(-tot[0], mbx, tot[4]), h01 = fma(-tot[0], mby, tot[5]);
  const double h10 = fma(-tot[1], mbx, tot[6]), h11 = fma(-tot[1], mby, tot[7]);
  const double num = h01 - h10, den = h00 + h11;
  const double h2 = fma(num, num, den * den);
  double cs = 1.0, sn = 0.0;
  if (h2 > 0.0) { const double rh = rsqrt(h2); cs = den * rh; sn = num * rh; }
  const double cax = ox + max_, cay = oy + may_;
  const double tx = (ox + mbx) - (cs * cax - sn * cay);            // icp.py:25
  const double ty = (oy + mby) - (sn * cax + cs * cay);
  __syncthreads();                                                 // all reads of src64[0..1] done
  for (int i = tid; i < n; i += blockDim.x) {                      // icp.py:45
    const double x = src64[2 * i], y = src64[2 * i + 1];
    src64[2 * i] = cs * x - sn * y + tx;
    src64[2 * i + 1] = sn * x + cs * y + ty;
  }
  if (tid == 0) {
    const double R00 = state->pose_total[0], R01 = state->pose_total[1];
    const double R10 = state->pose_total[2], R11 = state->pose_total[3];
    const double T0 = state->pose_total[4], T1 = state->pose_total[5];
    state->pose_total[0] = cs * R00 - sn * R10; state->pose_total[1] = cs * R01 - sn * R11;
    state->pose_total[2] = sn * R00 + cs * R10; state->pose_total[3] = sn * R01 + cs * R11;
    state->pose_total[4] = cs * T0 - sn * T1 + tx; state->pose_total[5] = sn * T0 + cs * T1 + ty;
    state->pose_last[0] = cs; state->pose_last[1] = -sn; state->pose_last[2] = sn;
    state->pose_last[3] = cs; state->pose_last[4] = tx; state->pose_last[5] = ty;
    state->error = mean_error; state->mean_d2 = tot[9] * inv; state->inliers = (int)(cnt + 0.5);
    const int it = state->iterations + 1;
    state->iterations = it;
    const bool converged = fabs(state->prev_error - mean_error) < tolerance;   // icp.py:49-50
    state->prev_error = mean_error;                                             // icp.py:51
    if (converged || it >= max_iterations) state->done = 1;
  }
}

// ------------------------------------------------------------------------------------------
// peer exchange: the all-gather of the records fused into a store-to-every-peer kernel.
// Every rank owns one peer-visible buffer  [2 slots][world][n] records + [2][world] int64 flags.
// publish: each thread stores its 32-byte record into slot `slot`, row `rank` of EVERY rank's
// buffer (NVLink peer stores), then the last CTA raises flags[slot][rank] = seq on every rank
// (system-scope release).  wait: spins (system-scope acquire) until all `world` flags of the
// slot reached seq.  Slots alternate per iteration, which is enough: a peer can publish
// iteration k+2 only after it has seen this rank's k+1, i.e. after this rank finished reading k.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) s2m_publish_kernel(const b200icp_s2m_record* __restrict__ local,
                                                          int n, void* const* __restrict__ peers, int world,
                                                          int rank, int slot, long long seq,
                                                          unsigned int* __restrict__ block_counter,
                                                          const b200icp_s2m_state* __restrict__ state) {
  if (state->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double4 v = reinterpret_cast<const double4*>(local)[i];
    for (int r = 0; r < world; ++r) {
      double4* dst = reinterpret_cast<double4*>(peers[r]) + ((int64_t)slot * world + rank) * n + i;
      *dst = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(block_counter, 1u);
    if (done == gridDim.x - 1) {                       // last CTA: every record of this rank is out
      *block_counter = 0;
      __threadfence_system();
      for (int r = 0; r < world; ++r) {
        long long* flags = reinterpret_cast<long long*>(reinterpret_cast<double4*>(peers[r]) + (int64_t)2 * world * n);
        asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(flags + slot * world + rank), "l"(seq) : "memory");
      }
    }
  }
}

__global__ void s2m_wait_kernel(const void* mine, int n, int world, int slot, long long seq,
                                b200icp_s2m_state* state) {
  if (state->done) return;
  const long long* flags = reinterpret_cast<const long long*>(reinterpret_cast<const double4*>(mine) + (int64_t)2 * world * n);
  if ((int)threadIdx.x < world) {
    const long long* f = flags + slot * world + threadIdx.x;
    const long long t0 = clock64();
    long long v = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > 4000000000LL) {            // ~2 s: a peer is gone; fail instead of hanging
        state->done = 2;
        break;
      }
    }
  }
}

int fail(const char* msg, int code) {
  b200icp_set_error_str(msg);
  return code;
}

int cuda_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return B200ICP_OK;
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  b200icp_set_error_str(buf);
  return B200ICP_ERR_CUDA;
}

struct Workspace {        // byte offsets inside the caller's workspace
  int64_t amb_list, tile_count, tile_list, partials, exact, total;
  int n_items, tiles;
};

Workspace layout_workspace(int n, int64_t m) {
  auto up = [](int64_t b) { return (b + 255) / 256 * 256; };
  Workspace w;
  const int64_t n_chunks = (m + kChunk - 1) / kChunk;
  w.tiles = (n + kSrcPerCta - 1) / kSrcPerCta;
  w.n_items = (int)(n_chunks < kItemsPerTile ? n_chunks : kItemsPerTile);
  int64_t off = 256;                                         // [0]: ambiguous-source counter
  w.amb_list = off;   off += up((int64_t)n * 4);
  w.tile_count = off; off += up((int64_t)w.tiles * 4);
  w.tile_list = off;  off += up((int64_t)w.tiles * n_chunks * 4);
  w.partials = off;   off += up((int64_t)w.n_items * n * (int64_t)sizeof(Partial));
  w.exact = off;      off += up((int64_t)n * kExactParts * (int64_t)sizeof(ExactPartial));
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

int b200icp_s2m_chunk(void) { return kChunk; }

int64_t b200icp_s2m_workspace_bytes(int32_t n_scan, int64_t m) {
  if (n_scan < 1 || m < 1) return -1;
  return layout_workspace(n_scan, m).total;
}

int b200icp_s2m_prepare_map(const b200icp_s2m_shard* shard, void* stream) {
  if (!shard || !shard->points || !shard->cx || !shard->cy || !shard->chunk_origin || !shard->chunk_radius)
    return fail("s2m_prepare_map: NULL pointer", B200ICP_ERR_INVALID_ARGUMENT);
  if (shard->m < 1) return fail("s2m_prepare_map: empty shard", B200ICP_ERR_INVALID_ARGUMENT);
  if (shard->dtype != B200ICP_F32 && shard->dtype != B200ICP_F64)
    return fail("s2m_prepare_map: bad dtype", B200ICP_ERR_INVALID_ARGUMENT);
  const int64_t n_chunks = (shard->m + kChunk - 1) / kChunk;
  if (n_chunks > (1LL << 24)) return fail("s2m_prepare_map: shard too large", B200ICP_ERR_UNSUPPORTED_SHAPE);
  s2m_prepare_kernel<<<(unsigned)n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      shard->points, shard->dtype, shard->m, shard->cx, shard->cy, shard->chunk_origin,
      shard->chunk_radius);
  return cuda_check("s2m_prepare_kernel");
}

int b200icp_s2m_init(const void* scan, int32_t dtype, int32_t n, const double* init_pose,
                     double* src64, b200icp_s2m_state* state, void* stream) {
  if (!scan || !src64 || !state || n < 1) return fail("s2m_init: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  if (dtype != B200ICP_F32 && dtype != B200ICP_F64) return fail("s2m_init: bad dtype", B200ICP_ERR_INVALID_ARGUMENT);
  s2m_init_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      scan, dtype, n, init_pose, src64, state);
  return cuda_check("s2m_init_kernel");
}

int b200icp_s2m_bound(const b200icp_s2m_shard* shard, const double* src64, int32_t n, float* ub,
                      const b200icp_s2m_state* state, void* stream) {
  if (!shard || !src64 || !ub || !state || n < 1) return fail("s2m_bound: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_chunks = (int)((shard->m + kChunk - 1) / kChunk);
  if (cudaMemsetAsync(ub, 0x7f, (size_t)n * 4, st) != cudaSuccess) return cuda_check("cudaMemsetAsync");
  s2m_bound_kernel<<<dim3((n + 127) / 128, (n_chunks + kBoundChunks - 1) / kBoundChunks), 128, 0, st>>>(
      shard->chunk_origin, shard->chunk_radius, n_chunks, src64, n, ub, state);
  return cuda_check("s2m_bound_kernel");
}

int b200icp_s2m_search(const b200icp_s2m_shard* shard, const double* src64, int32_t n, const float* ub,
                       b200icp_s2m_record* records, void* workspace, int64_t workspace_bytes,
                       const b200icp_s2m_state* state, void* stream) {
  if (!shard || !src64 || !ub || !records || !workspace || !state || n < 1)
    return fail("s2m_search: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  if (workspace_bytes < b200icp_s2m_workspace_bytes(n, shard->m))
    return fail("s2m_search: workspace too small", B200ICP_ERR_INVALID_ARGUMENT);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_chunks = (int)((shard->m + kChunk - 1) / kChunk);
  const Workspace w = layout_workspace(n, shard->m);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  int32_t* amb_count = reinterpret_cast<int32_t*>(base);
  int32_t* amb_list = reinterpret_cast<int32_t*>(base + w.amb_list);
  int32_t* tile_count = reinterpret_cast<int32_t*>(base + w.tile_count);
  int32_t* tile_list = reinterpret_cast<int32_t*>(base + w.tile_list);
  Partial* partials = reinterpret_cast<Partial*>(base + w.partials);
  if (cudaMemsetAsync(amb_count, 0, 64, st) != cudaSuccess) return cuda_check("cudaMemsetAsync");
  int rc = B200ICP_OK;
  s2m_cull_kernel<<<w.tiles, 256, 0, st>>>(shard->chunk_origin, shard->chunk_radius, n_chunks, src64, n,
                                          ub, tile_count, tile_list, state);
  rc = cuda_check("s2m_cull_kernel");
  if (rc) return rc;
  s2m_sweep_kernel<<<dim3(w.n_items, w.tiles), kSweepThreads, 0, st>>>(
      shard->cx, shard->cy, shard->chunk_origin, shard->chunk_radius, n_chunks, tile_count, tile_list,
      src64, n, partials, state);
  rc = cuda_check("s2m_sweep_kernel");
  if (rc) return rc;
  s2m_resolve_kernel<<<(n + 127) / 128, 128, 0, st>>>(
      shard->points, shard->dtype, shard->m, shard->global_offset, shard->cx, shard->cy,
      shard->chunk_origin, shard->chunk_radius, src64, n, partials, tile_count, records, amb_list,
      amb_count, state);
  rc = cuda_check("s2m_resolve_kernel");
  if (rc) return rc;
  ExactPartial* exact_partials = reinterpret_cast<ExactPartial*>(base + w.exact);
  s2m_exact_scan_kernel<<<dim3(kExactParts, kExactRows), 256, 0, st>>>(
      shard->points, shard->dtype, shard->m, src64, amb_list, amb_count, exact_partials, state);
  rc = cuda_check("s2m_exact_scan_kernel");
  if (rc) return rc;
  s2m_exact_reduce_kernel<<<32, 256, 0, st>>>(shard->points, shard->dtype, shard->global_offset,
                                              amb_list, amb_count, exact_partials, records, state);
  return cuda_check("s2m_exact_reduce_kernel");
}

int b200icp_peer_alloc(int64_t bytes, void** ptr_out, void* handle_out) {
  if (bytes < 1 || !ptr_out || !handle_out) return fail("peer_alloc: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return cuda_check("cudaMalloc");
  if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess) return cuda_check("cudaMemset");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return cuda_check("cudaIpcGetMemHandle"); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, 64);
  *ptr_out = p;
  return B200ICP_OK;
}

int b200icp_peer_open(const void* handle, void** ptr_out) {
  if (!handle || !ptr_out) return fail("peer_open: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return cuda_check("cudaIpcOpenMemHandle");
  *ptr_out = p;
  return B200ICP_OK;
}

int b200icp_peer_close(void* ptr) {
  if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) return cuda_check("cudaIpcCloseMemHandle");
  return B200ICP_OK;
}

int b200icp_peer_free(void* ptr) {
  if (ptr && cudaFree(ptr) != cudaSuccess) return cuda_check("cudaFree");
  return B200ICP_OK;
}

int b200icp_s2m_publish(const b200icp_s2m_record* records, int32_t n, void* const* peers, int32_t world,
                        int32_t rank, int32_t slot, int64_t seq, void* counter,
                        const b200icp_s2m_state* state, void* stream) {
  if (!records || !peers || !counter || !state || n < 1 || world < 1 || rank < 0 || rank >= world || (slot != 0 && slot != 1))
    return fail("s2m_publish: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  s2m_publish_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      records, n, peers, world, rank, slot, (long long)seq, reinterpret_cast<unsigned int*>(counter), state);
  return cuda_check("s2m_publish_kernel");
}

int b200icp_s2m_wait(const void* my_buffer, int32_t n, int32_t world, int32_t slot, int64_t seq,
                     b200icp_s2m_state* state, void* stream) {
  if (!my_buffer || !state || n < 1 || world < 1 || world > 1024 || (slot != 0 && slot != 1))
    return fail("s2m_wait: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  s2m_wait_kernel<<<1, ((world + 31) / 32) * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      my_buffer, n, world, slot, (long long)seq, state);
  return cuda_check("s2m_wait_kernel");
}

int b200icp_s2m_update(const b200icp_s2m_record* records_all, int32_t n_ranks, double* src64,
                       int32_t n, int32_t max_iterations, double tolerance, double max_corr_dist,
                       int32_t* idx_out, b200icp_s2m_state* state, void* stream) {
  if (!records_all || !src64 || !state || n < 1 || n_ranks < 1)
    return fail("s2m_update: bad arguments", B200ICP_ERR_INVALID_ARGUMENT);
  s2m_update_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      records_all, n_ranks, src64, n, max_iterations, tolerance, max_corr_dist, idx_out, state);
  return cuda_check("s2m_update_kernel");
}

}  // extern "C"
