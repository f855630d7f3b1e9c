// voxel.cu -- 2D voxel-grid down-sampling on the device (sm_100a).
//
// The reference down-samples both clouds before every registration
// (duc/ICP_LIDAR/gicp_lidar.py:8-11,20-21: point_cloud.voxel_down_sample(voxel_size)) and removes
// duplicates the same way (duc/ICP_LIDAR/process.py:68-73); labels_segmentation/d.py:10-16 spells
// the 2D grid out: cell = floor(p / voxel_size).  Open3D keeps ONE point per occupied voxel, the
// mean of the points that fell into it.  Its output order is the iteration order of a hash map
// (unspecified); here the voxels come out sorted by (cell_y, cell_x), which is the order d.py
// builds with lexsort.  Parity-unpinned against Open3D (not vendored by the reference): the
// semantics are pinned by oracle.icp_oracle.voxel_down_sample_2d.
//
// keys (cell_y, cell_x) -> cub radix sort of (key, index) -> one thread per point: segment heads
// average their run in the original point order (deterministic).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cstdint>
#include <cstdio>

#include "b200icp.h"

void b200icp_set_error_str(const char* msg);

namespace {

__device__ __forceinline__ double2 vx_load(const void* base, int dtype, int64_t i) {
  if (dtype == B200ICP_F64) return reinterpret_cast<const double2*>(base)[i];
  const float2 v = reinterpret_cast<const float2*>(base)[i];
  return make_double2((double)v.x, (double)v.y);
}

__global__ void voxel_key_kernel(const void* points, int dtype, int64_t n, double inv_voxel,
                                 uint64_t* keys, int32_t* idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 q = vx_load(points, dtype, i);
  // cell = floor(p / voxel) (d.py:11-12); biased to unsigned so the radix order is (cy, cx)
  const int64_t gx = (int64_t)floor(q.x * inv_voxel), gy = (int64_t)floor(q.y * inv_voxel);
  keys[i] = ((uint64_t)(uint32_t)((int32_t)gy ^ 0x80000000)) << 32 | (uint32_t)((int32_t)gx ^ 0x80000000);
  idx[i] = (int32_t)i;
}

__global__ void voxel_mean_kernel(const void* points, int dtype, int64_t n, const uint64_t* keys,
                                  const int32_t* idx, int32_t* head_flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head_flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per sorted position; heads walk their run (radix sort is stable, so a run lists its
// points in the original order) and write the mean to the slot given by the scanned head count
__global__ void voxel_emit_kernel(const void* points, int dtype, int64_t n, const uint64_t* keys,
                                  const int32_t* idx, const int32_t* head_flag,
                                  const int32_t* head_rank, void* out, int64_t* count_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == n - 1) *count_out = head_rank[i] + (int64_t)1;     // inclusive scan - 1 = slot
  if (!head_flag[i]) return;
  double sx = 0.0, sy = 0.0;
  int64_t k = i;
  const uint64_t key = keys[i];
  for (; k < n && keys[k] == key; ++k) {
    const double2 q = vx_load(points, dtype, idx[k]);
    sx += q.x; sy += q.y;
  }
  const double inv = 1.0 / (double)(k - i);
  const int64_t slot = head_rank[i];
  if (dtype == B200ICP_F64) reinterpret_cast<double2*>(out)[slot] = make_double2(sx * inv, sy * inv);
  else reinterpret_cast<float2*>(out)[slot] = make_float2((float)(sx * inv), (float)(sy * inv));
}

// inclusive scan of head flags minus one (single CTA; n is a scan, i.e. <= a few 10^5 points)
__global__ void __launch_bounds__(1024) voxel_scan_kernel(const int32_t* flag, int64_t n, int32_t* rank) {
  __shared__ int32_t tot[1024];
  const int tid = threadIdx.x;
  const int64_t per = (n + 1023) / 1024, b = min(n, tid * per), e = min(n, b + per);
  int32_t s = 0;
  for (int64_t k = b; k < e; ++k) s += flag[k];
  const int lane = tid & 31, warp = tid >> 5;
  int32_t inc = s;                                      // warp-shuffle scan of the segment sums
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int32_t w = tot[lane];
    int32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    tot[lane] = winc - w;
  }
  __syncthreads();
  int32_t run = tot[warp] + inc - s;
  for (int64_t k = b; k < e; ++k) { run += flag[k]; rank[k] = run - 1; }
}

struct VoxelWs {
  int64_t keys_in, keys_out, idx_in, idx_out, flags, rank, cub, total;
  size_t cub_bytes;
};

VoxelWs voxel_layout(int64_t n) {
  auto up = [](int64_t b) { return (b + 255) / 256 * 256; };
  VoxelWs w;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  w.cub_bytes = cub_bytes;
  int64_t off = 0;
  w.keys_in = off;  off += up(n * 8);
  w.keys_out = off; off += up(n * 8);
  w.idx_in = off;   off += up(n * 4);
  w.idx_out = off;  off += up(n * 4);
  w.flags = off;    off += up(n * 4);
  w.rank = off;     off += up(n * 4);
  w.cub = off;      off += up((int64_t)cub_bytes);
  w.total = off;
  return w;
}

int vfail(const char* msg, int code) {
  b200icp_set_error_str(msg);
  return code;
}

}  // namespace

extern "C" {

int64_t b200icp_voxel_workspace_bytes(int64_t n) {
  if (n < 0 || n > 0x7fffffffLL) return -1;
  return voxel_layout(n > 0 ? n : 1).total;
}

int b200icp_voxel_downsample(const void* points, int32_t dtype, int64_t n, double voxel_size,
                             void* out_points, int64_t* count_out, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  if (!points || !out_points || !count_out || !workspace) return vfail("voxel_downsample: NULL pointer", B200ICP_ERR_INVALID_ARGUMENT);
  if (dtype != B200ICP_F32 && dtype != B200ICP_F64) return vfail("voxel_downsample: bad dtype", B200ICP_ERR_INVALID_ARGUMENT);
  if (n < 1 || n > 0x7fffffffLL || !(voxel_size > 0.0)) return vfail("voxel_downsample: need n >= 1 and voxel_size > 0", B200ICP_ERR_INVALID_ARGUMENT);
  const VoxelWs w = voxel_layout(n);
  if (workspace_bytes < w.total) return vfail("voxel_downsample: workspace too small", B200ICP_ERR_INVALID_ARGUMENT);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  uint64_t* keys_in = reinterpret_cast<uint64_t*>(base + w.keys_in);
  uint64_t* keys_out = reinterpret_cast<uint64_t*>(base + w.keys_out);
  int32_t* idx_in = reinterpret_cast<int32_t*>(base + w.idx_in);
  int32_t* idx_out = reinterpret_cast<int32_t*>(base + w.idx_out);
  int32_t* flags = reinterpret_cast<int32_t*>(base + w.flags);
  int32_t* rank = reinterpret_cast<int32_t*>(base + w.rank);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  voxel_key_kernel<<<blocks, 256, 0, st>>>(points, dtype, n, 1.0 / voxel_size, keys_in, idx_in);
  size_t cub_bytes = w.cub_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(base + w.cub, cub_bytes, keys_in, keys_out, idx_in, idx_out,
                                                  (int)n, 0, 64, st);
  if (e != cudaSuccess) return vfail(cudaGetErrorString(e), B200ICP_ERR_CUDA);
  voxel_mean_kernel<<<blocks, 256, 0, st>>>(points, dtype, n, keys_out, idx_out, flags);
  voxel_scan_kernel<<<1, 1024, 0, st>>>(flags, n, rank);
  voxel_emit_kernel<<<blocks, 256, 0, st>>>(points, dtype, n, keys_out, idx_out, flags, rank, out_points, count_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return vfail(cudaGetErrorString(e), B200ICP_ERR_CUDA);
  return B200ICP_OK;
}

}  // extern "C"
