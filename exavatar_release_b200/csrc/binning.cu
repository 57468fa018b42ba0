// binning.cu -- K2: duplicate-with-keys scattered straight into per-tile segments; K3: per-tile depth sort.
//
// Replaces the reference rasteriser's duplicateWithKeys + global 64-bit cub::DeviceRadixSort (5-6 passes over
// every duplicate in HBM) + identifyTileRanges (App. A.2).  Here every tile owns a
// contiguous segment (offsets from project.cu's tile scan), K2 drops (depth, id) pairs into it, and K3 sorts each
// segment inside one CTA's shared memory: each key crosses HBM once in and its id once out.
//
// Order contract (App. A.2): ascending view depth, ties by ascending Gaussian index.  The 64-bit sort key
// (depth_bits << 32 | id) gives exactly that for positive floats and makes the result independent of the order in
// which the scatter's atomics landed, i.e. the pipeline is deterministic.
#include "common.cuh"

namespace b2r {

__global__ void __launch_bounds__(256) scatter_kernel(const B2RScene sc, const Ctx cx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sc.P) return;
  const int4 aux = cx.aux[i];
  if (aux.w == 0) return;
  const float4 g0 = reinterpret_cast<const float4*>(cx.geom + i)[0];
  const float4 g1 = reinterpret_cast<const float4*>(cx.geom + i)[1];
  const int x0 = aux.x & 0xffff, y0 = aux.x >> 16, x1 = aux.y & 0xffff, y1 = aux.y >> 16;
  const bool no_cull = (sc.flags & B2R_FLAG_NO_TILE_CULL) != 0;
  const uint32_t depth_bits = __float_as_uint(g1.z);
  const uint64_t cap = cx.dup_capacity;
  for (int ty = y0; ty < y1; ty++)
    for (int tx = x0; tx < x1; tx++) {
      bool keep = true;
      if (!no_cull) {  // must be the same predicate as in project_kernel
        const float rx0 = (float)(tx * TILE), ry0 = (float)(ty * TILE);
        const float rx1 = fminf(rx0 + (float)(TILE - 1), (float)(sc.width - 1));
        const float ry1 = fminf(ry0 + (float)(TILE - 1), (float)(sc.height - 1));
        keep = !(region_max_p2(g0.x, g0.y, g0.z, g0.w, g1.x, rx0, ry0, rx1, ry1) < g1.w);
      }
      if (keep) {
        const uint32_t pos = atomicAdd(cx.tile_cursor + ty * cx.gx + tx, 1u);
        if (pos < cap) cx.keys[pos] = make_uint2(depth_bits, (uint32_t)i);
      }
    }
}

// All-ascending bitonic network ("flip" then "disperse" steps).  Indices >= n behave as +inf keys, so no padding is
// stored and n need not be a power of two.
template <typename KeyPtr>
__device__ __forceinline__ void bitonic_sort(KeyPtr key, const int n, const int npow2) {
  const int half = npow2 >> 1;
  for (int k = 2; k <= npow2; k <<= 1) {
    const int hk = k >> 1;
    for (int t = threadIdx.x; t < half; t += blockDim.x) {  // flip: i <-> block_end - offset
      const int b = t / hk, off = t - b * hk;
      const int i = b * k + off, j = b * k + k - 1 - off;
      if (j < n) {
        const unsigned long long a = key[i], c = key[j];
        if (a > c) { key[i] = c; key[j] = a; }
      }
    }
    __syncthreads();
    for (int s = hk >> 1; s > 0; s >>= 1) {  // disperse: i <-> i + s
      for (int t = threadIdx.x; t < half; t += blockDim.x) {
        const int b = t / s, off = t - b * s;
        const int i = 2 * s * b + off, j = i + s;
        if (j < n) {
          const unsigned long long a = key[i], c = key[j];
          if (a > c) { key[i] = c; key[j] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per tile; handles tiles whose list length n satisfies lo < n <= hi.  Lists longer than SMEM_CAP are sorted
// in place in global memory by the same network (rare: > 16k splats on one tile).
template <int SMEM_CAP>
__global__ void __launch_bounds__(256) sort_tiles_kernel(const Ctx cx, const int lo, const int hi) {
  extern __shared__ __align__(16) unsigned long long skeys[];
  const int t = blockIdx.x;
  const uint2 r = cx.ranges[t];
  const int n = (int)(r.y - r.x);
  if (n <= lo || n > hi) return;
  const uint2* src = cx.keys + r.x;
  uint32_t* dst = cx.dup_ids + r.x;
  if (n == 1) {
    if (threadIdx.x == 0) dst[0] = src[0].y;
    return;
  }
  int npow2 = 2;
  while (npow2 < n) npow2 <<= 1;
  if (n <= SMEM_CAP) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint2 kv = src[i];
      skeys[i] = ((unsigned long long)kv.x << 32) | kv.y;
    }
    __syncthreads();
    bitonic_sort(skeys, n, npow2);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (uint32_t)skeys[i];
  } else {
    // global fallback: keys are stored (depth_bits, id) = little-endian (lo, hi) words, so re-pack to depth-major first
    unsigned long long* gk = reinterpret_cast<unsigned long long*>(cx.keys + r.x);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint2 kv = src[i];
      gk[i] = ((unsigned long long)kv.x << 32) | kv.y;
    }
    __syncthreads();
    bitonic_sort(gk, n, npow2);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (uint32_t)gk[i];
  }
}

constexpr int SORT_SMALL = 2048;   // 16 KB of keys: several CTAs per SM
constexpr int SORT_LARGE = 16384;  // 128 KB of keys: one CTA per SM

int launch_binning(const B2RScene& sc, const Ctx& cx, bool rescan, cudaStream_t st) {
  // the two-phase entry re-derives ranges and cursors for the capacity the caller finally chose
  if (rescan) launch_tile_scan(cx, st);
  if (sc.P > 0) scatter_kernel<<<(sc.P + 255) / 256, 256, 0, st>>>(sc, cx);
  cudaFuncSetAttribute(sort_tiles_kernel<SORT_LARGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_LARGE * 8);
  sort_tiles_kernel<SORT_SMALL><<<cx.tiles, 256, SORT_SMALL * 8, st>>>(cx, 0, SORT_SMALL);
  sort_tiles_kernel<SORT_LARGE><<<cx.tiles, 256, SORT_LARGE * 8, st>>>(cx, SORT_SMALL, 0x7fffffff);
  return check_launch();
}

}  // namespace b2r
