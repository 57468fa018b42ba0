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
  int4 aux = make_int4(0, 0, 0, 0);
  float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
  if (i < sc.P) {
    aux = cx.aux[i];
    if (aux.z > 0) {
      g0 = reinterpret_cast<const float4*>(cx.geom + i)[0];
      g1 = reinterpret_cast<const float4*>(cx.geom + i)[1];
    }
  }
  const uint64_t cap = cx.dup_capacity;
  const int gx = cx.gx;
  uint32_t* cursor = cx.tile_cursor;
  uint2* keys = cx.keys;
  // same enumeration and the same predicate as the counting pass in project_kernel
  warp_for_each_kept_tile(aux.z > 0, aux.x & 0xffff, aux.x >> 16, aux.y & 0xffff, aux.y >> 16, g0.x, g0.y, g0.z, g0.w,
                          g1.x, g1.w, __float_as_uint(g1.z), (uint32_t)i, (sc.flags & B2R_FLAG_NO_TILE_CULL) != 0,
                          sc.width, sc.height, [&](int tx, int ty, uint32_t depth_bits, uint32_t id) {
                            const uint32_t pos = atomicAdd(cursor + ty * gx + tx, 1u);
                            if (pos < cap) keys[pos] = make_uint2(depth_bits, id);
                          });
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

// Sorts every tile whose list length n satisfies lo < n <= hi; CTAs stride over the tiles (the small class is launched
// with one CTA per tile, the large class with one CTA per SM so that a frame without long lists costs ~2 us).  Lists
// longer than SMEM_CAP are sorted in place in global memory by the same network (rare: > 16k splats on one tile).
template <int SMEM_CAP>
__global__ void __launch_bounds__(256) sort_tiles_kernel(const Ctx cx, const int lo, const int hi) {
  extern __shared__ __align__(16) unsigned long long skeys[];
  for (int t = blockIdx.x; t < cx.tiles; t += gridDim.x) {
    const uint2 r = cx.ranges[cx.tile_order[t]];  // longest lists first
    const int n = (int)(r.y - r.x);
    if (n <= lo || n > hi) continue;
    const uint2* src = cx.keys + r.x;
    uint32_t* dst = cx.dup_ids + r.x;
    if (n == 1) {
      if (threadIdx.x == 0) dst[0] = src[0].y;
      continue;
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
      __syncthreads();  // skeys is reused by the next tile of this CTA
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
}

constexpr int SORT_SMALL = 2048;   // 16 KB of keys: several CTAs per SM
constexpr int SORT_LARGE = 16384;  // 128 KB of keys: one CTA per SM

int launch_binning(const B2RScene& sc, const Ctx& cx, bool rescan, cudaStream_t st) {
  // the two-phase entry re-derives ranges and cursors for the capacity the caller finally chose
  if (rescan) launch_tile_scan(cx, st);
  if (sc.P > 0) { ProfScope p(K_SCATTER, st); scatter_kernel<<<(sc.P + 255) / 256, 256, 0, st>>>(sc, cx); }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  cudaFuncSetAttribute(sort_tiles_kernel<SORT_LARGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_LARGE * 8);
  { ProfScope p(K_SORT_SMALL, st); sort_tiles_kernel<SORT_SMALL><<<cx.tiles, 256, SORT_SMALL * 8, st>>>(cx, 0, SORT_SMALL); }
  {
    ProfScope p(K_SORT_LARGE, st);
    sort_tiles_kernel<SORT_LARGE><<<cx.tiles < sms ? cx.tiles : sms, 256, SORT_LARGE * 8, st>>>(cx, SORT_SMALL, 0x7fffffff);
  }
  return check_launch();
}

}  // namespace b2r
