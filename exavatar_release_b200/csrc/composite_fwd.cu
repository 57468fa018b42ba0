// composite_fwd.cu -- K4: forward alpha-composite (App. A.3), one CTA per 16x16 tile.
//
// Replaces the reference rasteriser's forward render kernel for ExAvatar's render path
// (avatar/common/nets/module.py:632 -> render_img, render_depthmap, render_mask).
//
// Design (B200): the kernel is issue-bound, not HBM-bound, so the structure is built around
// removing per-pixel work rather than around bytes:
//   * the tile's depth-sorted id list is streamed in batches of 256; each thread gathers one 48-byte splat record
//     (three 16-byte cp.async = LDGSTS, no register staging) into a double-buffered shared-memory stage while the
//     previous batch is being composited;
//   * each warp owns an 8x4 pixel sub-tile.  For every 32 staged splats the warp runs ONE lane-parallel test
//     "can this splat reach alpha >= 1/255 anywhere in my 8x4 rect" (region_max_p2), ballots, and walks only the
//     surviving splats with lanes = pixels.  For ExAvatar's millimetre-scale avatar splats (radius 3-6 px) this
//     removes most (pixel, splat) evaluations of a 16x16 tile; every dropped pair is one App. A.3 would skip;
//   * warps retire independently (all 32 pixels saturated, T(1-a) < 1e-4) and the CTA stops staging when all have;
//   * exp via a single MUFU.EX2 on the pre-scaled conic; outputs leave as 16-byte vector stores.
#include "common.cuh"

namespace b2r {

constexpr int FWD_BATCH = 256;

struct FwdStage {
  float4 a[FWD_BATCH];  // px, py, A2, B2
  float4 b[FWD_BATCH];  // C2, opacity, depth, thr2
  float4 c[FWD_BATCH];  // r, g, b, bits
};

__device__ __forceinline__ void store4(float* base, bool vec_ok, int lane, float v, bool inside) {
  if (vec_ok) {
    const float v1 = __shfl_down_sync(0xffffffffu, v, 1);
    const float v2 = __shfl_down_sync(0xffffffffu, v, 2);
    const float v3 = __shfl_down_sync(0xffffffffu, v, 3);
    if ((lane & 3) == 0 && inside) *reinterpret_cast<float4*>(base) = make_float4(v, v1, v2, v3);
  } else if (inside) {
    *base = v;
  }
}

__global__ void __launch_bounds__(256) composite_fwd_kernel(const B2RScene sc, const Ctx cx, const B2RForwardOutputs out,
                                                            const int vec_ok) {
  __shared__ FwdStage stage[2];
  const int tile = blockIdx.x;
  const int tx = tile % cx.gx, ty = tile / cx.gx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = sc.width, H = sc.height;
  const int wx0 = tx * TILE + (warp & 1) * 8, wy0 = ty * TILE + (warp >> 1) * 4;
  const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float rx0 = (float)wx0, ry0 = (float)wy0;
  const float rx1 = fminf((float)(wx0 + 7), (float)(W - 1)), ry1 = fminf((float)(wy0 + 3), (float)(H - 1));

  const uint2 range = cx.ranges[tile];
  const int n = (int)(range.y - range.x);
  const uint32_t* ids = cx.dup_ids + range.x;
  const int nb = (n + FWD_BATCH - 1) / FWD_BATCH;

  float T = 1.f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Dp = 0.f, Aa = 0.f;
  uint32_t last = 0;
  bool done = !inside;

  auto issue = [&](int b) {
    const int idx = b * FWD_BATCH + threadIdx.x;
    if (idx < n) {
      const uint32_t id = __ldg(ids + idx);
      const float4* src = reinterpret_cast<const float4*>(cx.geom + id);
      FwdStage& s = stage[b & 1];
      cp_async16(&s.a[threadIdx.x], src);
      cp_async16(&s.b[threadIdx.x], src + 1);
      cp_async16(&s.c[threadIdx.x], src + 2);
    }
    cp_async_commit();
  };

  int staged = 0;
  if (nb > 0) issue(0);
  for (int b = 0; b < nb; b++) {
    cp_async_wait<0>();
    if (__syncthreads_and(done)) break;  // batch b visible; everyone is past batch b-1
    if (b + 1 < nb) issue(b + 1);
    const int count = min(FWD_BATCH, n - b * FWD_BATCH);
    staged += count;
    const FwdStage& s = stage[b & 1];
    bool warp_live = __any_sync(0xffffffffu, !done);
    for (int c0 = 0; c0 < count && warp_live; c0 += 32) {
      const int idx = c0 + lane;
      bool hit = false;
      if (idx < count) {
        const float4 a = s.a[idx];
        const float4 bb = s.b[idx];
        hit = !(region_max_p2(a.x, a.y, a.z, a.w, bb.x, rx0, ry0, rx1, ry1) < bb.w);
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const int j = c0 + k;
        const float4 a = s.a[j];
        const float4 bb = s.b[j];
        const float4 col = s.c[j];
        if (!done) {
          const float dx = a.x - pxf, dy = a.y - pyf;
          const float p2 = a.z * dx * dx + bb.x * dy * dy + a.w * dx * dy;
          if (p2 <= 0.f) {
            const float alpha = fminf(K_ALPHA_MAX, bb.y * ex2_approx(p2));
            if (alpha >= K_ALPHA_MIN) {
              const float test = T * (1.f - alpha);
              if (test < K_T_MIN) {
                done = true;  // this splat is not applied (App. A.3)
              } else {
                const float w = alpha * T;
                Cr = fmaf(col.x, w, Cr);
                Cg = fmaf(col.y, w, Cg);
                Cb = fmaf(col.z, w, Cb);
                Dp = fmaf(bb.z, w, Dp);
                Aa += w;
                T = test;
                last = (uint32_t)(b * FWD_BATCH + j + 1);
              }
            }
          }
        }
      }
      warp_live = __any_sync(0xffffffffu, !done);
    }
  }
  cp_async_wait<0>();

  if (threadIdx.x == 0 && staged) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_fwd), (unsigned long long)staged);

  const size_t N = (size_t)W * H;
  const size_t pix = (size_t)py * W + px;
  const float bg0 = __ldg(sc.bg), bg1 = __ldg(sc.bg + 1), bg2 = __ldg(sc.bg + 2);
  const bool v = vec_ok != 0;
  // rows of a sub-tile outside the image never store; shuffles stay warp-uniform
  store4(out.color + pix, v, lane, fmaf(T, bg0, Cr), inside);
  store4(out.color + N + pix, v, lane, fmaf(T, bg1, Cg), inside);
  store4(out.color + 2 * N + pix, v, lane, fmaf(T, bg2, Cb), inside);
  store4(out.depth + pix, v, lane, Dp, inside);
  store4(out.alpha + pix, v, lane, Aa, inside);
  store4(cx.final_T + pix, v, lane, T, inside);
  store4(reinterpret_cast<float*>(cx.n_contrib) + pix, v, lane, __uint_as_float(last), inside);
}

int launch_composite_fwd(const B2RScene& sc, const Ctx& cx, const B2RForwardOutputs& out, cudaStream_t st) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int vec_ok = (sc.width % 4 == 0) && al16(out.color) && al16(out.depth) && al16(out.alpha) && al16(cx.final_T) &&
                     al16(cx.n_contrib);
  composite_fwd_kernel<<<cx.tiles, 256, 0, st>>>(sc, cx, out, vec_ok);
  return check_launch();
}

}  // namespace b2r
