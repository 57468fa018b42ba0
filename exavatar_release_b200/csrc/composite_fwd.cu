// composite_fwd.cu -- K4: forward alpha-composite (App. A.3).
//
// Replaces the reference rasteriser's forward render kernel for ExAvatar's render path
// (avatar/common/nets/module.py:632 -> render_img, render_depthmap, render_mask).
//
// Design (B200).  The kernel is issue- and latency-bound, not HBM-bound (ncu in
// profiles/), so the structure removes per-pixel work and idle SMs rather than bytes:
//   * work unit = one 8x8 quarter of a 16x16 tile, one 64-thread CTA (2 warps, each an 8x4 pixel rect).  Tile cost
//     spans three orders of magnitude; quarter-tile CTAs launched longest-list-first (cx.tile_order) keep all 148 SMs
//     busy where whole-tile CTAs in index order left them idle > 50 % of the time (profiles/r01_notes.md);
//   * the tile's depth-sorted id list is streamed in batches of 128; each thread gathers two 48-byte splat records
//     (three 16-byte cp.async = LDGSTS each, no register staging) into a double-buffered shared-memory stage while
//     the previous batch is composited;
//   * for every 32 staged splats a warp runs ONE lane-parallel test "can this splat reach alpha >= 1/255 anywhere in
//     my 8x4 rect" (region_max_p2), ballots, and walks only the survivors with lanes = pixels.  For ExAvatar's
//     millimetre-scale avatar splats (radius 3-6 px) this removes most (pixel, splat) evaluations of a 16x16 tile;
//     every dropped pair is one App. A.3 would skip anyway;
//   * warps retire independently (all 32 pixels saturated, T(1-a) < 1e-4); the CTA stops staging when both have;
//   * exp via one MUFU.EX2 on the pre-scaled conic; outputs leave as 16-byte vector stores.
#include "common.cuh"

namespace b2r {

constexpr int FWD_THREADS = 64;
constexpr int FWD_BATCH = 128;
constexpr int FWD_PER_THREAD = FWD_BATCH / FWD_THREADS;

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

__global__ void __launch_bounds__(FWD_THREADS) composite_fwd_kernel(const B2RScene sc, const Ctx cx,
                                                                    const B2RForwardOutputs out, const int vec_ok) {
  __shared__ FwdStage stage[2];
  const int tile = (int)cx.tile_order[blockIdx.x >> 2];
  const int quad = blockIdx.x & 3;
  const int tx = tile % cx.gx, ty = tile / cx.gx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = sc.width, H = sc.height;
  const int wx0 = tx * TILE + (quad & 1) * 8, wy0 = ty * TILE + (quad >> 1) * 8 + warp * 4;
  if (wx0 >= W || ty * TILE + (quad >> 1) * 8 >= H) return;  // quarter entirely outside the image (CTA-uniform)
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
    FwdStage& s = stage[b & 1];
#pragma unroll
    for (int u = 0; u < FWD_PER_THREAD; u++) {
      const int slot = threadIdx.x + u * FWD_THREADS;
      const int idx = b * FWD_BATCH + slot;
      if (idx < n) {
        const uint32_t id = __ldg(ids + idx);
        const float4* src = reinterpret_cast<const float4*>(cx.geom + id);
        cp_async16(&s.a[slot], src);
        cp_async16(&s.b[slot], src + 1);
        cp_async16(&s.c[slot], src + 2);
      }
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
      // Walk the surviving splats two at a time: the two exponent evaluations are independent (ILP 2); only the
      // transmittance recurrence is serial, and it is applied branch-free (selects), so the only branches in this
      // loop are warp-uniform.
      while (mask) {
        const int k0 = __ffs(mask) - 1;
        mask &= mask - 1;
        const bool two = mask != 0;
        const int k1 = two ? __ffs(mask) - 1 : k0;
        mask &= mask - 1;  // no-op when mask == 0
        const int j0 = c0 + k0, j1 = c0 + k1;
        const float4 a0 = s.a[j0], b0 = s.b[j0], col0 = s.c[j0];
        const float4 a1 = s.a[j1], b1 = s.b[j1], col1 = s.c[j1];
        const float dx0 = a0.x - pxf, dy0 = a0.y - pyf;
        const float dx1 = a1.x - pxf, dy1 = a1.y - pyf;
        const float p20 = a0.z * dx0 * dx0 + b0.x * dy0 * dy0 + a0.w * dx0 * dy0;
        const float p21 = a1.z * dx1 * dx1 + b1.x * dy1 * dy1 + a1.w * dx1 * dy1;
        const float al0 = fminf(K_ALPHA_MAX, b0.y * ex2_approx(p20));
        const float al1 = fminf(K_ALPHA_MAX, b1.y * ex2_approx(p21));
        const bool ok0 = (p20 <= 0.f) & (al0 >= K_ALPHA_MIN);
        const bool ok1 = two & (p21 <= 0.f) & (al1 >= K_ALPHA_MIN);
        {
          const bool v = ok0 & !done;
          const float test = T * (1.f - al0);
          const bool stop = v & (test < K_T_MIN);  // this splat is not applied (App. A.3)
          done |= stop;
          const bool use = v & !stop;
          const float w = al0 * T;  // predicated accumulates: a skipped splat must not touch the sums at all
          Cr = use ? fmaf(col0.x, w, Cr) : Cr;
          Cg = use ? fmaf(col0.y, w, Cg) : Cg;
          Cb = use ? fmaf(col0.z, w, Cb) : Cb;
          Dp = use ? fmaf(b0.z, w, Dp) : Dp;
          Aa = use ? Aa + w : Aa;
          T = use ? test : T;
          last = use ? (uint32_t)(b * FWD_BATCH + j0 + 1) : last;
        }
        {
          const bool v = ok1 & !done;
          const float test = T * (1.f - al1);
          const bool stop = v & (test < K_T_MIN);
          done |= stop;
          const bool use = v & !stop;
          const float w = al1 * T;  // predicated accumulates: a skipped splat must not touch the sums at all
          Cr = use ? fmaf(col1.x, w, Cr) : Cr;
          Cg = use ? fmaf(col1.y, w, Cg) : Cg;
          Cb = use ? fmaf(col1.z, w, Cb) : Cb;
          Dp = use ? fmaf(b1.z, w, Dp) : Dp;
          Aa = use ? Aa + w : Aa;
          T = use ? test : T;
          last = use ? (uint32_t)(b * FWD_BATCH + j1 + 1) : last;
        }
      }
      warp_live = __any_sync(0xffffffffu, !done);
    }
  }
  cp_async_wait<0>();

  // consumed_fwd counts list entries per TILE: four quarter-CTAs each add a quarter of what they staged
  if (threadIdx.x == 0 && staged) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_fwd), (unsigned long long)staged);

  const size_t N = (size_t)W * H;
  const size_t pix = (size_t)py * W + px;
  const float bg0 = __ldg(sc.bg), bg1 = __ldg(sc.bg + 1), bg2 = __ldg(sc.bg + 2);
  const bool v = vec_ok != 0;
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
  {
    ProfScope p(K_COMPOSITE_FWD, st);
    composite_fwd_kernel<<<cx.tiles * 4, FWD_THREADS, 0, st>>>(sc, cx, out, vec_ok);
  }
  return check_launch();
}

}  // namespace b2r
