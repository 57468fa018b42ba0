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
  float4 c[FWD_BATCH];  // r, g, b, id bits
};
constexpr int FWD_GROUP = 4;   // splats blended per trip of the hit loop (their evaluations overlap: ILP 4)
constexpr int FWD_CQ = 36;     // survivor queue: <= 3 left over + 32 new per chunk (+ pad)
struct FwdCompact {            // warp-private queue of cull survivors in list order
  float4 r[3][FWD_CQ];         // [0] px,py,A2,B2  [1] C2,opacity,depth,1-based list position (int bits)  [2] r,g,b,-
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
  __shared__ FwdCompact compact[2];
  B2R_TRACE_BEGIN();
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

  // Blend state.  A finished pixel (T(1-a) < 1e-4 reached, or outside the image) is marked by the SIGN of T: the
  // magnitude is still the transmittance that goes to final_T, and "alive" is one compare folded into the chain of
  // predicate tests below -- no separate flag register to maintain.
  float T = inside ? 1.f : -1.f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Dp = 0.f, Aa = 0.f;
  uint32_t last = 0;
  FwdCompact& cw = compact[warp];
  const unsigned lanes_below = (1u << lane) - 1u;
  int fill = 0;  // warp-uniform: survivors waiting in the queue (< FWD_GROUP between chunks)

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

  // One trip = FWD_GROUP queued splats.  Their exponent evaluations are independent of the blend state and of each
  // other, so they overlap (shared loads, FMA chain, MUFU); only the short T recurrence that follows is serial.
  // Branch-free (App. A.3): a splat is skipped unless the pixel is alive, power <= 0 and alpha >= 1/255; a splat that
  // would drop T(1-a) below 1e-4 finishes the pixel WITHOUT being applied.
  auto blend_group = [&](const int k) {
    float al[FWD_GROUP];
    bool ok[FWD_GROUP];
    float4 col[FWD_GROUP];
    float dep[FWD_GROUP];
    uint32_t pos[FWD_GROUP];
#pragma unroll
    for (int u = 0; u < FWD_GROUP; u++) {
      const float4 a = cw.r[0][k + u], bb = cw.r[1][k + u];
      col[u] = cw.r[2][k + u];
      const float dx = a.x - pxf, dy = a.y - pyf;
      const float p2 = a.z * dx * dx + bb.x * dy * dy + a.w * dx * dy;
      const float ar = bb.y * ex2_approx(p2);
      al[u] = fminf(K_ALPHA_MAX, ar);
      ok[u] = (ar >= K_ALPHA_MIN) & (p2 <= 0.f);
      dep[u] = bb.z;
      pos[u] = (uint32_t)__float_as_int(bb.w);
    }
#pragma unroll
    for (int u = 0; u < FWD_GROUP; u++) {
      const bool v = ok[u] & (T > 0.f);
      const float test = T * (1.f - al[u]);
      const float w = al[u] * T;
      const bool stop = v & (test < K_T_MIN);
      const bool use = v & !stop;
      Cr = use ? fmaf(col[u].x, w, Cr) : Cr;  // predicated accumulates: a skipped splat must not touch the sums at all
      Cg = use ? fmaf(col[u].y, w, Cg) : Cg;
      Cb = use ? fmaf(col[u].z, w, Cb) : Cb;
      Dp = use ? fmaf(dep[u], w, Dp) : Dp;
      Aa = use ? Aa + w : Aa;
      last = use ? pos[u] : last;
      T = use ? test : (stop ? -T : T);
    }
  };

  int staged = 0;
  if (nb > 0) issue(0);
  for (int b = 0; b < nb; b++) {
    cp_async_wait<0>();
    if (__syncthreads_and(!(T > 0.f))) break;  // batch b visible; everyone is past batch b-1
    if (b + 1 < nb) issue(b + 1);
    const int count = min(FWD_BATCH, n - b * FWD_BATCH);
    staged += count;
    const FwdStage& s = stage[b & 1];
    bool warp_live = __any_sync(0xffffffffu, T > 0.f);
    for (int c0 = 0; c0 < count && warp_live; c0 += 32) {
      const int idx = c0 + lane;
      bool hit = false;
      float4 a, bb;
      if (idx < count) {
        a = s.a[idx];
        bb = s.b[idx];
        hit = !(region_max_p2(a.x, a.y, a.z, a.w, bb.x, rx0, ry0, rx1, ry1) < bb.w);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask == 0u) continue;
      // Append the survivors (list order) to the warp's queue: the hit loop then walks consecutive slots, FWD_GROUP at
      // a time, instead of find-first-set + index arithmetic per splat; what does not fill a group waits for the
      // next chunk.
      if (hit) {
        const int slot = fill + __popc(mask & lanes_below);
        cw.r[0][slot] = a;
        cw.r[1][slot] = make_float4(bb.x, bb.y, bb.z, __int_as_float(b * FWD_BATCH + idx + 1));
        cw.r[2][slot] = s.c[idx];
      }
      fill += __popc(mask);
      __syncwarp();
      int k = 0;
      for (; k + FWD_GROUP <= fill; k += FWD_GROUP) blend_group(k);
      const int left = fill - k;
      if (k > 0) {  // move the <= 3 leftover records to the front (sources are slots >= 4: no overlap)
        __syncwarp();
        if (lane < 3 * left) {
          const int t = (lane >= left) + (lane >= 2 * left), j = lane - t * left;
          cw.r[t][j] = cw.r[t][k + j];
        }
      }
      fill = left;
      warp_live = __any_sync(0xffffffffu, T > 0.f);  // also orders the queue reads / moves before the next append
    }
  }
  if (fill > 0) {  // flush: pad the last group with splats of opacity 0 (alpha 0 < 1/255 => skipped)
    if (lane >= fill && lane < FWD_GROUP) {
      cw.r[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      cw.r[1][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      cw.r[2][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    blend_group(0);
  }
  cp_async_wait<0>();

  // consumed_fwd counts list entries per TILE: four quarter-CTAs each add a quarter of what they staged
  if (threadIdx.x == 0 && staged) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_fwd), (unsigned long long)staged);

  T = fabsf(T);
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
  B2R_TRACE_END(n);
}

int launch_composite_fwd(const B2RScene& sc, const Ctx& cx, const B2RForwardOutputs& out, cudaStream_t st) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int vec_ok = (sc.width % 4 == 0) && al16(out.color) && al16(out.depth) && al16(out.alpha) && al16(cx.final_T) &&
                     al16(cx.n_contrib);
  {
    ProfScope p(K_COMPOSITE_FWD, st);
    launch_k(composite_fwd_kernel, cx.tiles * 4, FWD_THREADS, 0, st, false, sc, cx, out, vec_ok);
  }
  return check_launch();
}

}  // namespace b2r

#ifdef B2R_CTA_TRACE
extern "C" int b2r_debug_trace_fwd(unsigned long long* buf) {
  return (int)cudaMemcpyToSymbol(b2r::g_cta_trace, &buf, sizeof(buf));
}
#endif
