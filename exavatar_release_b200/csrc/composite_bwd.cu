// composite_bwd.cu -- K5: backward alpha-composite (App. A.4), one 64-thread CTA per 8x8 quarter tile, launched
// longest-list-first (same work decomposition as composite_fwd.cu).
//
// Replaces the reference rasteriser's backward render kernel (reached from loss.backward(), avatar/main/train.py:46,
// through the autograd node created at avatar/common/nets/module.py:632).  That kernel issues ~10 fp32 global atomics
// per (pixel, splat) -- the known bottleneck.  Here:
//   * same staging and 8x4 sub-tile culling as the forward composite, walked back to front;
//   * each warp reduces its 32 pixels' 9-10 partial gradients with a value-halving butterfly (12 shuffles instead of
//     50) and adds them to a per-CTA shared-memory accumulator, one row per staged splat;
//   * after each batch every splat row leaves the SM as three 16-byte vector reductions (REDG.E.ADD.F32x4) -- one
//     accumulation per (splat, tile) and 3 L2 transactions instead of ~2560 scalar atomics.
// Gradient conventions are App. A.6's: the 0.99 clamp is ignored, masks are constants.
//
// Accumulator row (12 floats; constants folded in by project_bwd.cu):
//   q0 = { sum dLdG*(2 gdx A2 + gdy B2), sum dLdG*(2 gdy C2 + gdx B2), sum dLdG*gdx*dx, sum dLdG*gdx*dy }
//   q1 = { sum dLdG*gdy*dy, sum G*dLdalpha, sum w*g_depth, 0 }      q2 = { sum w*g_r, sum w*g_g, sum w*g_b, 0 }
#include <cstdlib>

#include "common.cuh"

namespace b2r {

constexpr int BWD_THREADS = 64;
constexpr int BWD_BATCH = 64;
constexpr int BWD_PER_THREAD = BWD_BATCH / BWD_THREADS;

struct BwdStage {
  float4 a[BWD_BATCH];
  float4 b[BWD_BATCH];
  float4 c[BWD_BATCH];
  uint32_t id[BWD_BATCH];
};

// Sum N per-lane values across the warp: after the call, the total of value `slot` lives in v[0] of the (two) lanes
// whose bits 4..1 select that slot.  Each level halves the number of live values instead of reducing them all.
template <int N, int M>
__device__ __forceinline__ void halving_reduce(float* v, const int lane) {
  if constexpr (M >= 2) {
    constexpr int NK = (N + 1) / 2;
    const bool up = (lane & M) != 0;
#pragma unroll
    for (int i = 0; i < NK; i++) {
      const float lo = v[i];
      const float hi = (i + NK < N) ? v[i + NK] : 0.f;
      const float send = up ? lo : hi;
      const float keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
    }
    halving_reduce<NK, M / 2>(v, lane);
  } else {
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  }
}

// Which value index ends up in this lane's v[0] (or -1).
template <int NV>
__device__ __forceinline__ int halving_slot(const int lane) {
  int idx = 0, real = NV, level = NV;
#pragma unroll
  for (int m = 16; m >= 2; m >>= 1) {
    const int nk = (level + 1) / 2;
    if (lane & m) {
      idx += nk;
      real = real - nk;
    } else {
      real = min(real, nk);
    }
    level = nk;
  }
  return (real >= 1 && (lane & 1) == 0) ? idx : -1;
}

template <bool HAS_DA>
__global__ void __launch_bounds__(BWD_THREADS) composite_bwd_kernel(const B2RScene sc, const Ctx cx, const B2RBackwardArgs args,
                                                            float* __restrict__ gacc) {
  constexpr int NV = HAS_DA ? 10 : 9;
  __shared__ BwdStage stage[2];
  // one accumulator row per (warp, staged splat): a warp visits a splat at most once per batch, so the reduced
  // values are written with plain stores -- shared-memory fp32 atomicAdd would compile to a CAS spin loop
  __shared__ __align__(16) float acc[2][BWD_BATCH][12];
  __shared__ int warp_max_s[2];
  __shared__ unsigned long long touched_s[2];  // per warp: which rows of this batch hold fresh values

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
  const size_t N = (size_t)W * H;
  const size_t pix = (size_t)py * W + px;

  const uint2 range = cx.ranges[tile];
  const uint32_t* ids = cx.dup_ids + range.x;

  const int my_n = inside ? (int)cx.n_contrib[pix] : 0;
  const float T_final = inside ? cx.final_T[pix] : 0.f;
  const float g_r = inside ? __ldg(args.dL_dcolor + pix) : 0.f;
  const float g_g = inside ? __ldg(args.dL_dcolor + N + pix) : 0.f;
  const float g_b = inside ? __ldg(args.dL_dcolor + 2 * N + pix) : 0.f;
  float g_d = 0.f, g_a = 0.f;
  if (HAS_DA && inside) {
    if (args.dL_ddepth) g_d = __ldg(args.dL_ddepth + pix);
    if (args.dL_dalpha) g_a = __ldg(args.dL_dalpha + pix);
  }
  const float bg_dot = __ldg(sc.bg) * g_r + __ldg(sc.bg + 1) * g_g + __ldg(sc.bg + 2) * g_b;

  const int warp_n = __reduce_max_sync(0xffffffffu, my_n);
  if (lane == 0) warp_max_s[warp] = warp_n;
  __syncthreads();
  int nmax = 0;
  nmax = max(warp_max_s[0], warp_max_s[1]);
  if (nmax == 0) return;
  const int nb = (nmax + BWD_BATCH - 1) / BWD_BATCH;

  const int my_slot = halving_slot<NV>(lane);
  // value index -> accumulator column: mx my ca cb cc op | r g b | d
  const int my_col = my_slot < 0 ? -1 : (my_slot < 6 ? my_slot : (my_slot < 9 ? my_slot + 2 : 6));

  float T = T_final, last_alpha = 0.f;
  float acr = 0.f, acg = 0.f, acb = 0.f, lcr = 0.f, lcg = 0.f, lcb = 0.f;
  float acd = 0.f, lcd = 0.f, aca = 0.f;

  auto issue = [&](int b) {
    BwdStage& s = stage[b & 1];
#pragma unroll
    for (int u = 0; u < BWD_PER_THREAD; u++) {
      const int slot = threadIdx.x + u * BWD_THREADS;
      const int idx = b * BWD_BATCH + slot;
      if (idx < nmax) {
        const uint32_t id = __ldg(ids + idx);
        const float4* src = reinterpret_cast<const float4*>(cx.geom + id);
        cp_async16(&s.a[slot], src);
        cp_async16(&s.b[slot], src + 1);
        cp_async16(&s.c[slot], src + 2);
        s.id[slot] = id;
      }
    }
    cp_async_commit();
  };

  issue(nb - 1);
  for (int b = nb - 1; b >= 0; b--) {
    cp_async_wait<0>();
    __syncthreads();  // batch b staged; flush of batch b+1 finished
    if (b > 0) issue(b - 1);
    const int count = min(BWD_BATCH, nmax - b * BWD_BATCH);
    const BwdStage& s = stage[b & 1];
    unsigned long long touched = 0ull;
    if (warp_n > b * BWD_BATCH) {
      for (int c0 = ((count - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
        const int idx = c0 + lane;
        bool hit = false;
        if (idx < count && b * BWD_BATCH + idx < warp_n) {
          const float4 a = s.a[idx];
          const float4 bb = s.b[idx];
          hit = !(region_max_p2(a.x, a.y, a.z, a.w, bb.x, rx0, ry0, rx1, ry1) < bb.w);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = 31 - __clz(mask);
          mask &= ~(1u << k);
          const int j = c0 + k;
          const float4 a = s.a[j];
          const float4 bb = s.b[j];
          const float dx = a.x - pxf, dy = a.y - pyf;
          const float p2 = a.z * dx * dx + bb.x * dy * dy + a.w * dx * dy;
          const float G = ex2_approx(p2);
          const float alpha = fminf(K_ALPHA_MAX, bb.y * G);
          const bool valid = (b * BWD_BATCH + j < my_n) && (p2 <= 0.f) && (alpha >= K_ALPHA_MIN);
          if (!__any_sync(0xffffffffu, valid)) continue;
          touched |= 1ull << j;
          // branch-free: every lane evaluates, `valid` only selects what is kept (no divergent code in this loop)
          const float4 col = s.c[j];
          const float rcp = __fdividef(1.f, 1.f - alpha);
          const float Tn = T * rcp;
          const float w = alpha * Tn;
          const float om = 1.f - last_alpha;
          const float nacr = last_alpha * lcr + om * acr;
          const float nacg = last_alpha * lcg + om * acg;
          const float nacb = last_alpha * lcb + om * acb;
          float dLda = (col.x - nacr) * g_r + (col.y - nacg) * g_g + (col.z - nacb) * g_b;
          float nacd = 0.f, naca = 0.f;
          if (HAS_DA) {
            nacd = last_alpha * lcd + om * acd;
            naca = last_alpha + om * aca;
            dLda += (bb.z - nacd) * g_d + (1.f - naca) * g_a;
          }
          dLda = dLda * Tn + (-T_final * rcp) * bg_dot;
          const float dLdG = bb.y * dLda;  // clamp ignored (App. A.6 i)
          const float gdx = G * dx, gdy = G * dy;
          const float hx = dLdG * gdx, hy = dLdG * gdy;
          float v[NV];
          v[0] = valid ? dLdG * (2.f * gdx * a.z + gdy * a.w) : 0.f;
          v[1] = valid ? dLdG * (2.f * gdy * bb.x + gdx * a.w) : 0.f;
          v[2] = valid ? hx * dx : 0.f;
          v[3] = valid ? hx * dy : 0.f;
          v[4] = valid ? hy * dy : 0.f;
          v[5] = valid ? G * dLda : 0.f;
          v[6] = valid ? w * g_r : 0.f;
          v[7] = valid ? w * g_g : 0.f;
          v[8] = valid ? w * g_b : 0.f;
          if (HAS_DA) v[NV - 1] = valid ? w * g_d : 0.f;
          T = valid ? Tn : T;
          acr = valid ? nacr : acr; lcr = valid ? col.x : lcr;
          acg = valid ? nacg : acg; lcg = valid ? col.y : lcg;
          acb = valid ? nacb : acb; lcb = valid ? col.z : lcb;
          if (HAS_DA) {
            acd = valid ? nacd : acd; lcd = valid ? bb.z : lcd;
            aca = valid ? naca : aca;
          }
          last_alpha = valid ? alpha : last_alpha;
          halving_reduce<NV, 16>(v, lane);
          if (my_col >= 0) acc[warp][j][my_col] = v[0];
        }
      }
    }
    if (lane == 0) touched_s[warp] = touched;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < BWD_PER_THREAD; u++) {  // flush this batch: one accumulator row per (thread, u)
      const int slot = threadIdx.x + u * BWD_THREADS;
      if (slot < count) {
        const bool t0 = (touched_s[0] >> slot) & 1ull, t1 = (touched_s[1] >> slot) & 1ull;
        if (t0 | t1) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4* row0 = reinterpret_cast<const float4*>(&acc[0][slot][0]);
          const float4* row1 = reinterpret_cast<const float4*>(&acc[1][slot][0]);
          const float4 p0 = t0 ? row0[0] : z, p1 = t0 ? row0[1] : z, p2 = t0 ? row0[2] : z;
          const float4 r0 = t1 ? row1[0] : z, r1 = t1 ? row1[1] : z, r2 = t1 ? row1[2] : z;
          float* dst = gacc + (size_t)s.id[slot] * 12;
          red_add_v4(dst, p0.x + r0.x, p0.y + r0.y, p0.z + r0.z, p0.w + r0.w);
          red_add_v4(dst + 4, p1.x + r1.x, p1.y + r1.y, HAS_DA ? p1.z + r1.z : 0.f, 0.f);
          red_add_v4(dst + 8, p2.x + r2.x, p2.y + r2.y, p2.z + r2.z, 0.f);
        }
      }
    }
  }
  if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_bwd), (unsigned long long)nmax);
}

int launch_composite_bwd2(const B2RScene& sc, const Ctx& cx, const B2RBackwardArgs& a, float* gacc, cudaStream_t st);

int launch_composite_bwd(const B2RScene& sc, const Ctx& cx, const B2RBackwardArgs& a, float* gacc, cudaStream_t st) {
  // default: transposed-accumulation variant (composite_bwd2.cu); B2R_BWD_V1=1 selects the butterfly variant below
  static const bool use_v1 = getenv("B2R_BWD_V1") != nullptr;
  if (!use_v1) return launch_composite_bwd2(sc, cx, a, gacc, st);
  cudaMemsetAsync(gacc, 0, (size_t)(sc.P > 0 ? sc.P : 1) * 12 * sizeof(float), st);
  ProfScope p(K_COMPOSITE_BWD, st);
  if (a.dL_ddepth || a.dL_dalpha)
    composite_bwd_kernel<true><<<cx.tiles * 4, BWD_THREADS, 0, st>>>(sc, cx, a, gacc);
  else
    composite_bwd_kernel<false><<<cx.tiles * 4, BWD_THREADS, 0, st>>>(sc, cx, a, gacc);
  return check_launch();
}

}  // namespace b2r
