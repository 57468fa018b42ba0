// composite_bwd.cu -- K5: backward alpha-composite (App. A.4), one CTA per 16x16 tile.
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
#include "common.cuh"

namespace b2r {

constexpr int BWD_BATCH = 256;

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
__global__ void __launch_bounds__(256) composite_bwd_kernel(const B2RScene sc, const Ctx cx, const B2RBackwardArgs args,
                                                            float* __restrict__ gacc) {
  constexpr int NV = HAS_DA ? 10 : 9;
  __shared__ BwdStage stage[2];
  __shared__ __align__(16) float acc[BWD_BATCH][12];
  __shared__ int warp_max_s[8];

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
  for (int i = threadIdx.x; i < BWD_BATCH * 12; i += blockDim.x) (&acc[0][0])[i] = 0.f;
  __syncthreads();
  int nmax = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) nmax = max(nmax, warp_max_s[i]);
  if (nmax == 0) return;
  const int nb = (nmax + BWD_BATCH - 1) / BWD_BATCH;

  const int my_slot = halving_slot<NV>(lane);
  // value index -> accumulator column: mx my ca cb cc op | r g b | d
  const int my_col = my_slot < 0 ? -1 : (my_slot < 6 ? my_slot : (my_slot < 9 ? my_slot + 2 : 6));

  float T = T_final, last_alpha = 0.f;
  float acr = 0.f, acg = 0.f, acb = 0.f, lcr = 0.f, lcg = 0.f, lcb = 0.f;
  float acd = 0.f, lcd = 0.f, aca = 0.f;

  auto issue = [&](int b) {
    const int idx = b * BWD_BATCH + threadIdx.x;
    if (idx < nmax) {
      const uint32_t id = __ldg(ids + idx);
      const float4* src = reinterpret_cast<const float4*>(cx.geom + id);
      BwdStage& s = stage[b & 1];
      cp_async16(&s.a[threadIdx.x], src);
      cp_async16(&s.b[threadIdx.x], src + 1);
      cp_async16(&s.c[threadIdx.x], src + 2);
      s.id[threadIdx.x] = id;
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
          float v[NV];
#pragma unroll
          for (int i = 0; i < NV; i++) v[i] = 0.f;
          if (valid) {
            const float4 col = s.c[j];
            const float rcp = __fdividef(1.f, 1.f - alpha);
            T *= rcp;
            const float w = alpha * T;
            float dLda = 0.f;
            acr = last_alpha * lcr + (1.f - last_alpha) * acr; lcr = col.x; dLda += (col.x - acr) * g_r;
            acg = last_alpha * lcg + (1.f - last_alpha) * acg; lcg = col.y; dLda += (col.y - acg) * g_g;
            acb = last_alpha * lcb + (1.f - last_alpha) * acb; lcb = col.z; dLda += (col.z - acb) * g_b;
            if (HAS_DA) {
              acd = last_alpha * lcd + (1.f - last_alpha) * acd; lcd = bb.z; dLda += (bb.z - acd) * g_d;
              aca = last_alpha + (1.f - last_alpha) * aca; dLda += (1.f - aca) * g_a;
            }
            dLda *= T;
            last_alpha = alpha;
            dLda += (-T_final * rcp) * bg_dot;
            const float dLdG = bb.y * dLda;  // clamp ignored (App. A.6 i)
            const float gdx = G * dx, gdy = G * dy;
            v[0] = dLdG * (2.f * gdx * a.z + gdy * a.w);
            v[1] = dLdG * (2.f * gdy * bb.x + gdx * a.w);
            const float hx = dLdG * gdx, hy = dLdG * gdy;
            v[2] = hx * dx;
            v[3] = hx * dy;
            v[4] = hy * dy;
            v[5] = G * dLda;
            v[6] = w * g_r;
            v[7] = w * g_g;
            v[8] = w * g_b;
            if (HAS_DA) v[NV - 1] = w * g_d;
          }
          halving_reduce<NV, 16>(v, lane);
          if (my_col >= 0) atomicAdd(&acc[j][my_col], v[0]);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < count) {  // flush this batch: one row per thread
      float4* row = reinterpret_cast<float4*>(&acc[threadIdx.x][0]);
      const float4 q0 = row[0], q1 = row[1], q2 = row[2];
      const bool any = (q0.x != 0.f) | (q0.y != 0.f) | (q0.z != 0.f) | (q0.w != 0.f) | (q1.x != 0.f) | (q1.y != 0.f) |
                       (q1.z != 0.f) | (q2.x != 0.f) | (q2.y != 0.f) | (q2.z != 0.f);
      if (any) {
        float* dst = gacc + (size_t)s.id[threadIdx.x] * 12;
        red_add_v4(dst, q0.x, q0.y, q0.z, q0.w);
        red_add_v4(dst + 4, q1.x, q1.y, q1.z, 0.f);
        red_add_v4(dst + 8, q2.x, q2.y, q2.z, 0.f);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        row[0] = z; row[1] = z; row[2] = z;
      }
    }
  }
  if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_bwd), (unsigned long long)nmax);
}

int launch_composite_bwd(const B2RScene& sc, const Ctx& cx, const B2RBackwardArgs& a, float* gacc, cudaStream_t st) {
  cudaMemsetAsync(gacc, 0, (size_t)(sc.P > 0 ? sc.P : 1) * 12 * sizeof(float), st);
  if (a.dL_ddepth || a.dL_dalpha)
    composite_bwd_kernel<true><<<cx.tiles, 256, 0, st>>>(sc, cx, a, gacc);
  else
    composite_bwd_kernel<false><<<cx.tiles, 256, 0, st>>>(sc, cx, a, gacc);
  return check_launch();
}

}  // namespace b2r
