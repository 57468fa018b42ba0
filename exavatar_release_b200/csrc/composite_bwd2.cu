// composite_bwd2.cu -- K5 (default variant): backward alpha-composite (App. A.4) with a transposed accumulation phase.
//
// Same work decomposition as composite_fwd.cu (one 64-thread CTA per 8x8 quarter tile, longest list first, cp.async
// staging, 8x4 sub-tile culling, back to front).  What differs from the first (butterfly-reduction) backward is how the 32 pixels of a warp
// are summed into per-splat gradients:
//
//   phase A (lanes = pixels)  for every surviving splat the warp replays the blend state of its 32 pixels and writes
//                             just two numbers per pixel into a warp-private shared-memory queue slot:
//                             q = dL/dG * G  and  w = alpha * T.  No cross-lane traffic.
//   phase B (lanes = splats)  when 32 splats are queued, lane l takes splat l and loops over the 32 pixels, rebuilding
//                             dx, dy from the pixel index and accumulating the nine sums in registers
//                             (q dx, q dy, q dx^2, q dx dy, q dy^2, q, w g_r, w g_g, w g_b).  No shuffles, no
//                             shared-memory accumulator, no per-batch flush: each lane leaves with three 16-byte
//                             vector reductions (REDG.E.ADD.F32x4) for its splat.
//
// Versus the butterfly variant this trades ~46 shuffle/select/add instructions per (warp, splat) for ~17 FMA-class
// instructions, keeps the queue across staging batches, and drops one CTA barrier per batch.
// Conventions (App. A.6): the 0.99 clamp is ignored on the way back; masks are constants.  A splat that is skipped
// at a pixel enters the recurrences with alpha = 0, which is the identity for every state variable.
#include "common.cuh"

namespace b2r {

constexpr int B2_THREADS = 64;
constexpr int B2_BATCH = 64;
constexpr int B2_PER_THREAD = B2_BATCH / B2_THREADS;
#ifndef B2_QUEUE_DEPTH
#define B2_QUEUE_DEPTH 16
#endif
constexpr int B2_QUEUE = B2_QUEUE_DEPTH;  // 32 or 16
#ifndef B2_DRAIN_UNROLL
#define B2_DRAIN_UNROLL 32
#endif

struct B2Stage {
  float4 a[B2_BATCH];
  float4 b[B2_BATCH];
  float4 c[B2_BATCH];
  uint32_t id[B2_BATCH];
};

template <bool HAS_DA>
__global__ void __launch_bounds__(B2_THREADS) composite_bwd2_kernel(const B2RScene sc, const Ctx cx,
                                                                    const B2RBackwardArgs args, float* __restrict__ gacc) {
  __shared__ B2Stage stage[2];
  __shared__ float2 tb[2][B2_QUEUE][33];        // [warp][queued splat][pixel], padded rows: conflict-free both ways
  __shared__ float4 qm0[2][B2_QUEUE];           // sx - wx0, sy - wy0, opacity, id bits
  __shared__ float4 qm1[2][B2_QUEUE];           // A2, B2, C2, -
  __shared__ float4 gpix[2][32];                // per pixel of the warp: g_r, g_g, g_b, g_depth
  __shared__ int warp_max_s[2];

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
  gpix[warp][lane] = make_float4(g_r, g_g, g_b, g_d);

  const int warp_n = __reduce_max_sync(0xffffffffu, my_n);
  if (lane == 0) warp_max_s[warp] = warp_n;
  __syncthreads();
  const int nmax = max(warp_max_s[0], warp_max_s[1]);
  if (nmax == 0) return;
  const int nb = (nmax + B2_BATCH - 1) / B2_BATCH;

  float T = T_final, last_alpha = 0.f;
  float acr = 0.f, acg = 0.f, acb = 0.f, lcr = 0.f, lcg = 0.f, lcb = 0.f;
  float acd = 0.f, lcd = 0.f, aca = 0.f;
  int qpos = 0;  // warp-uniform

  // phase B: with a 32-deep queue lane l owns queued splat l and walks all 32 pixels; with a 16-deep queue (half the
  // transposition buffer, more CTAs per SM) lanes l and l+16 share splat l, walk 16 pixels each and are combined
  // with one shuffle per sum
  auto drain = [&](const int count) {
    constexpr int HALVES = 32 / B2_QUEUE;
    constexpr int PIX = 32 / HALVES;
    __syncwarp();
    const int h = lane % B2_QUEUE, half = lane / B2_QUEUE;
    const bool live = h < count;
    float Sx = 0.f, Sy = 0.f, Sxx = 0.f, Sxy = 0.f, Syy = 0.f, Sq = 0.f, Sr = 0.f, Sg = 0.f, Sb = 0.f, Sd = 0.f;
    float4 m0 = make_float4(0.f, 0.f, 1.f, 0.f), m1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      m0 = qm0[warp][h];
      m1 = qm1[warp][h];
      constexpr int kUnroll = B2_DRAIN_UNROLL < PIX ? B2_DRAIN_UNROLL : PIX;
#pragma unroll kUnroll
      for (int k = 0; k < PIX; k++) {
        const int p = half * PIX + k;
        const float2 t = tb[warp][h][p];
        const float4 g = gpix[warp][p];
        const float dx = m0.x - (float)(p & 7), dy = m0.y - (float)(p >> 3);
        const float hx = t.x * dx, hy = t.x * dy;
        Sx += hx;
        Sy += hy;
        Sxx = fmaf(hx, dx, Sxx);
        Sxy = fmaf(hx, dy, Sxy);
        Syy = fmaf(hy, dy, Syy);
        Sq += t.x;
        Sr = fmaf(t.y, g.x, Sr);
        Sg = fmaf(t.y, g.y, Sg);
        Sb = fmaf(t.y, g.z, Sb);
        if (HAS_DA) Sd = fmaf(t.y, g.w, Sd);
      }
    }
    if (HALVES == 2) {
      Sx += __shfl_xor_sync(0xffffffffu, Sx, 16);
      Sy += __shfl_xor_sync(0xffffffffu, Sy, 16);
      Sxx += __shfl_xor_sync(0xffffffffu, Sxx, 16);
      Sxy += __shfl_xor_sync(0xffffffffu, Sxy, 16);
      Syy += __shfl_xor_sync(0xffffffffu, Syy, 16);
      Sq += __shfl_xor_sync(0xffffffffu, Sq, 16);
      Sr += __shfl_xor_sync(0xffffffffu, Sr, 16);
      Sg += __shfl_xor_sync(0xffffffffu, Sg, 16);
      Sb += __shfl_xor_sync(0xffffffffu, Sb, 16);
      if (HAS_DA) Sd += __shfl_xor_sync(0xffffffffu, Sd, 16);
    }
    if (live && half == 0) {
      // accumulator row convention of project_bwd.cu
      float* dst = gacc + (size_t)__float_as_uint(m0.w) * 12;
      red_add_v4(dst, 2.f * m1.x * Sx + m1.y * Sy, 2.f * m1.z * Sy + m1.y * Sx, Sxx, Sxy);
      red_add_v4(dst + 4, Syy, __fdividef(Sq, m0.z), Sd, 0.f);
      red_add_v4(dst + 8, Sr, Sg, Sb, 0.f);
    }
    __syncwarp();
  };

  auto issue = [&](int b) {
    B2Stage& s = stage[b & 1];
#pragma unroll
    for (int u = 0; u < B2_PER_THREAD; u++) {
      const int slot = threadIdx.x + u * B2_THREADS;
      const int idx = b * B2_BATCH + slot;
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
    __syncthreads();  // batch b staged; both warps are done with batch b+1
    if (b > 0) issue(b - 1);
    const int count = min(B2_BATCH, nmax - b * B2_BATCH);
    const B2Stage& s = stage[b & 1];
    if (warp_n > b * B2_BATCH) {
      for (int c0 = ((count - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
        const int idx = c0 + lane;
        bool hit = false;
        if (idx < count && b * B2_BATCH + idx < warp_n) {
          const float4 a = s.a[idx];
          const float4 bb = s.b[idx];
          hit = !(region_max_p2(a.x, a.y, a.z, a.w, bb.x, rx0, ry0, rx1, ry1) < bb.w);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        // Two survivors per trip: their exponent evaluations (shared loads, MUFU) are independent and overlap; only the
        // short blend-state recurrence is serial.  The longest lists bound this kernel by per-warp latency, not issue.
        while (mask) {
          const int k0 = 31 - __clz(mask);
          mask &= ~(1u << k0);
          const bool two = mask != 0u;
          const int k1 = two ? 31 - __clz(mask) : k0;
          if (two) mask &= ~(1u << k1);
          const int j0 = c0 + k0, j1 = c0 + k1;
          const float4 a0 = s.a[j0], b0 = s.b[j0], col0 = s.c[j0];
          const float4 a1 = s.a[j1], b1 = s.b[j1], col1 = s.c[j1];
          const float dx0 = a0.x - pxf, dy0 = a0.y - pyf, dx1 = a1.x - pxf, dy1 = a1.y - pyf;
          const float p20 = a0.z * dx0 * dx0 + b0.x * dy0 * dy0 + a0.w * dx0 * dy0;
          const float p21 = a1.z * dx1 * dx1 + b1.x * dy1 * dy1 + a1.w * dx1 * dy1;
          const float G0 = ex2_approx(p20), G1 = ex2_approx(p21);
          const float al0 = fminf(K_ALPHA_MAX, b0.y * G0), al1 = fminf(K_ALPHA_MAX, b1.y * G1);
          const bool v0 = (b * B2_BATCH + j0 < my_n) && (p20 <= 0.f) && (al0 >= K_ALPHA_MIN);
          const bool v1 = two && (b * B2_BATCH + j1 < my_n) && (p21 <= 0.f) && (al1 >= K_ALPHA_MIN);
          const bool any0 = __any_sync(0xffffffffu, v0), any1 = __any_sync(0xffffffffu, v1);
#pragma unroll
          for (int u = 0; u < 2; u++) {
          if (!(u == 0 ? any0 : any1)) continue;  // warp-uniform
          const float4 a = u == 0 ? a0 : a1, bb = u == 0 ? b0 : b1, col = u == 0 ? col0 : col1;
          const float G = u == 0 ? G0 : G1, alpha = u == 0 ? al0 : al1;
          const bool valid = u == 0 ? v0 : v1;
          const int j = u == 0 ? j0 : j1;
          // ---- phase A: branch-free state replay; a skipped splat enters with alpha = 0 (identity) ----
          const float ae = valid ? alpha : 0.f;
          const float Ge = valid ? G : 0.f;
          const float rcp = __fdividef(1.f, 1.f - ae);
          const float Tn = valid ? T * rcp : T;
          const float om = 1.f - last_alpha;
          acr = fmaf(last_alpha, lcr, om * acr);
          acg = fmaf(last_alpha, lcg, om * acg);
          acb = fmaf(last_alpha, lcb, om * acb);
          float dLda = (col.x - acr) * g_r + (col.y - acg) * g_g + (col.z - acb) * g_b;
          if (HAS_DA) {
            acd = fmaf(last_alpha, lcd, om * acd);
            aca = fmaf(om, aca, last_alpha);
            dLda += (bb.z - acd) * g_d + (1.f - aca) * g_a;
            lcd = bb.z;
          }
          lcr = col.x; lcg = col.y; lcb = col.z;
          last_alpha = ae;
          dLda = dLda * Tn - T_final * rcp * bg_dot;
          T = Tn;
          tb[warp][qpos][lane] = make_float2(bb.y * dLda * Ge, ae * Tn);  // q = dL/dG * G (clamp ignored), w
          if (lane == 0) {
            qm0[warp][qpos] = make_float4(a.x - rx0, a.y - ry0, bb.y, __uint_as_float(s.id[j]));
            qm1[warp][qpos] = make_float4(a.z, a.w, bb.x, 0.f);
          }
          if (++qpos == B2_QUEUE) {
            drain(B2_QUEUE);
            qpos = 0;
          }
          }  // u
        }
      }
    }
  }
  if (qpos > 0) drain(qpos);
  if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_bwd), (unsigned long long)nmax);
}

int launch_composite_bwd2(const B2RScene& sc, const Ctx& cx, const B2RBackwardArgs& a, float* gacc, cudaStream_t st) {
  if (!(a.flags & B2R_BWD_SCRATCH_ZEROED)) cudaMemsetAsync(gacc, 0, (size_t)(sc.P > 0 ? sc.P : 1) * 12 * sizeof(float), st);
  ProfScope p(K_COMPOSITE_BWD, st);
  if (a.dL_ddepth || a.dL_dalpha)
    launch_k(composite_bwd2_kernel<true>, cx.tiles * 4, B2_THREADS, 0, st, false, sc, cx, a, gacc);
  else
    launch_k(composite_bwd2_kernel<false>, cx.tiles * 4, B2_THREADS, 0, st, false, sc, cx, a, gacc);
  return check_launch();
}

}  // namespace b2r
