// composite_bwd3.cu -- K5 (default): backward alpha-composite (App. A.4), compacted-survivor variant.
//
// Same decomposition as composite_bwd2.cu (64-thread CTA per 8x8 quarter tile, longest list first, cp.async staging,
// 8x4 sub-tile culling, back to front, phase A lanes = pixels -> warp queue -> phase B lanes = splats).  ncu on
// composite_bwd2 (profiles/r01d) showed the kernel issue-bound at ~100 warp-instructions per (warp, splat) hit, of
// which ~20 were bit-walking of the survivor mask on the uniform datapath, ~12 the three-channel colour recurrence
// and ~8 bookkeeping executed by one lane.  This variant removes them:
//   * survivors of the 8x4 cull are COMPACTED once per 32-entry chunk: the lane that tested a splat copies its
//     48-byte record (plus its list position) to slot popc(mask above me) of a warp-private buffer, so the hit loop
//     walks consecutive shared-memory slots with a running pointer -- no per-hit find-first-set / index arithmetic;
//     an odd survivor count is padded with a record of opacity 0 (alpha = 0 is the identity of every recurrence);
//   * the colour recurrence of App. A.4 is carried as ONE scalar.  Only the dot product with the pixel's incoming
//     gradient is ever used:  B_k = accum_k . g  obeys the same recurrence  B = la * v_last + (1 - la) * B  with
//     v_k = c_k . g (+ z_k g_depth + g_alpha: depth and alpha are two more "colour channels", alpha's colour is 1);
//   * 1/(1 - alpha) is one MUFU.RCP (1 - alpha >= 0.01, no range fix-up needed);
//   * the Gaussian id rides in the record itself (geom.g2.w, written by project.cu), so nothing but the two
//     numbers (q, w) per pixel and the record copy per splat is written to the queue.
// Conventions (App. A.6): the 0.99 clamp is ignored on the way back; masks are constants.
#include <cstdlib>

#include "common.cuh"

namespace b2r {

constexpr int B3_THREADS = 64;
constexpr int B3_BATCH = 64;
constexpr int B3_QUEUE = 16;  // queued splats per warp before phase B runs (two lanes share a splat)

struct B3Stage {
  float4 a[B3_BATCH];  // px, py, A2, B2
  float4 b[B3_BATCH];  // C2, opacity, depth, thr2
  float4 c[B3_BATCH];  // r, g, b, id | clamp bits << 29
};
struct B3Compact {     // survivors of one 32-entry chunk, back-to-front; slot 32 = room for the odd-count pad
  float4 a[33];
  float4 b[33];        // .w = list position (int bits) instead of thr2
  float4 c[33];
};

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool HAS_DA>
__global__ void __launch_bounds__(B3_THREADS) composite_bwd3_kernel(const B2RScene sc, const Ctx cx,
                                                                    const B2RBackwardArgs args, float* __restrict__ gacc) {
  __shared__ B3Stage stage[2];
  __shared__ B3Compact compact[2];
  __shared__ float2 tb[2][B3_QUEUE][33];  // [warp][queued splat][pixel], padded rows: conflict-free both ways
  __shared__ float4 qm0[2][B3_QUEUE];     // px, py, A2, B2
  __shared__ float4 qm1[2][B3_QUEUE];     // C2, opacity, id bits, -
  __shared__ float4 gpix[2][32];          // per pixel of the warp: g_r, g_g, g_b, g_depth
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
  const float Tfb = T_final * (__ldg(sc.bg) * g_r + __ldg(sc.bg + 1) * g_g + __ldg(sc.bg + 2) * g_b);
  gpix[warp][lane] = make_float4(g_r, g_g, g_b, g_d);

  const int warp_n = __reduce_max_sync(0xffffffffu, my_n);
  if (lane == 0) warp_max_s[warp] = warp_n;
  __syncthreads();
  const int nmax = max(warp_max_s[0], warp_max_s[1]);
  if (nmax == 0) return;
  const int nb = (nmax + B3_BATCH - 1) / B3_BATCH;

  // blend state, walked back to front: T, and the scalar form of the "accumulated colour behind me" recurrence
  float T = T_final, B = 0.f, la = 0.f, olm = 1.f, lv = 0.f;
  int qpos = 0;  // warp-uniform
  B3Compact& cw = compact[warp];
  const bool lane0 = lane == 0;
  const unsigned lanes_above = 0xfffffffeu << lane;  // lanes with a higher index (= later list entries)

  // phase B: lanes l and l+16 share queued splat l, walk 16 pixels each and are combined with one shuffle per sum
  auto drain = [&](const int count) {
    constexpr int PIX = 16;
    __syncwarp();
    const int h = lane & (B3_QUEUE - 1), half = lane >> 4;
    const bool live = h < count;
    float Sx = 0.f, Sy = 0.f, Sxx = 0.f, Sxy = 0.f, Syy = 0.f, Sq = 0.f, Sr = 0.f, Sg = 0.f, Sb = 0.f, Sd = 0.f;
    float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = make_float4(0.f, 1.f, 0.f, 0.f);
    if (live) {
      m0 = qm0[warp][h];
      m1 = qm1[warp][h];
      const float mx = m0.x - rx0, my = m0.y - ry0;
#pragma unroll
      for (int k = 0; k < PIX; k++) {
        const int p = half * PIX + k;
        const float2 t = tb[warp][h][p];
        const float4 g = gpix[warp][p];
        const float dx = mx - (float)(p & 7), dy = my - (float)(p >> 3);
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
    if (live && half == 0) {
      // accumulator row convention of project_bwd.cu
      float* dst = gacc + (size_t)(__float_as_uint(m1.z) & 0x1fffffffu) * 12;
      red_add_v4(dst, 2.f * m0.z * Sx + m0.w * Sy, 2.f * m1.x * Sy + m0.w * Sx, Sxx, Sxy);
      red_add_v4(dst + 4, Syy, __fdividef(Sq, m1.y), Sd, 0.f);
      red_add_v4(dst + 8, Sr, Sg, Sb, 0.f);
    }
    __syncwarp();
  };

  auto issue = [&](int b) {
    B3Stage& s = stage[b & 1];
    const int idx = b * B3_BATCH + threadIdx.x;
    if (idx < nmax) {
      const uint32_t id = __ldg(ids + idx);
      const float4* src = reinterpret_cast<const float4*>(cx.geom + id);
      cp_async16(&s.a[threadIdx.x], src);
      cp_async16(&s.b[threadIdx.x], src + 1);
      cp_async16(&s.c[threadIdx.x], src + 2);
    }
    cp_async_commit();
  };

  // one replayed splat: a, bb (bb.w = list position), col are warp-uniform; everything else is per pixel
  auto replay = [&](const float4 a, const float4 bb, const float4 col, const float p2, const float araw,
                    const bool valid) {
    const float ae = valid ? fminf(K_ALPHA_MAX, araw) : 0.f;  // a skipped splat enters with alpha = 0 (identity)
    const float om = 1.f - ae;
    const float rcp = rcp_approx(om);
    const float Tn = T * rcp;
    float v = fmaf(col.z, g_b, fmaf(col.y, g_g, col.x * g_r));
    if (HAS_DA) v += fmaf(bb.z, g_d, g_a);
    B = fmaf(la, lv, olm * B);
    const float dLda = fmaf(v - B, Tn, -Tfb * rcp);
    la = ae; olm = om; lv = v; T = Tn;
    (void)p2;
    tb[warp][qpos][lane] = make_float2(valid ? araw * dLda : 0.f, ae * Tn);  // q = dL/dG * G (clamp ignored), w
    if (lane0) {
      qm0[warp][qpos] = a;
      qm1[warp][qpos] = make_float4(bb.x, bb.y, col.w, 0.f);
    }
    if (++qpos == B3_QUEUE) {
      drain(B3_QUEUE);
      qpos = 0;
    }
  };

  issue(nb - 1);
  for (int b = nb - 1; b >= 0; b--) {
    cp_async_wait<0>();
    __syncthreads();  // batch b staged; both warps are done with batch b+1
    if (b > 0) issue(b - 1);
    const int count = min(B3_BATCH, nmax - b * B3_BATCH);
    const B3Stage& s = stage[b & 1];
    if (warp_n <= b * B3_BATCH) continue;  // warp-uniform: none of my pixels reaches this batch
    for (int c0 = ((count - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
      const int idx = c0 + lane;
      const int pos = b * B3_BATCH + idx;
      bool hit = false;
      float4 a, bb;
      if (idx < count && pos < warp_n) {
        a = s.a[idx];
        bb = s.b[idx];
        hit = !(region_max_p2(a.x, a.y, a.z, a.w, bb.x, rx0, ry0, rx1, ry1) < bb.w);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask == 0u) continue;
      const int n = __popc(mask);
      if (hit) {  // back to front: the highest surviving list position goes to slot 0
        const int slot = __popc(mask & lanes_above);
        cw.a[slot] = a;
        cw.b[slot] = make_float4(bb.x, bb.y, bb.z, __int_as_float(pos));
        cw.c[slot] = s.c[idx];
      }
      if (lane0 && (n & 1)) {  // pad to an even count with a splat that can never be valid
        cw.a[n] = make_float4(0.f, 0.f, 0.f, 0.f);
        cw.b[n] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
        cw.c[n] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      // Two survivors per trip: their exponent evaluations (shared loads, MUFU) are independent and overlap; only the
      // short blend-state recurrence is serial.
      for (int k = 0; k < n; k += 2) {
        const float4 a0 = cw.a[k], b0 = cw.b[k], col0 = cw.c[k];
        const float4 a1 = cw.a[k + 1], b1 = cw.b[k + 1], col1 = cw.c[k + 1];
        const float dx0 = a0.x - pxf, dy0 = a0.y - pyf, dx1 = a1.x - pxf, dy1 = a1.y - pyf;
        const float p20 = a0.z * dx0 * dx0 + b0.x * dy0 * dy0 + a0.w * dx0 * dy0;
        const float p21 = a1.z * dx1 * dx1 + b1.x * dy1 * dy1 + a1.w * dx1 * dy1;
        const float ar0 = b0.y * ex2_approx(p20), ar1 = b1.y * ex2_approx(p21);
        const bool v0 = (__float_as_int(b0.w) < my_n) && (p20 <= 0.f) && (ar0 >= K_ALPHA_MIN);
        const bool v1 = (__float_as_int(b1.w) < my_n) && (p21 <= 0.f) && (ar1 >= K_ALPHA_MIN);
        if (__any_sync(0xffffffffu, v0)) replay(a0, b0, col0, p20, ar0, v0);
        if (__any_sync(0xffffffffu, v1)) replay(a1, b1, col1, p21, ar1, v1);
      }
      __syncwarp();  // the compact buffer is rewritten by the next chunk
    }
  }
  if (qpos > 0) drain(qpos);
  if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_bwd), (unsigned long long)nmax);
}

int launch_composite_bwd2(const B2RScene& sc, const Ctx& cx, const B2RBackwardArgs& a, float* gacc, cudaStream_t st);

int launch_composite_bwd(const B2RScene& sc, const Ctx& cx, const B2RBackwardArgs& a, float* gacc, cudaStream_t st) {
  // B2R_BWD_V2=1 selects the previous variant (composite_bwd2.cu) for A/B measurements
  static const bool use_v2 = getenv("B2R_BWD_V2") != nullptr;
  if (use_v2) return launch_composite_bwd2(sc, cx, a, gacc, st);
  cudaMemsetAsync(gacc, 0, (size_t)(sc.P > 0 ? sc.P : 1) * 12 * sizeof(float), st);
  ProfScope p(K_COMPOSITE_BWD, st);
  if (a.dL_ddepth || a.dL_dalpha)
    composite_bwd3_kernel<true><<<cx.tiles * 4, B3_THREADS, 0, st>>>(sc, cx, a, gacc);
  else
    composite_bwd3_kernel<false><<<cx.tiles * 4, B3_THREADS, 0, st>>>(sc, cx, a, gacc);
  return check_launch();
}

}  // namespace b2r
