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
#ifndef B3_QUEUE_DEPTH
#define B3_QUEUE_DEPTH 16
#endif
#ifndef B3_MIN_BLOCKS
#define B3_MIN_BLOCKS 10  // 96 registers.  Measured on C2 / C4 (frames/s with 4 frames in flight; solo kernel us):
#endif                    // depth 16 @ 96 regs 6737 / 105 (kept); 16 @ 80 regs 6728 / 105; 8 @ 72-80 regs 6400 / 121-127;
                          // 32 @ 128 regs 6494 / 97 -- the deeper queue is faster alone but costs throughput in flight
constexpr int B3_QUEUE = B3_QUEUE_DEPTH;  // queued splats per warp before phase B runs (32 / B3_QUEUE lanes share a splat)

struct B3Stage {
  float4 a[B3_BATCH];  // px, py, A2, B2
  float4 b[B3_BATCH];  // C2, opacity, depth, thr2
  float4 c[B3_BATCH];  // r, g, b, id | clamp bits << 29
};
constexpr int B3_GROUP = 4;   // splats replayed per trip of the hit loop (their evaluations overlap: ILP 4)
constexpr int B3_CQ = 36;     // survivor queue: <= 3 left over + 32 new per chunk (+ pad)
struct B3Compact {            // warp-private queue of cull survivors, back to front
  float4 r[3][B3_CQ];         // [0] px,py,A2,B2  [1] C2,opacity,depth,list position (int bits)  [2] r,g,b,id bits
};
static_assert(B3_QUEUE % B3_GROUP == 0, "phase B runs when whole groups fill the transposition queue");

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool HAS_DA>
__global__ void __launch_bounds__(B3_THREADS, B3_MIN_BLOCKS) composite_bwd3_kernel(const B2RScene sc, const Ctx cx,
                                                                    const B2RBackwardArgs args, float* __restrict__ gacc) {
  __shared__ B3Stage stage[2];
  __shared__ B3Compact compact[2];
  __shared__ float2 tb[2][B3_QUEUE][33];  // [warp][queued splat][pixel], padded rows: conflict-free both ways
  __shared__ float4 qm0[2][B3_QUEUE];     // px, py, A2, B2
  __shared__ float4 qm1[2][B3_QUEUE];     // C2, opacity, id bits, -
  __shared__ float4 gpix[2][32];          // per pixel of the warp: g_r, g_g, g_b, g_depth
  __shared__ int warp_max_s[2];
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
  if (nmax == 0) {
    B2R_TRACE_END(0);
    return;
  }
  const int nb = (nmax + B3_BATCH - 1) / B3_BATCH;

  // blend state, walked back to front: T, and the scalar form of the "accumulated colour behind me" recurrence
  float T = T_final, B = 0.f, la = 0.f, olm = 1.f, lv = 0.f;
  int qpos = 0;  // warp-uniform
  B3Compact& cw = compact[warp];
  const bool lane0 = lane == 0;
  const unsigned lanes_above = 0xfffffffeu << lane;  // lanes with a higher index (= later list entries)

  // phase B: 32 / B3_QUEUE lanes share a queued splat, each walks its pixel rows; combined with shuffles
  auto drain = [&](const int count) {
    constexpr int SHARE = 32 / B3_QUEUE;  // lanes per queued splat
    constexpr int ROWS = 4 / SHARE;       // pixel rows (of eight) per lane
    static_assert(B3_QUEUE == 8 || B3_QUEUE == 16 || B3_QUEUE == 32, "queue depth");
    __syncwarp();
    const int h = lane & (B3_QUEUE - 1), part = lane / B3_QUEUE;
    const bool live = h < count;
    float Sx = 0.f, Sy = 0.f, Sxx = 0.f, Sxy = 0.f, Syy = 0.f, Sq = 0.f, Sr = 0.f, Sg = 0.f, Sb = 0.f, Sd = 0.f;
    float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = make_float4(0.f, 1.f, 0.f, 0.f);
    if (live) {
      m0 = qm0[warp][h];
      m1 = qm1[warp][h];
      // ROWS pixel rows of eight per lane.  Within a row dy is constant, so only q, q dx and q dx^2 are summed per
      // pixel; the dy moments are formed once per row from the row sums.
      const float mx = m0.x - rx0, my = m0.y - ry0;
      float dxs[8];
#pragma unroll
      for (int c = 0; c < 8; c++) dxs[c] = mx - (float)c;
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        const float dy = my - (float)(part * ROWS + r);
        float Rq = 0.f, Rx = 0.f, Rxx = 0.f;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int p = (part * ROWS + r) * 8 + c;
          const float2 t = tb[warp][h][p];
          const float4 g = gpix[warp][p];
          const float hx = t.x * dxs[c];
          Rq += t.x;
          Rx += hx;
          Rxx = fmaf(hx, dxs[c], Rxx);
          Sr = fmaf(t.y, g.x, Sr);
          Sg = fmaf(t.y, g.y, Sg);
          Sb = fmaf(t.y, g.z, Sb);
          if (HAS_DA) Sd = fmaf(t.y, g.w, Sd);
        }
        const float hy = Rq * dy;
        Sq += Rq;
        Sx += Rx;
        Sxx += Rxx;
        Sy += hy;
        Syy = fmaf(hy, dy, Syy);
        Sxy = fmaf(Rx, dy, Sxy);
      }
    }
#pragma unroll
    for (int o = B3_QUEUE; o < 32; o <<= 1) {  // combine the lanes that share a splat
      Sx += __shfl_xor_sync(0xffffffffu, Sx, o);
      Sy += __shfl_xor_sync(0xffffffffu, Sy, o);
      Sxx += __shfl_xor_sync(0xffffffffu, Sxx, o);
      Sxy += __shfl_xor_sync(0xffffffffu, Sxy, o);
      Syy += __shfl_xor_sync(0xffffffffu, Syy, o);
      Sq += __shfl_xor_sync(0xffffffffu, Sq, o);
      Sr += __shfl_xor_sync(0xffffffffu, Sr, o);
      Sg += __shfl_xor_sync(0xffffffffu, Sg, o);
      Sb += __shfl_xor_sync(0xffffffffu, Sb, o);
      if (HAS_DA) Sd += __shfl_xor_sync(0xffffffffu, Sd, o);
    }
    if (live && part == 0) {
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

  // One trip = B3_GROUP queued survivors.  Stage 1 (independent per splat, so the four overlap): exponent, alpha,
  // validity, the scalar "colour" v = c . g, and the per-splat record for phase B (written by lane 0).  Stage 2: the
  // short serial recurrences (T, B) and the two numbers per pixel that go to the transposition queue.  No branches
  // inside a trip; a splat no pixel of the warp accepts still takes a queue slot (all-zero column).
  auto replay_group = [&](const int k) {
    float araw[B3_GROUP], vv[B3_GROUP];
    bool valid[B3_GROUP];
#pragma unroll
    for (int u = 0; u < B3_GROUP; u++) {
      const float4 a = cw.r[0][k + u], bb = cw.r[1][k + u], col = cw.r[2][k + u];
      const float dx = a.x - pxf, dy = a.y - pyf;
      const float p2 = a.z * dx * dx + bb.x * dy * dy + a.w * dx * dy;
      araw[u] = bb.y * ex2_approx(p2);
      valid[u] = (__float_as_int(bb.w) < my_n) && (p2 <= 0.f) && (araw[u] >= K_ALPHA_MIN);
      float v = fmaf(col.z, g_b, fmaf(col.y, g_g, col.x * g_r));
      if (HAS_DA) v += fmaf(bb.z, g_d, g_a);
      vv[u] = v;
      if (lane0) {
        qm0[warp][qpos + u] = a;
        qm1[warp][qpos + u] = make_float4(bb.x, bb.y, col.w, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < B3_GROUP; u++) {
      const float ae = valid[u] ? fminf(K_ALPHA_MAX, araw[u]) : 0.f;  // a skipped splat enters with alpha = 0 (identity)
      const float om = 1.f - ae;
      const float rcp = rcp_approx(om);
      const float Tn = T * rcp;
      B = fmaf(la, lv, olm * B);
      const float dLda = fmaf(vv[u] - B, Tn, -Tfb * rcp);
      la = ae; olm = om; lv = vv[u]; T = Tn;
      tb[warp][qpos + u][lane] = make_float2(valid[u] ? araw[u] * dLda : 0.f, ae * Tn);  // q = dL/dG * G (clamp ignored), w
    }
    qpos += B3_GROUP;
    if (qpos == B3_QUEUE) {
      drain(B3_QUEUE);
      qpos = 0;
    }
  };

  int fill = 0;  // warp-uniform: survivors waiting in the queue (< B3_GROUP between chunks)
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
      if (hit) {  // back to front: the highest surviving list position is queued first
        const int slot = fill + __popc(mask & lanes_above);
        cw.r[0][slot] = a;
        cw.r[1][slot] = make_float4(bb.x, bb.y, bb.z, __int_as_float(pos));
        cw.r[2][slot] = s.c[idx];
      }
      fill += __popc(mask);
      __syncwarp();
      int k = 0;
      for (; k + B3_GROUP <= fill; k += B3_GROUP) replay_group(k);
      const int left = fill - k;
      __syncwarp();
      if (k > 0 && lane < 3 * left) {  // move the <= 3 leftover records to the front (sources are slots >= 4)
        const int t = (lane >= left) + (lane >= 2 * left), j = lane - t * left;
        cw.r[t][j] = cw.r[t][k + j];
      }
      fill = left;
      __syncwarp();  // queue reads / moves before the next append
    }
  }
  if (fill > 0) {  // flush: pad the last group with splats that can never be valid (list position INT_MAX); their
                   // opacity is 1 because phase B divides by it
    if (lane >= fill && lane < B3_GROUP) {
      cw.r[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      cw.r[1][lane] = make_float4(0.f, 1.f, 0.f, __int_as_float(0x7fffffff));
      cw.r[2][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    replay_group(0);
  }
  if (qpos > 0) drain(qpos);
  if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&cx.status->consumed_bwd), (unsigned long long)nmax);
  B2R_TRACE_END(nmax);
}

int launch_composite_bwd(const B2RScene& sc, const Ctx& cx, const B2RBackwardArgs& a, float* gacc, cudaStream_t st) {
  if (!(a.flags & B2R_BWD_SCRATCH_ZEROED)) cudaMemsetAsync(gacc, 0, (size_t)(sc.P > 0 ? sc.P : 1) * 12 * sizeof(float), st);
  ProfScope p(K_COMPOSITE_BWD, st);
  if (a.dL_ddepth || a.dL_dalpha)
    launch_k(composite_bwd3_kernel<true>, cx.tiles * 4, B3_THREADS, 0, st, false, sc, cx, a, gacc);
  else
    launch_k(composite_bwd3_kernel<false>, cx.tiles * 4, B3_THREADS, 0, st, false, sc, cx, a, gacc);
  return check_launch();
}

}  // namespace b2r

#ifdef B2R_CTA_TRACE
extern "C" int b2r_debug_trace_bwd(unsigned long long* buf) {
  return (int)cudaMemcpyToSymbol(b2r::g_cta_trace, &buf, sizeof(buf));
}
#endif
