/*
 * gs_oracle.c -- CPU oracle for ExAvatar's Gaussian-rasterisation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (exavatar_release_b200/)
 * may import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py use it, as the checker or
 * as the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference repo (mks0601/ExAvatar_RELEASE) does not vendor
 * the rasteriser it calls (`diff_gaussian_rasterization_depth`,
 * avatar/common/nets/module.py:11; environment.yml:272 pins
 * `diff-gaussian-rasterization==0.0.0` without a commit) and holds no tests or
 * golden vectors for it.  This file restates the *published* algorithm of that
 * package family (graphdeco-inria diff-gaussian-rasterization + the depth/alpha
 * forks) as summarised in SURVEY.md Appendix A, and is anchored on the
 * reference's own call site (module.py:592-647) and camera conventions
 * (avatar/common/utils/transforms.py:38-70).  It is validated by closed-form
 * known-answer tests and by an independent fp64 autograd
 * restatement (oracle/dense_autograd.py).
 *
 * Each function cites the SURVEY appendix section it follows; the reference
 * file:line it serves is given where one exists.
 *
 * Build: see oracle/build.py (gcc -O2 -fopenmp -ffp-contract=off; fp32 and,
 * with -DGSO_FP64, fp64 variants).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef GSO_FP64
typedef double real;
#define R_EXP exp
#define R_SQRT sqrt
#define R_CEIL ceil
#else
typedef float real;
#define R_EXP expf
#define R_SQRT sqrtf
#define R_CEIL ceilf
#endif

/* Constants of SURVEY App. A (float-rounded in both precisions so that the
 * discrete decisions sit at the same thresholds). */
#define TILE 16
#define K_NEAR ((real)0.2f)
#define K_DILATE ((real)0.3f)
#define K_ALPHA_MAX ((real)0.99f)
#define K_ALPHA_MIN ((real)(1.0f / 255.0f))
#define K_T_MIN ((real)0.0001f)
#define K_FRUSTUM ((real)1.3f)
#define K_EPS_W ((real)0.0000001f)
#define K_EIG_FLOOR ((real)0.1f)

#define SH_C0 ((real)0.28209479177387814f)
#define SH_C1 ((real)0.4886025119029199f)
static const real SH_C2[5] = {(real)1.0925484305920792f, (real)-1.0925484305920792f, (real)0.31539156525252005f,
                              (real)-1.0925484305920792f, (real)0.5462742152960396f};
static const real SH_C3[7] = {(real)-0.5900435899266435f, (real)2.890611442640554f, (real)-0.4570457994644658f,
                              (real)0.3731763325901154f,  (real)-0.4570457994644658f, (real)1.445305721320277f,
                              (real)-0.5900435899266435f};

typedef struct {
  uint32_t tile;
  uint32_t id;
  real depth;
} dup_t;

typedef struct gso_ctx {
  int P, W, H, gx, gy, M, deg;
  int has_sh, has_cov_precomp;
  real scale_modifier, tanfovx, tanfovy;
  real view[16], proj[16], campos[3], bg[3];
  /* copies of the inputs (backward needs them) */
  real *means3D, *shs, *opac, *scales, *rots, *cov_pre;
  /* per-Gaussian state (App. A.1) */
  int32_t* radii;
  uint32_t* tiles_touched;
  int32_t* rect; /* 4 per Gaussian: xmin ymin xmax ymax */
  real *xy, *depth, *conic_o, *cov3D, *rgb;
  uint8_t* clamped; /* 3 per Gaussian */
  /* binning (App. A.2) */
  int64_t D;
  uint32_t* list;  /* sorted ids, length D */
  int64_t* ranges; /* 2 per tile */
  /* per pixel (App. A.3) */
  real* final_T;
  uint32_t* n_contrib;
  uint32_t* n_walked; /* list entries a pixel iterated over before it stopped */
  real* out_color;    /* 3*H*W, includes background */
  int64_t consumed_fwd, consumed_bwd;
} gso_ctx;

static real* dup_real(const real* src, size_t n) {
  if (!src || n == 0) return NULL;
  real* p = (real*)malloc(n * sizeof(real));
  memcpy(p, src, n * sizeof(real));
  return p;
}

/* column-major 4x4, element (r,c) at m[4*c+r]: the layout GaussianRenderer hands over after
 * its .permute(1,0) (module.py:605-607). */
static inline void xform4x3(const real* p, const real* m, real* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static inline void xform4x4(const real* p, const real* m, real* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
  o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* Rotation of an UN-normalised quaternion (r,x,y,z): R_std (App. A.1 step 3). */
static inline void quat_to_R(const real* q, real R[3][3]) {
  real r = q[0], x = q[1], y = q[2], z = q[3];
  R[0][0] = (real)1 - (real)2 * (y * y + z * z);
  R[0][1] = (real)2 * (x * y - r * z);
  R[0][2] = (real)2 * (x * z + r * y);
  R[1][0] = (real)2 * (x * y + r * z);
  R[1][1] = (real)1 - (real)2 * (x * x + z * z);
  R[1][2] = (real)2 * (y * z - r * x);
  R[2][0] = (real)2 * (x * z - r * y);
  R[2][1] = (real)2 * (y * z + r * x);
  R[2][2] = (real)1 - (real)2 * (x * x + y * y);
}

/* App. A.1 step 3: Sigma = R S^2 R^T, six upper-triangular floats. */
static void cov3d_from_scale_rot(const real* scale, real mod, const real* q, real* c6) {
  real R[3][3];
  quat_to_R(q, R);
  real s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
  /* M = S * R_std^T  (row i scaled by s_i); Sigma = M^T M */
  real Mm[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Mm[i][j] = s[i] * R[j][i];
  real S[3][3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) S[a][b] = Mm[0][a] * Mm[0][b] + Mm[1][a] * Mm[1][b] + Mm[2][a] * Mm[2][b];
  c6[0] = S[0][0];
  c6[1] = S[0][1];
  c6[2] = S[0][2];
  c6[3] = S[1][1];
  c6[4] = S[1][2];
  c6[5] = S[2][2];
}

/* App. A.1 step 4.  Returns A = J*Rv (2x3), the clamped t and grad multipliers; cov2 = (a,b,c) with dilation. */
typedef struct {
  real t[3];
  real A[2][3];
  real xmul, ymul;
  real a, b, c;
} ewa_t;

static void ewa_project(const real* mean, const real* c6, const gso_ctx* g, real fx, real fy, ewa_t* e) {
  real t[3];
  xform4x3(mean, g->view, t);
  real limx = K_FRUSTUM * g->tanfovx, limy = K_FRUSTUM * g->tanfovy;
  real txtz = t[0] / t[2], tytz = t[1] / t[2];
  e->xmul = (txtz < -limx || txtz > limx) ? (real)0 : (real)1;
  e->ymul = (tytz < -limy || tytz > limy) ? (real)0 : (real)1;
  real cx = txtz < -limx ? -limx : (txtz > limx ? limx : txtz);
  real cy = tytz < -limy ? -limy : (tytz > limy ? limy : tytz);
  t[0] = cx * t[2];
  t[1] = cy * t[2];
  e->t[0] = t[0];
  e->t[1] = t[1];
  e->t[2] = t[2];
  real J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
  real J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
  const real* v = g->view;
  /* Rv[m][k] = v[4k+m] */
  for (int k = 0; k < 3; k++) {
    e->A[0][k] = J00 * v[4 * k + 0] + J02 * v[4 * k + 2];
    e->A[1][k] = J11 * v[4 * k + 1] + J12 * v[4 * k + 2];
  }
  real S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
  real B[2][3];
  for (int r = 0; r < 2; r++)
    for (int k = 0; k < 3; k++) B[r][k] = e->A[r][0] * S[0][k] + e->A[r][1] * S[1][k] + e->A[r][2] * S[2][k];
  e->a = B[0][0] * e->A[0][0] + B[0][1] * e->A[0][1] + B[0][2] * e->A[0][2] + K_DILATE;
  e->b = B[0][0] * e->A[1][0] + B[0][1] * e->A[1][1] + B[0][2] * e->A[1][2];
  e->c = B[1][0] * e->A[1][0] + B[1][1] * e->A[1][1] + B[1][2] * e->A[1][2] + K_DILATE;
}

/* App. A.7 / transforms.py:112-167 polynomial, layout (P,M,3), + 0.5, clamp at 0 (module.py:265-266). */
static void sh_to_rgb(int deg, int M, const real* sh, const real* mean, const real* campos, real* rgb, uint8_t* clamped) {
  real d[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
  real n = R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  real x = d[0] / n, y = d[1] / n, z = d[2] / n;
  (void)M;
  for (int c = 0; c < 3; c++) {
#define SH(k) sh[(k) * 3 + c]
    real r = SH_C0 * SH(0);
    if (deg > 0) {
      r = r - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
      if (deg > 1) {
        real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) + SH_C2[2] * ((real)2 * zz - xx - yy) * SH(6) +
            SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
        if (deg > 2) {
          r = r + SH_C3[0] * y * ((real)3 * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
              SH_C3[2] * y * ((real)4 * zz - xx - yy) * SH(11) +
              SH_C3[3] * z * ((real)2 * zz - (real)3 * xx - (real)3 * yy) * SH(12) +
              SH_C3[4] * x * ((real)4 * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
              SH_C3[6] * x * (xx - (real)3 * yy) * SH(15);
        }
      }
    }
#undef SH
    r += (real)0.5;
    clamped[c] = r < 0;
    rgb[c] = r < 0 ? (real)0 : r;
  }
}

static int cmp_dup(const void* pa, const void* pb) {
  const dup_t* a = (const dup_t*)pa;
  const dup_t* b = (const dup_t*)pb;
  if (a->tile != b->tile) return a->tile < b->tile ? -1 : 1;
  if (a->depth != b->depth) return a->depth < b->depth ? -1 : 1;
  if (a->id != b->id) return a->id < b->id ? -1 : 1; /* stable sort of index-ordered input (App. A.2) */
  return 0;
}

void gso_free(gso_ctx* g) {
  if (!g) return;
  free(g->means3D); free(g->shs); free(g->opac); free(g->scales); free(g->rots); free(g->cov_pre);
  free(g->radii); free(g->tiles_touched); free(g->rect); free(g->xy); free(g->depth); free(g->conic_o);
  free(g->cov3D); free(g->rgb); free(g->clamped); free(g->list); free(g->ranges); free(g->final_T);
  free(g->n_contrib); free(g->n_walked); free(g->out_color);
  free(g);
}

/*
 * Forward: App. A.1 (preprocess) -> A.2 (binning) -> A.3 (composite).
 * Serves GaussianRasterizer.forward as called at module.py:632-640.
 * shs: (P,M,3) or NULL; colors_precomp: (P,3) or NULL; cov3D_precomp (P,6) or NULL.
 * Outputs: out_color (3,H,W), out_depth (H,W), out_alpha (H,W), radii (P).
 */
gso_ctx* gso_forward(int P, int W, int H, int sh_degree, int M, const real* means3D, const real* shs,
                     const real* colors_precomp, const real* opacities, const real* scales, const real* rotations,
                     const real* cov3D_precomp, real scale_modifier, const real* viewmatrix, const real* projmatrix,
                     const real* campos, real tanfovx, real tanfovy, const real* bg, real* out_color, real* out_depth,
                     real* out_alpha, int32_t* out_radii) {
  gso_ctx* g = (gso_ctx*)calloc(1, sizeof(gso_ctx));
  g->P = P; g->W = W; g->H = H; g->deg = sh_degree; g->M = M;
  g->gx = (W + TILE - 1) / TILE;
  g->gy = (H + TILE - 1) / TILE;
  g->has_sh = shs != NULL;
  g->has_cov_precomp = cov3D_precomp != NULL;
  g->scale_modifier = scale_modifier; g->tanfovx = tanfovx; g->tanfovy = tanfovy;
  memcpy(g->view, viewmatrix, sizeof(g->view));
  memcpy(g->proj, projmatrix, sizeof(g->proj));
  memcpy(g->campos, campos, sizeof(g->campos));
  memcpy(g->bg, bg, sizeof(g->bg));
  g->means3D = dup_real(means3D, (size_t)P * 3);
  g->shs = dup_real(shs, (size_t)P * M * 3);
  g->opac = dup_real(opacities, (size_t)P);
  g->scales = dup_real(scales, (size_t)P * 3);
  g->rots = dup_real(rotations, (size_t)P * 4);
  g->cov_pre = dup_real(cov3D_precomp, (size_t)P * 6);
  size_t Pn = P > 0 ? (size_t)P : 1;
  g->radii = (int32_t*)calloc(Pn, sizeof(int32_t));
  g->tiles_touched = (uint32_t*)calloc(Pn, sizeof(uint32_t));
  g->rect = (int32_t*)calloc(Pn * 4, sizeof(int32_t));
  g->xy = (real*)calloc(Pn * 2, sizeof(real));
  g->depth = (real*)calloc(Pn, sizeof(real));
  g->conic_o = (real*)calloc(Pn * 4, sizeof(real));
  g->cov3D = (real*)calloc(Pn * 6, sizeof(real));
  g->rgb = (real*)calloc(Pn * 3, sizeof(real));
  g->clamped = (uint8_t*)calloc(Pn * 3, 1);
  const real fx = (real)W / ((real)2 * tanfovx), fy = (real)H / ((real)2 * tanfovy);
  const int gx = g->gx, gy = g->gy;

  /* ---- App. A.1 ---- */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    const real* p = means3D + 3 * (size_t)i;
    real pv[3];
    xform4x3(p, g->view, pv);
    if (pv[2] <= K_NEAR) continue; /* step 1 */
    real ph[4];
    xform4x4(p, g->proj, ph); /* step 2 */
    real pw = (real)1 / (ph[3] + K_EPS_W);
    real pproj[2] = {ph[0] * pw, ph[1] * pw};
    real* c6 = g->cov3D + 6 * (size_t)i;
    if (cov3D_precomp)
      memcpy(c6, cov3D_precomp + 6 * (size_t)i, 6 * sizeof(real));
    else
      cov3d_from_scale_rot(scales + 3 * (size_t)i, scale_modifier, rotations + 4 * (size_t)i, c6); /* step 3 */
    ewa_t e;
    ewa_project(p, c6, g, fx, fy, &e); /* step 4 */
    real det = e.a * e.c - e.b * e.b; /* step 5 */
    if (det == (real)0) continue;
    real det_inv = (real)1 / det;
    real conic[3] = {e.c * det_inv, -e.b * det_inv, e.a * det_inv};
    real mid = (real)0.5 * (e.a + e.c); /* step 6 */
    real disc = mid * mid - det;
    if (disc < K_EIG_FLOOR) disc = K_EIG_FLOOR;
    real l1 = mid + R_SQRT(disc), l2 = mid - R_SQRT(disc);
    real lm = l1 > l2 ? l1 : l2;
    real radf = R_CEIL((real)3 * R_SQRT(lm));
    int radius = (int)radf;
    real px = ((pproj[0] + (real)1) * (real)W - (real)1) * (real)0.5; /* step 7 */
    real py = ((pproj[1] + (real)1) * (real)H - (real)1) * (real)0.5;
    /* step 8: C (int) truncation, clamp to grid */
    int x0 = (int)((px - (real)radius) / (real)TILE), y0 = (int)((py - (real)radius) / (real)TILE);
    int x1 = (int)((px + (real)radius + (real)(TILE - 1)) / (real)TILE);
    int y1 = (int)((py + (real)radius + (real)(TILE - 1)) / (real)TILE);
    x0 = x0 < 0 ? 0 : (x0 > gx ? gx : x0);
    y0 = y0 < 0 ? 0 : (y0 > gy ? gy : y0);
    x1 = x1 < 0 ? 0 : (x1 > gx ? gx : x1);
    y1 = y1 < 0 ? 0 : (y1 > gy ? gy : y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;
    if (shs) /* step 9 */
      sh_to_rgb(sh_degree, M, shs + (size_t)i * M * 3, p, g->campos, g->rgb + 3 * (size_t)i, g->clamped + 3 * (size_t)i);
    else
      memcpy(g->rgb + 3 * (size_t)i, colors_precomp + 3 * (size_t)i, 3 * sizeof(real));
    g->depth[i] = pv[2]; /* step 10 */
    g->radii[i] = radius;
    g->xy[2 * i] = px;
    g->xy[2 * i + 1] = py;
    g->conic_o[4 * i] = conic[0];
    g->conic_o[4 * i + 1] = conic[1];
    g->conic_o[4 * i + 2] = conic[2];
    g->conic_o[4 * i + 3] = opacities[i];
    g->rect[4 * i] = x0; g->rect[4 * i + 1] = y0; g->rect[4 * i + 2] = x1; g->rect[4 * i + 3] = y1;
    g->tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
  }

  /* ---- App. A.2 ---- */
  int64_t D = 0;
  int64_t* offs = (int64_t*)malloc(Pn * sizeof(int64_t));
  for (int i = 0; i < P; i++) { offs[i] = D; D += g->tiles_touched[i]; }
  g->D = D;
  dup_t* dups = (dup_t*)malloc((size_t)(D > 0 ? D : 1) * sizeof(dup_t));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (g->radii[i] <= 0) continue;
    int64_t o = offs[i];
    const int32_t* r = g->rect + 4 * (size_t)i;
    for (int y = r[1]; y < r[3]; y++)
      for (int x = r[0]; x < r[2]; x++) {
        dups[o].tile = (uint32_t)(y * gx + x);
        dups[o].id = (uint32_t)i;
        dups[o].depth = g->depth[i];
        o++;
      }
  }
  free(offs);
  qsort(dups, (size_t)D, sizeof(dup_t), cmp_dup);
  int Tn = gx * gy;
  g->list = (uint32_t*)malloc((size_t)(D > 0 ? D : 1) * sizeof(uint32_t));
  g->ranges = (int64_t*)calloc((size_t)Tn * 2, sizeof(int64_t));
  for (int64_t k = 0; k < D; k++) {
    g->list[k] = dups[k].id;
    uint32_t t = dups[k].tile;
    if (k == 0 || dups[k - 1].tile != t) g->ranges[2 * t] = k;
    if (k == D - 1 || dups[k + 1].tile != t) g->ranges[2 * t + 1] = k + 1;
  }
  free(dups);

  /* ---- App. A.3 ---- */
  size_t N = (size_t)W * H;
  g->final_T = (real*)malloc(N * sizeof(real));
  g->n_contrib = (uint32_t*)malloc(N * sizeof(uint32_t));
  g->n_walked = (uint32_t*)malloc(N * sizeof(uint32_t));
  g->out_color = (real*)malloc(3 * N * sizeof(real));
  int64_t consumed = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : consumed)
  for (int t = 0; t < Tn; t++) {
    int tx = t % gx, ty = t / gx;
    int64_t r0 = g->ranges[2 * t], r1 = g->ranges[2 * t + 1];
    uint32_t tile_walk = 0;
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        real T = 1, C[3] = {0, 0, 0}, Dp = 0, Aa = 0;
        uint32_t contributor = 0, last = 0;
        for (int64_t k = r0; k < r1; k++) {
          contributor++;
          uint32_t id = g->list[k];
          real dx = g->xy[2 * id] - (real)px, dy = g->xy[2 * id + 1] - (real)py;
          const real* co = g->conic_o + 4 * (size_t)id;
          real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > (real)0) continue;
          real alpha = co[3] * R_EXP(power);
          if (alpha > K_ALPHA_MAX) alpha = K_ALPHA_MAX;
          if (alpha < K_ALPHA_MIN) continue;
          real test = T * ((real)1 - alpha);
          if (test < K_T_MIN) break; /* this splat is NOT applied */
          const real* col = g->rgb + 3 * (size_t)id;
          for (int c = 0; c < 3; c++) C[c] += col[c] * alpha * T;
          Dp += g->depth[id] * alpha * T;
          Aa += alpha * T;
          T = test;
          last = contributor;
        }
        size_t pix = (size_t)py * W + px;
        g->final_T[pix] = T;
        g->n_contrib[pix] = last;
        g->n_walked[pix] = contributor;
        if (contributor > tile_walk) tile_walk = contributor;
        for (int c = 0; c < 3; c++) {
          real v = C[c] + T * g->bg[c];
          g->out_color[c * N + pix] = v;
          out_color[c * N + pix] = v;
        }
        out_depth[pix] = Dp;
        out_alpha[pix] = Aa;
      }
    consumed += tile_walk;
  }
  g->consumed_fwd = consumed;
  if (out_radii) memcpy(out_radii, g->radii, (size_t)P * sizeof(int32_t));
  return g;
}

static inline void atomic_add_d(double* p, double v) {
#pragma omp atomic
  *p += v;
}

/*
 * Backward: App. A.4 (composite) -> A.5 (projection).  Gradient conventions of App. A.6.
 * dL_dcolor (3,H,W), dL_ddepth (H,W) or NULL, dL_dalpha (H,W) or NULL.
 * Outputs (any may be NULL): dmeans3D (P,3), dmeans2D (P,3), dshs (P,M,3), dcolors (P,3),
 * dopac (P), dscales (P,3), drots (P,4), dcov3D (P,6).
 * Per-Gaussian sums are accumulated in double and rounded once at the end.
 */
void gso_backward(gso_ctx* g, const real* dL_dcolor, const real* dL_ddepth, const real* dL_dalpha_px, real* dmeans3D,
                  real* dmeans2D, real* dshs, real* dcolors, real* dopac, real* dscales, real* drots, real* dcov3D) {
  const int P = g->P, W = g->W, H = g->H, gx = g->gx, Tn = g->gx * g->gy;
  const size_t N = (size_t)W * H;
  size_t Pn = P > 0 ? (size_t)P : 1;
  /* accumulators: mean2D(2) conic(3) opacity(1) colour(3) depth(1) */
  double* acc = (double*)calloc(Pn * 10, sizeof(double));
  int64_t consumed = 0;
  const real ddelx_dx = (real)0.5 * (real)W, ddely_dy = (real)0.5 * (real)H;

#pragma omp parallel for schedule(dynamic, 1) reduction(+ : consumed)
  for (int t = 0; t < Tn; t++) {
    int tx = t % gx, ty = t / gx;
    int64_t r0 = g->ranges[2 * t];
    uint32_t tile_max = 0;
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        size_t pix = (size_t)py * W + px;
        const real T_final = g->final_T[pix];
        real T = T_final;
        uint32_t last = g->n_contrib[pix];
        if (last > tile_max) tile_max = last;
        real gpx[3] = {dL_dcolor[pix], dL_dcolor[N + pix], dL_dcolor[2 * N + pix]};
        real gd = dL_ddepth ? dL_ddepth[pix] : (real)0;
        real ga = dL_dalpha_px ? dL_dalpha_px[pix] : (real)0;
        real accum_c[3] = {0, 0, 0}, last_c[3] = {0, 0, 0};
        real accum_d = 0, last_d = 0, accum_a = 0, last_alpha = 0;
        real bg_dot = g->bg[0] * gpx[0] + g->bg[1] * gpx[1] + g->bg[2] * gpx[2];
        for (int64_t k = r0 + (int64_t)last - 1; k >= r0; k--) {
          uint32_t id = g->list[k];
          real dx = g->xy[2 * id] - (real)px, dy = g->xy[2 * id + 1] - (real)py;
          const real* co = g->conic_o + 4 * (size_t)id;
          real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > (real)0) continue;
          real G = R_EXP(power);
          real alpha = co[3] * G;
          if (alpha > K_ALPHA_MAX) alpha = K_ALPHA_MAX;
          if (alpha < K_ALPHA_MIN) continue;
          T = T / ((real)1 - alpha);
          real w = alpha * T;
          real dLda = 0;
          const real* col = g->rgb + 3 * (size_t)id;
          double* a = acc + 10 * (size_t)id;
          for (int c = 0; c < 3; c++) {
            accum_c[c] = last_alpha * last_c[c] + ((real)1 - last_alpha) * accum_c[c];
            last_c[c] = col[c];
            dLda += (col[c] - accum_c[c]) * gpx[c];
            atomic_add_d(a + 6 + c, (double)(w * gpx[c]));
          }
          real zd = g->depth[id];
          accum_d = last_alpha * last_d + ((real)1 - last_alpha) * accum_d;
          last_d = zd;
          dLda += (zd - accum_d) * gd;
          atomic_add_d(a + 9, (double)(w * gd));
          accum_a = last_alpha + ((real)1 - last_alpha) * accum_a;
          dLda += ((real)1 - accum_a) * ga;
          dLda *= T;
          last_alpha = alpha;
          dLda += (-T_final / ((real)1 - alpha)) * bg_dot;
          /* App. A.6(i): the 0.99 clamp is ignored on the way back */
          real dLdG = co[3] * dLda;
          real gdx = G * dx, gdy = G * dy;
          real dG_ddx = -gdx * co[0] - gdy * co[1];
          real dG_ddy = -gdy * co[2] - gdx * co[1];
          atomic_add_d(a + 0, (double)(dLdG * dG_ddx * ddelx_dx));
          atomic_add_d(a + 1, (double)(dLdG * dG_ddy * ddely_dy));
          atomic_add_d(a + 2, (double)((real)-0.5 * gdx * dx * dLdG));
          atomic_add_d(a + 3, (double)((real)-0.5 * gdx * dy * dLdG));
          atomic_add_d(a + 4, (double)((real)-0.5 * gdy * dy * dLdG));
          atomic_add_d(a + 5, (double)(G * dLda));
        }
      }
    consumed += tile_max;
  }
  g->consumed_bwd = consumed;

  const real fx = (real)W / ((real)2 * g->tanfovx), fy = (real)H / ((real)2 * g->tanfovy);
  const int M = g->M;
  if (dmeans3D) memset(dmeans3D, 0, Pn * 3 * sizeof(real));
  if (dmeans2D) memset(dmeans2D, 0, Pn * 3 * sizeof(real));
  if (dshs && M > 0) memset(dshs, 0, Pn * (size_t)M * 3 * sizeof(real));
  if (dcolors) memset(dcolors, 0, Pn * 3 * sizeof(real));
  if (dopac) memset(dopac, 0, Pn * sizeof(real));
  if (dscales) memset(dscales, 0, Pn * 3 * sizeof(real));
  if (drots) memset(drots, 0, Pn * 4 * sizeof(real));
  if (dcov3D) memset(dcov3D, 0, Pn * 6 * sizeof(real));

  /* ---- App. A.5 ---- */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (!(g->radii[i] > 0)) continue;
    const double* a = acc + 10 * (size_t)i;
    real dm2[2] = {(real)a[0], (real)a[1]};
    real dcon[3] = {(real)a[2], (real)a[3], (real)a[4]};
    real dcol[3] = {(real)a[6], (real)a[7], (real)a[8]};
    real ddep = (real)a[9];
    if (dmeans2D) { dmeans2D[3 * i] = dm2[0]; dmeans2D[3 * i + 1] = dm2[1]; }
    if (dopac) dopac[i] = (real)a[5];
    if (dcolors) for (int c = 0; c < 3; c++) dcolors[3 * i + c] = dcol[c];

    const real* p = g->means3D + 3 * (size_t)i;
    const real* c6 = g->cov3D + 6 * (size_t)i;
    ewa_t e;
    ewa_project(p, c6, g, fx, fy, &e);
    real A = e.a, B = e.b, Cc = e.c;
    real denom = A * Cc - B * B;
    real d2inv = (real)1 / (denom * denom + K_EPS_W); /* App. A.6(ii) */
    real dLa = 0, dLb = 0, dLc = 0;
    real dS[6] = {0, 0, 0, 0, 0, 0};
    if (d2inv != (real)0) {
      dLa = d2inv * (-Cc * Cc * dcon[0] + (real)2 * B * Cc * dcon[1] + (denom - A * Cc) * dcon[2]);
      dLc = d2inv * (-A * A * dcon[2] + (real)2 * A * B * dcon[1] + (denom - A * Cc) * dcon[0]);
      dLb = d2inv * (real)2 * (B * Cc * dcon[0] - (denom + (real)2 * B * B) * dcon[1] + A * B * dcon[2]);
      const real* A0 = e.A[0];
      const real* A1 = e.A[1];
      /* diagonal entries once, off-diagonals carry both (i,j) and (j,i) */
      dS[0] = A0[0] * A0[0] * dLa + A0[0] * A1[0] * dLb + A1[0] * A1[0] * dLc;
      dS[3] = A0[1] * A0[1] * dLa + A0[1] * A1[1] * dLb + A1[1] * A1[1] * dLc;
      dS[5] = A0[2] * A0[2] * dLa + A0[2] * A1[2] * dLb + A1[2] * A1[2] * dLc;
      dS[1] = (real)2 * A0[0] * A0[1] * dLa + (A0[0] * A1[1] + A0[1] * A1[0]) * dLb + (real)2 * A1[0] * A1[1] * dLc;
      dS[2] = (real)2 * A0[0] * A0[2] * dLa + (A0[0] * A1[2] + A0[2] * A1[0]) * dLb + (real)2 * A1[0] * A1[2] * dLc;
      dS[4] = (real)2 * A0[2] * A0[1] * dLa + (A0[1] * A1[2] + A0[2] * A1[1]) * dLb + (real)2 * A1[1] * A1[2] * dLc;
    }
    if (dcov3D && g->has_cov_precomp) for (int k = 0; k < 6; k++) dcov3D[6 * i + k] = dS[k];
    /* dL/dA rows: dA0 = 2 (Sigma A0) dLa + (Sigma A1) dLb ; dA1 = 2 (Sigma A1) dLc + (Sigma A0) dLb */
    real S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    real SA0[3], SA1[3];
    for (int k = 0; k < 3; k++) {
      SA0[k] = e.A[0][0] * S[k][0] + e.A[0][1] * S[k][1] + e.A[0][2] * S[k][2];
      SA1[k] = e.A[1][0] * S[k][0] + e.A[1][1] * S[k][1] + e.A[1][2] * S[k][2];
    }
    real dA0[3], dA1[3];
    for (int k = 0; k < 3; k++) {
      dA0[k] = (real)2 * SA0[k] * dLa + SA1[k] * dLb;
      dA1[k] = (real)2 * SA1[k] * dLc + SA0[k] * dLb;
    }
    const real* v = g->view;
    /* dJ[r][m] = sum_k dA_r[k] * Rv[m][k],  Rv[m][k] = v[4k+m] */
    real dJ00 = dA0[0] * v[0] + dA0[1] * v[4] + dA0[2] * v[8];
    real dJ02 = dA0[0] * v[2] + dA0[1] * v[6] + dA0[2] * v[10];
    real dJ11 = dA1[0] * v[1] + dA1[1] * v[5] + dA1[2] * v[9];
    real dJ12 = dA1[0] * v[2] + dA1[1] * v[6] + dA1[2] * v[10];
    real tz = (real)1 / e.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    real dtx = e.xmul * -fx * tz2 * dJ02; /* App. A.6(iii) */
    real dty = e.ymul * -fy * tz2 * dJ12;
    real dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + ((real)2 * fx * e.t[0]) * tz3 * dJ02 + ((real)2 * fy * e.t[1]) * tz3 * dJ12;
    /* dmean = Rv^T dt */
    real dm[3];
    dm[0] = v[0] * dtx + v[1] * dty + v[2] * dtz;
    dm[1] = v[4] * dtx + v[5] * dty + v[6] * dtz;
    dm[2] = v[8] * dtx + v[9] * dty + v[10] * dtz;

    /* projection path (dL/dmean2D is already NDC-scaled, App. A.5) */
    const real* pm = g->proj;
    real mh[4];
    xform4x4(p, pm, mh);
    real mw = (real)1 / (mh[3] + K_EPS_W);
    real mul1 = (pm[0] * p[0] + pm[4] * p[1] + pm[8] * p[2] + pm[12]) * mw * mw;
    real mul2 = (pm[1] * p[0] + pm[5] * p[1] + pm[9] * p[2] + pm[13]) * mw * mw;
    dm[0] += (pm[0] * mw - pm[3] * mul1) * dm2[0] + (pm[1] * mw - pm[3] * mul2) * dm2[1];
    dm[1] += (pm[4] * mw - pm[7] * mul1) * dm2[0] + (pm[5] * mw - pm[7] * mul2) * dm2[1];
    dm[2] += (pm[8] * mw - pm[11] * mul1) * dm2[0] + (pm[9] * mw - pm[11] * mul2) * dm2[1];
    /* depth path: depth = row 2 of V . [p,1] */
    dm[0] += v[2] * ddep;
    dm[1] += v[6] * ddep;
    dm[2] += v[10] * ddep;

    /* SH path */
    if (g->has_sh) {
      const real* sh = g->shs + (size_t)i * M * 3;
      real* dsh = dshs ? dshs + (size_t)i * M * 3 : NULL;
      real d[3] = {p[0] - g->campos[0], p[1] - g->campos[1], p[2] - g->campos[2]};
      real n = R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      real x = d[0] / n, y = d[1] / n, z = d[2] / n;
      real ddir[3] = {0, 0, 0};
      int deg = g->deg;
      for (int c = 0; c < 3; c++) {
        real gc = g->clamped[3 * i + c] ? (real)0 : dcol[c];
#define SH(k) sh[(k) * 3 + c]
#define DSH(k, val) do { if (dsh) dsh[(k) * 3 + c] = (val) * gc; } while (0)
        real drx = 0, dry = 0, drz = 0; /* d rgb_c / d (x,y,z) */
        DSH(0, SH_C0);
        if (deg > 0) {
          DSH(1, -SH_C1 * y); DSH(2, SH_C1 * z); DSH(3, -SH_C1 * x);
          drx += -SH_C1 * SH(3); dry += -SH_C1 * SH(1); drz += SH_C1 * SH(2);
          if (deg > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * ((real)2 * zz - xx - yy));
            DSH(7, SH_C2[3] * xz); DSH(8, SH_C2[4] * (xx - yy));
            drx += SH_C2[0] * y * SH(4) + SH_C2[2] * (real)2 * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * (real)2 * x * SH(8);
            dry += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * (real)2 * -y * SH(6) + SH_C2[4] * (real)2 * -y * SH(8);
            drz += SH_C2[1] * y * SH(5) + SH_C2[2] * (real)2 * (real)2 * z * SH(6) + SH_C2[3] * x * SH(7);
            if (deg > 2) {
              DSH(9, SH_C3[0] * y * ((real)3 * xx - yy));
              DSH(10, SH_C3[1] * xy * z);
              DSH(11, SH_C3[2] * y * ((real)4 * zz - xx - yy));
              DSH(12, SH_C3[3] * z * ((real)2 * zz - (real)3 * xx - (real)3 * yy));
              DSH(13, SH_C3[4] * x * ((real)4 * zz - xx - yy));
              DSH(14, SH_C3[5] * z * (xx - yy));
              DSH(15, SH_C3[6] * x * (xx - (real)3 * yy));
              drx += SH_C3[0] * SH(9) * (real)3 * (real)2 * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -(real)2 * xy +
                     SH_C3[3] * SH(12) * -(real)3 * (real)2 * xz + SH_C3[4] * SH(13) * (-(real)3 * xx + (real)4 * zz - yy) +
                     SH_C3[5] * SH(14) * (real)2 * xz + SH_C3[6] * SH(15) * (real)3 * (xx - yy);
              dry += SH_C3[0] * SH(9) * (real)3 * (xx - yy) + SH_C3[1] * SH(10) * xz +
                     SH_C3[2] * SH(11) * (-(real)3 * yy + (real)4 * zz - xx) + SH_C3[3] * SH(12) * -(real)3 * (real)2 * yz +
                     SH_C3[4] * SH(13) * -(real)2 * xy + SH_C3[5] * SH(14) * -(real)2 * yz + SH_C3[6] * SH(15) * -(real)3 * (real)2 * xy;
              drz += SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * (real)4 * (real)2 * yz +
                     SH_C3[3] * SH(12) * (real)3 * ((real)2 * zz - xx - yy) + SH_C3[4] * SH(13) * (real)4 * (real)2 * xz +
                     SH_C3[5] * SH(14) * (xx - yy);
            }
          }
        }
#undef SH
#undef DSH
        ddir[0] += drx * gc;
        ddir[1] += dry * gc;
        ddir[2] += drz * gc;
      }
      /* through dir = d/|d| : J = (I - dir dir^T)/|d| */
      real dot = x * ddir[0] + y * ddir[1] + z * ddir[2];
      dm[0] += (ddir[0] - x * dot) / n;
      dm[1] += (ddir[1] - y * dot) / n;
      dm[2] += (ddir[2] - z * dot) / n;
    }
    if (dmeans3D) { dmeans3D[3 * i] = dm[0]; dmeans3D[3 * i + 1] = dm[1]; dmeans3D[3 * i + 2] = dm[2]; }

    /* Cov3D path: Sigma = R S^2 R^T (gradient w.r.t. UN-normalised quaternion) */
    if (!g->has_cov_precomp && (dscales || drots)) {
      const real* q = g->rots + 4 * (size_t)i;
      const real* sc = g->scales + 3 * (size_t)i;
      real mod = g->scale_modifier;
      real R[3][3];
      quat_to_R(q, R);
      real s[3] = {mod * sc[0], mod * sc[1], mod * sc[2]};
      /* symmetric dL/dSigma: off-diagonals split in two halves */
      real G3[3][3] = {{dS[0], (real)0.5 * dS[1], (real)0.5 * dS[2]},
                       {(real)0.5 * dS[1], dS[3], (real)0.5 * dS[4]},
                       {(real)0.5 * dS[2], (real)0.5 * dS[4], dS[5]}};
      /* N = R S (columns of R scaled): Sigma = N N^T ; dL/dN = 2 G3 N */
      real Nm[3][3], dN[3][3];
      for (int a2 = 0; a2 < 3; a2++)
        for (int k = 0; k < 3; k++) Nm[a2][k] = R[a2][k] * s[k];
      for (int a2 = 0; a2 < 3; a2++)
        for (int k = 0; k < 3; k++) dN[a2][k] = (real)2 * (G3[a2][0] * Nm[0][k] + G3[a2][1] * Nm[1][k] + G3[a2][2] * Nm[2][k]);
      if (dscales)
        for (int k = 0; k < 3; k++) dscales[3 * i + k] = mod * (R[0][k] * dN[0][k] + R[1][k] * dN[1][k] + R[2][k] * dN[2][k]);
      if (drots) {
        real dR[3][3];
        for (int a2 = 0; a2 < 3; a2++)
          for (int k = 0; k < 3; k++) dR[a2][k] = dN[a2][k] * s[k];
        real r = q[0], x = q[1], y = q[2], z = q[3];
        /* differentiate quat_to_R entry by entry */
        drots[4 * i + 0] = (real)2 * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
        drots[4 * i + 1] = (real)2 * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - (real)2 * x * dR[1][1] - r * dR[1][2] + z * dR[2][0] + r * dR[2][1] - (real)2 * x * dR[2][2]);
        drots[4 * i + 2] = (real)2 * (-(real)2 * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] - r * dR[2][0] + z * dR[2][1] - (real)2 * y * dR[2][2]);
        drots[4 * i + 3] = (real)2 * (-(real)2 * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - (real)2 * z * dR[1][1] + y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
      }
    }
  }
  free(acc);
}

/*
 * Decision-margin analysis for parity tests.  The composite takes discrete decisions per (pixel, splat):
 * power > 0, alpha < 1/255, T(1-alpha) < 1e-4.  Two fp32 implementations that differ by an ulp in exp() legitimately
 * take different branches when a pair sits on a threshold, and the pixel then differs by up to ~alpha*T.  This
 * re-walks the forward and flags every pixel at which some decision had a relative margin below eps (alpha test:
 * eps_alpha, transmittance test: eps_T, power test: |power| < 1e-6), and every Gaussian listed in such a pixel's
 * tile (its gradients inherit the flip).  Tests hold flagged elements to a looser bound and assert they are rare.
 */
void gso_fragility(const gso_ctx* g, real eps_alpha, real eps_T, uint8_t* pix_mask, uint8_t* gauss_mask) {
  const int W = g->W, H = g->H, gx = g->gx, Tn = g->gx * g->gy;
  memset(pix_mask, 0, (size_t)W * H);
  memset(gauss_mask, 0, (size_t)(g->P > 0 ? g->P : 1));
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < Tn; t++) {
    int tx = t % gx, ty = t / gx;
    int64_t r0 = g->ranges[2 * t], r1 = g->ranges[2 * t + 1];
    int tile_fragile = 0;
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        real T = 1;
        int fragile = 0;
        for (int64_t k = r0; k < r1; k++) {
          uint32_t id = g->list[k];
          real dx = g->xy[2 * id] - (real)px, dy = g->xy[2 * id + 1] - (real)py;
          const real* co = g->conic_o + 4 * (size_t)id;
          real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > (real)-1e-6 && power < (real)1e-6) fragile = 1;
          if (power > (real)0) continue;
          real araw = co[3] * R_EXP(power);
          real da = araw - K_ALPHA_MIN;
          if (da < 0) da = -da;
          /* the exponent is a sum of terms that cancel for thin, rotated splats: its fp32 rounding error scales with
             the magnitude of the terms, not of the result (a few ulp of `mag`), and alpha inherits it 1:1 */
          real mag = (real)0.5 * ((co[0] < 0 ? -co[0] : co[0]) * dx * dx + (co[2] < 0 ? -co[2] : co[2]) * dy * dy);
          real cross = co[1] * dx * dy;
          mag += cross < 0 ? -cross : cross;
          if (da <= (eps_alpha + (real)4e-7 * mag) * K_ALPHA_MIN) fragile = 1;
          real alpha = araw > K_ALPHA_MAX ? K_ALPHA_MAX : araw;
          if (alpha < K_ALPHA_MIN) continue;
          real test = T * ((real)1 - alpha);
          real dt = test - K_T_MIN;
          if (dt < 0) dt = -dt;
          if (dt <= eps_T * K_T_MIN) fragile = 1;
          if (test < K_T_MIN) break;
          T = test;
        }
        if (fragile) {
          pix_mask[(size_t)py * W + px] = 1;
          tile_fragile = 1;
          /* every splat that could contribute here under either outcome of the flipped decision */
          for (int64_t k = r0; k < r1; k++) {
            uint32_t id = g->list[k];
            real dx = g->xy[2 * id] - (real)px, dy = g->xy[2 * id + 1] - (real)py;
            const real* co = g->conic_o + 4 * (size_t)id;
            real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > (real)1e-6) continue;
            if (co[3] * R_EXP(power) < (real)0.5 * K_ALPHA_MIN) continue;
            gauss_mask[id] = 1; /* benign race: all writers store 1 */
          }
        }
      }
    (void)tile_fragile;
  }
}

/* markVisible [EXT]: z_view > 0.2 (App. A.1 step 1). */
void gso_mark_visible(int P, const real* means3D, const real* viewmatrix, uint8_t* present) {
  for (int i = 0; i < P; i++) {
    real pv[3];
    xform4x3(means3D + 3 * (size_t)i, viewmatrix, pv);
    present[i] = pv[2] > K_NEAR;
  }
}

/* ---- accessors for stage-by-stage parity checks ---- */
int64_t gso_num_dups(const gso_ctx* g) { return g->D; }
int64_t gso_consumed_fwd(const gso_ctx* g) { return g->consumed_fwd; }
int64_t gso_consumed_bwd(const gso_ctx* g) { return g->consumed_bwd; }
const real* gso_xy(const gso_ctx* g) { return g->xy; }
const real* gso_depth(const gso_ctx* g) { return g->depth; }
const real* gso_conic_opacity(const gso_ctx* g) { return g->conic_o; }
const real* gso_cov3D(const gso_ctx* g) { return g->cov3D; }
const real* gso_rgb(const gso_ctx* g) { return g->rgb; }
const int32_t* gso_rect(const gso_ctx* g) { return g->rect; }
const uint32_t* gso_tiles_touched(const gso_ctx* g) { return g->tiles_touched; }
const uint32_t* gso_list(const gso_ctx* g) { return g->list; }
const int64_t* gso_ranges(const gso_ctx* g) { return g->ranges; }
const real* gso_final_T(const gso_ctx* g) { return g->final_T; }
const uint32_t* gso_n_contrib(const gso_ctx* g) { return g->n_contrib; }
int gso_real_bytes(void) { return (int)sizeof(real); }
int gso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void gso_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
