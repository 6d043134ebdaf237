/*
 * oracle.c — CPU restatement of the VkScanlinePR compute shaders.
 * TEST INFRASTRUCTURE ONLY; parity pinned against the shipped SPIR-V by interpretation — see oracle.h.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
 */
#define _POSIX_C_SOURCE 200809L /* clock_gettime */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FRAG_SIZE 2
#define T_LINE 0x02u
#define T_QUADRIC 0x03u
#define T_CUBIC 0x04u
#define T_ARC 0x13u
#define CUBIC_ITERATION_NUMBER 24 /* MI1:7 */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
/* A NaN that an arithmetic instruction PRODUCES has an implementation-defined bit pattern: x86 SSE writes the negative
 * default NaN 0xFFC00000, NVIDIA GPUs — the reference's target and ours — the positive 0x7FFFFFFF, and pass it on
 * unchanged. Wherever the path looks at the BITS of a computed float (the sign tests of the bisection, MI1:421; the
 * parameters, cuts and transformed points it stores) the restatement therefore pins the GPU's pattern. Found by
 * tools/fuzz_parity.py: a line whose y extent is a denormal has 1/a = inf and (c - c0) * a = 0 * inf. */
static inline float canon(float x) { return (x != x) ? u2f(0x7FFFFFFFu) : x; }

/* GLSL int(x): round toward zero. Out-of-range is undefined in GLSL; we pin it to the
 * saturating behaviour of the GPU conversion instruction (NaN -> 0) so both sides agree. */
static inline int f2i(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)x;
}

/* MI0:10, MI1:12, GF:10: LERP(a,b,t) = a + t*(b-a), three separately rounded fp32 ops. */
/* SLPR_FLAG_CONTRACT_FMA / orc_set_contract_fma (NOT pinned by the reference: no driver here): the
 * other legitimate reading of the shaders, in which the compiler contracts a*b + c into one fused multiply-add. The
 * policy, mirrored by csrc/common.cuh: (1) LERP(a,b,t) = fma(t, b - a, a); (2) dot(vec4(x,y,0,1), m) accumulated left to
 * right, x*m.x then one fma per further term; (3) MI0's a = fma(3, x1 - x2, x3 - x0) and the discriminant
 * fma(B, B, -(A*C)); (4) the f-1 arc evaluator's numerator as an fma chain. Nothing else has the a*b + c shape. */
static int g_contract_fma = 0;
void orc_set_contract_fma(int on) { g_contract_fma = on; }
static inline float madd(float a, float b, float c) { /* a * b + c under the current policy */
    if (g_contract_fma) return fmaf(a, b, c);
    float m = a * b;
    return m + c;
}

static inline float lerpf(float a, float b, float t) {
    float d = b - a;
    if (g_contract_fma) return fmaf(t, d, a);
    float m = t * d;
    return a + m;
}

/* MI1:76-79, GF:54-57 */
static inline int float2int_rd(float x) { return x >= 0.0f ? f2i(x) : f2i(x - 1.0f); }

/* MI0:14-20 */
static inline int path_invisible(int32_t mask) {
    uint32_t m = (uint32_t)mask;
    return ((m & 0x11111000u) == 0) || ((m & 0x01101011u) == 0) ||
           ((m & 0x00011111u) == 0) || ((m & 0x11010110u) == 0);
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return imin(imax(v, lo), hi); } /* GLSL clamp */

/* ------------------------------------------------------------------ TP:33-85 */
void orc_transform(uint32_t n_points, const float *pos, const uint32_t *pos_path,
                   const float *rows, int width, int height,
                   float *tpos_out, int32_t *path_visible) {
    const float w = (float)width, h = (float)height; /* TP:8, SR:1153-1154 */
    for (uint32_t i = 0; i < n_points; ++i) {
        float x = pos[2 * i], y = pos[2 * i + 1];
        float op[4];
        for (int r = 0; r < 4; ++r) { /* TP:41-46; dot evaluated left to right */
            const float *m = rows + 4 * r;
            float s;
            if (g_contract_fma) {
                s = x * m[0];
                s = fmaf(y, m[1], s);
                s = fmaf(0.0f, m[2], s);
                s = fmaf(1.0f, m[3], s);
            } else {
                float a = x * m[0];
                float b = y * m[1];
                float c = 0.0f * m[2];
                float d = 1.0f * m[3];
                s = a + b;
                s = s + c;
                s = s + d;
            }
            op[r] = s;
        }
        op[0] = op[0] / op[3]; /* TP:53-55 */
        op[1] = op[1] / op[3];
        int xf = op[0] < 0 ? 0 : (op[0] < w ? 1 : 2); /* TP:67-68 */
        int yf = op[1] < 0 ? 0 : (op[1] < h ? 1 : 2);
        uint32_t flag = 0;
        switch ((yf << 4) | xf) { /* TP:71-82 */
        case 0x00: flag = 0x10000000u; break;
        case 0x01: flag = 0x01000000u; break;
        case 0x02: flag = 0x00100000u; break;
        case 0x10: flag = 0x00010000u; break;
        case 0x11: flag = 0x10000001u; break;
        case 0x12: flag = 0x00001000u; break;
        case 0x20: flag = 0x00000100u; break;
        case 0x21: flag = 0x00000010u; break;
        case 0x22: flag = 0x00000001u; break;
        default: break;
        }
        path_visible[pos_path[i]] = (int32_t)((uint32_t)path_visible[pos_path[i]] | flag);
        tpos_out[2 * i] = canon(op[0]);
        tpos_out[2 * i + 1] = canon(op[1]);
    }
}

/* ------------------------------------------------------------------ MI0 helpers */
/* MI0:76-129 */
static void solve_quad(float a, float b, float c, float *r0, float *r1) {
    if (a == 0) {
        float x = -c / b;
        *r0 = x; *r1 = x;
        return;
    }
    float A = a, B = b * 0.5f, C = c;
    float tx = 0.f, ty = 0.f;
    float bb = B * B, ac = A * C;
    float R = g_contract_fma ? fmaf(B, B, -ac) : bb - ac;
    if (R > 0.0f) {
        float SR = sqrtf(R);
        if (B > 0.0f) {
            float TB = B + SR;
            tx = -C / TB; ty = -TB / A;
        } else {
            float TB = -B + SR;
            tx = TB / A; ty = C / TB;
        }
    }
    *r0 = tx; *r1 = ty;
}

/* ---- SURVEY section 8 f-1 (NOT in the reference: its QUADRIC / ARC arms are `// TODO`). With orc_set_full_rvg(1, w)
 * the TODO arms get real arithmetic, defined HERE and mirrored operation for operation by the CUDA kernels
 * (csrc/geom.cuh, SLPR_FLAG_FULL_RVG): QUADRIC by de Casteljau with the reference's LERP; ARC = rational quadratic
 * with weights (1, w, 1), w per curve, control point Euclidean; monotonic cuts from the zero of the derivative's
 * numerator w(p1-p0)u^2 + (p2-p0)ut + w(p2-p1)t^2 through the reference's own solveQuadEquation; crossings by
 * the reference's 24-step bisection (MI1:392-436) with this evaluator. Default (off): reference behaviour. */
static int g_full_rvg = 0;
static const float *g_curve_weight = NULL;
void orc_set_full_rvg(int on, const float *curve_weight) { g_full_rvg = on; g_curve_weight = on ? curve_weight : NULL; }

static inline float eval_quadric(const float *p, float t) {
    float q0 = lerpf(p[0], p[1], t), q1 = lerpf(p[1], p[2], t);
    return lerpf(q0, q1, t);
}
static inline float eval_arc(const float *p, float t) { /* p[3] = weight of the middle control point */
    float u = 1.0f - t;
    float b0 = u * u, b2 = t * t;
    float tt = 2.0f * t;
    float tu = tt * u;
    float b1 = tu * p[3];
    float d01 = b0 + b1;
    float D = d01 + b2;
    float n0 = b0 * p[0];
    float n01 = madd(b1, p[1], n0);
    float N = madd(b2, p[2], n01);
    return N / D;
}

/* MI0:131-165 (dflt = 1.0f) and MI1:81-148 (dflt = 0.0f) */
static inline float interp_general(uint32_t type, float t, const float *p, float dflt) {
    if (type == T_LINE) return lerpf(p[0], p[1], t);
    if (type == T_CUBIC) {
        float q0 = lerpf(p[0], p[1], t), q1 = lerpf(p[1], p[2], t), q2 = lerpf(p[2], p[3], t);
        float l0 = lerpf(q0, q1, t), l1 = lerpf(q1, q2, t);
        return lerpf(l0, l1, t);
    }
    if (g_full_rvg && type == T_QUADRIC) return eval_quadric(p, t);
    if (g_full_rvg && type == T_ARC) return eval_arc(p, t);
    return dflt; /* QUADRIC / ARC: TODO arms in the reference */
}

/* MI0:167-183 (floor) */
static void xy_begin_end_floor(float p0x, float p0y, float p1x, float p1y,
                               int *xb, int *xe, int *yb, int *ye) {
    if (p0x <= p1x) { *xb = f2i(floorf(p0x / FRAG_SIZE) * FRAG_SIZE) + FRAG_SIZE; *xe = f2i(floorf(p1x / FRAG_SIZE) * FRAG_SIZE); }
    else            { *xb = f2i(floorf(p1x / FRAG_SIZE) * FRAG_SIZE) + FRAG_SIZE; *xe = f2i(floorf(p0x / FRAG_SIZE) * FRAG_SIZE); }
    if (p0y <= p1y) { *yb = f2i(floorf(p0y / FRAG_SIZE) * FRAG_SIZE) + FRAG_SIZE; *ye = f2i(floorf(p1y / FRAG_SIZE) * FRAG_SIZE); }
    else            { *yb = f2i(floorf(p1y / FRAG_SIZE) * FRAG_SIZE) + FRAG_SIZE; *ye = f2i(floorf(p0y / FRAG_SIZE) * FRAG_SIZE); }
}

/* MI1:150-170 (float2int_rd, plus direction) */
static void xy_begin_end_delta(float p0x, float p0y, float p1x, float p1y,
                               int *xb, int *xe, int *yb, int *ye, float *dx, float *dy) {
    if (p0x <= p1x) { *xb = float2int_rd(p0x / FRAG_SIZE) * FRAG_SIZE + FRAG_SIZE; *xe = float2int_rd(p1x / FRAG_SIZE) * FRAG_SIZE; *dx = FRAG_SIZE; }
    else            { *xb = float2int_rd(p1x / FRAG_SIZE) * FRAG_SIZE + FRAG_SIZE; *xe = float2int_rd(p0x / FRAG_SIZE) * FRAG_SIZE; *dx = -FRAG_SIZE; }
    if (p0y <= p1y) { *yb = float2int_rd(p0y / FRAG_SIZE) * FRAG_SIZE + FRAG_SIZE; *ye = float2int_rd(p1y / FRAG_SIZE) * FRAG_SIZE; *dy = FRAG_SIZE; }
    else            { *yb = float2int_rd(p1y / FRAG_SIZE) * FRAG_SIZE + FRAG_SIZE; *ye = float2int_rd(p0y / FRAG_SIZE) * FRAG_SIZE; *dy = -FRAG_SIZE; }
}

/* MI0:186-221 == MI1:174-213 */
static void update_cut_range(int w, int h, int *xb, int *xe, int *yb, int *ye, int *nx, int *ny) {
    int cut_x_min = 0, cut_x_max = (int)((uint32_t)w & 0xFFFFFFFEu) + FRAG_SIZE;
    int cut_y_min = 0, cut_y_max = (int)((uint32_t)h & 0xFFFFFFFEu) + FRAG_SIZE;
    if ((*xb < cut_x_min && *xe < cut_x_min) || (*xb > cut_x_max && *xe > cut_x_max) || (*xb > *xe)) {
        *nx = 0;
    } else {
        *xb = iclamp(*xb, cut_x_min, cut_x_max);
        *xe = iclamp(*xe, cut_x_min, cut_x_max);
        *nx = imax((*xe - *xb) / FRAG_SIZE + 1, 0);
    }
    if ((*yb < cut_y_min && *ye < cut_y_min) || (*yb > cut_y_max && *ye > cut_y_max) || (*yb > *ye)) {
        *ny = 0;
    } else {
        *yb = iclamp(*yb, cut_y_min, cut_y_max);
        *ye = iclamp(*ye, cut_y_min, cut_y_max);
        *ny = imax((*ye - *yb) / FRAG_SIZE + 1, 0);
    }
}

static inline void load_points(uint32_t type, uint32_t po, const float *tpos, float *px, float *py, uint32_t curve) {
    for (uint32_t i = 0; i < 4; ++i) { /* MI0:251-257, MI1:237-243 */
        if (i < (type & 7u)) { px[i] = tpos[2 * (po + i)]; py[i] = tpos[2 * (po + i) + 1]; }
        else { px[i] = 0.f; py[i] = 0.f; } /* uninitialised shared memory in the reference; never consumed */
    }
    if (g_full_rvg && type == T_ARC) { /* f-1: the arc's weight rides in the unused fourth slot */
        float w = g_curve_weight ? g_curve_weight[curve] : 1.0f;
        px[3] = w; py[3] = w;
    }
}

/* ------------------------------------------------------------------ MI0:226-410 */
void orc_monotonize_count(uint32_t n_curves, const uint32_t *curve_type,
                          const uint32_t *curve_pos_map, const uint32_t *curve_path,
                          const float *tpos, const int32_t *path_visible,
                          int width, int height, float *cut_cache_out, int32_t *count_out) {
#pragma omp parallel for schedule(static)
    for (int64_t ci = 0; ci < (int64_t)n_curves; ++ci) {
        uint32_t c = (uint32_t)ci;
        uint32_t type = curve_type[c];
        float px[4], py[4];
        load_points(type, curve_pos_map[c], tpos, px, py, c);
        uint32_t n_cuts = 0;
        int visible = !path_invisible(path_visible[curve_path[c]]); /* MI0:260-261 */
        float tq[5] = {0, 0, 0, 0, 0};
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        if (visible) {
            for (int ax = 0; ax < 2; ++ax) { /* MI0:267-308 */
                if (type == T_CUBIC) {
                    const float *p = ax ? py : px;
                    float x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3];
                    float r0 = 0.f, r1 = 0.f;
                    float a = madd(3.0f, x1 - x2, x3 - x0);
                    float b = 2.0f * ((x0 - x1) + (x2 - x1));
                    float cc = x1 - x0;
                    solve_quad(a, b, cc, &r0, &r1);
                    if (r0 > 0.0f && r0 < 1.0f) tq[n_cuts++] = r0;
                    if (r1 > 0.0f && r1 < 1.0f && r1 != r0) tq[n_cuts++] = r1;
                } else if (g_full_rvg && (type == T_QUADRIC || type == T_ARC)) { /* f-1, see interp_general */
                    const float *p = ax ? py : px;
                    float w = (type == T_ARC) ? p[3] : 1.0f;
                    float d10 = p[1] - p[0], d20 = p[2] - p[0], d21 = p[2] - p[1];
                    float A = w * d10, B = d20, C = w * d21;
                    float amb = A - B;
                    float a = amb + C;
                    float a2 = 2.0f * A;
                    float b = B - a2;
                    float r0 = 0.f, r1 = 0.f;
                    solve_quad(a, b, A, &r0, &r1);
                    if (r0 > 0.0f && r0 < 1.0f) tq[n_cuts++] = r0;
                    if (r1 > 0.0f && r1 < 1.0f && r1 != r0) tq[n_cuts++] = r1;
                }
            }
            q0 = tq[0]; q1 = tq[1]; q2 = tq[2]; q3 = tq[3]; /* MI0:310-313 */
            if (n_cuts >= 2) { /* MI0:315-322 */
                float t1 = q1, t0 = q0;
                if (t1 < t0) { q1 = t0; q0 = t1; }
            }
            if (n_cuts >= 3) { /* MI0:323-337 */
                float t2 = q2, t1 = q1;
                if (t2 < t1) {
                    q2 = t1;
                    float t0 = q0;
                    if (t2 < t0) { q1 = t0; q0 = t2; }
                    else { q1 = t2; }
                }
            }
            if (n_cuts >= 4) { /* MI0:338-359 — literal, including `float t2 = q3;` at :340 */
                float t3 = q3;
                float t2 = q3;
                if (t3 < t2) {
                    q3 = t2;
                    float t1 = q1;
                    if (t3 < t1) {
                        q2 = t1;
                        float t0 = q0;
                        if (t3 < t0) { q1 = t0; q0 = t3; }
                        else { q1 = t3; }
                    } else { q2 = t3; }
                }
            }
        }
        tq[0] = q0; tq[1] = q1; tq[2] = q2; tq[3] = q3; /* MI0:363-366 */
        cut_cache_out[5 * c + 0] = canon(q0); /* MI0:368-372 */
        cut_cache_out[5 * c + 1] = canon(q1);
        cut_cache_out[5 * c + 2] = canon(q2);
        cut_cache_out[5 * c + 3] = canon(q3);
        cut_cache_out[5 * c + 4] = u2f(n_cuts);
        if (visible) { tq[n_cuts] = 1.f; ++n_cuts; } /* MI0:374-377 */

        float p0x = px[0], p0y = py[0];
        int pcnt = 0;
        for (uint32_t i = 0; i < n_cuts; ++i) { /* MI0:383-408 */
            float t1 = tq[i];
            float p1x = interp_general(type, t1, px, 1.0f);
            float p1y = interp_general(type, t1, py, 1.0f);
            int xb, xe, yb, ye, nx = 0, ny = 0;
            xy_begin_end_floor(p0x, p0y, p1x, p1y, &xb, &xe, &yb, &ye);
            update_cut_range(width, height, &xb, &xe, &yb, &ye, &nx, &ny);
            pcnt += 1 + nx + ny;
            p0x = p1x; p0y = p1y;
        }
        count_out[c] = pcnt;
    }
}

/* ------------------------------------------------------------------ SCAN */
void orc_exclusive_scan(int64_t n, const int32_t *in, int32_t *out) {
    int32_t acc = 0;
    for (int64_t i = 0; i < n; ++i) { int32_t v = in[i]; out[i] = acc; acc = (int32_t)((uint32_t)acc + (uint32_t)v); }
    out[n] = acc;
}

/* ------------------------------------------------------------------ MI1:217-447 */
void orc_intersect(uint32_t n_curves, const uint32_t *curve_type,
                   const uint32_t *curve_pos_map, const uint32_t *curve_path,
                   const float *tpos, const int32_t *path_visible,
                   const float *cut_cache, const int32_t *offsets,
                   int width, int height, int32_t *inter) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t ci = 0; ci < (int64_t)n_curves; ++ci) {
        uint32_t c = (uint32_t)ci;
        uint32_t type = curve_type[c];
        float P[10]; /* point_coords slots 0..3 = x, 4..7 = y, 8 = tx, 9 = ty */
        load_points(type, curve_pos_map[c], tpos, P, P + 4, c);
        int visible = !path_invisible(path_visible[curve_path[c]]);
        float tq[5];
        tq[0] = cut_cache[5 * c + 0]; tq[1] = cut_cache[5 * c + 1];
        tq[2] = cut_cache[5 * c + 2]; tq[3] = cut_cache[5 * c + 3]; tq[4] = 0.f;
        uint32_t n_cuts = f2u(cut_cache[5 * c + 4]); /* MI1:255 */
        if (visible) { tq[n_cuts] = 1.f; ++n_cuts; }

        float t0_ms = 0.f;
        float p0x = P[0], p0y = P[4];
        int pcnt = offsets[c]; /* MI1:265 */
        for (uint32_t i = 0; i < n_cuts; ++i) {
            float t1_ms = tq[i];
            float p1x = interp_general(type, t1_ms, P, 0.0f);
            float p1y = interp_general(type, t1_ms, P + 4, 0.0f);
            /* MI1:271-276 */
            t1_ms = u2f(f2u(t1_ms) & 0xFFFFFFFCu);
            if (floorf(p1x) == p1x) t1_ms = u2f((f2u(t1_ms) & 0xFFFFFFFCu) | 2u);
            else t1_ms = u2f(f2u(t1_ms) | 3u);

            int xb, xe, yb, ye, n_x = 0, n_y = 0;
            float dx, dy;
            xy_begin_end_delta(p0x, p0y, p1x, p1y, &xb, &xe, &yb, &ye, &dx, &dy);
            update_cut_range(width, height, &xb, &xe, &yb, &ye, &n_x, &n_y);
            int n_loop = n_x + n_y + 1;
            float x = (float)(dx < 0 ? xe : xb); /* MI1:301-302 */
            float y = (float)(dy < 0 ? ye : yb);
            P[8] = t0_ms; P[9] = t0_ms;
            int32_t i_inte_last = (int32_t)f2u(-1.0f);

            for (int it = -1; it < n_loop; ++it) { /* MI1:310-441 */
                float t_solve = 0.0f, cst = 0.0f;
                int side = 0;
                float next_tx = P[8], next_ty = P[9];
                float t_min;
                if (it == -1) {
                    t_min = next_tx;
                    if (n_x == 0) { side = 0; t_solve = 2.f; }
                    else if (n_y == 0) { side = 1; t_solve = 2.f; }
                    else { side = 0; --n_x; cst = x; x += dx; }
                } else if (next_tx <= next_ty) {
                    t_min = next_tx; side = 0;
                    if (n_x > 0) { --n_x; cst = x; x += dx; } else t_solve = 2.f;
                } else {
                    t_min = next_ty; side = 1;
                    if (n_y > 0) { --n_y; cst = y; y += dy; } else t_solve = 2.f;
                }
                if (it >= 0) { /* MI1:361-375 */
                    int32_t i_out = (int32_t)f2u(t_min);
                    int32_t i_t = (int32_t)(f2u(t_min) & 0xFFFFFFFCu);
                    if (i_t == (int32_t)((uint32_t)i_inte_last & 0xFFFFFFFCu)) {
                        i_out = i_out | i_inte_last;
                        inter[2 * (pcnt - 1)] = (int32_t)c;
                        inter[2 * (pcnt - 1) + 1] = i_out;
                    }
                    inter[2 * pcnt] = (int32_t)c;
                    inter[2 * pcnt + 1] = i_out;
                    i_inte_last = i_out;
                    ++pcnt;
                }
                if (t_solve < 2.f) {
                    const float *cv = P + side * 4;
                    if (type == T_LINE) { /* MI1:379-385 */
                        float a = cv[1] - cv[0];
                        a = (a != 0.0f ? (1.0f / a) : 0.0f);
                        float v = (cst - cv[0]) * a;
                        v = (v < t_min) ? t_min : v;     /* GLSL max(x,y): y if x<y else x */
                        t_solve = (t1_ms < v) ? t1_ms : v; /* GLSL min(x,y): y if y<x else x */
                    } else if (!g_full_rvg && (type == T_QUADRIC || type == T_ARC)) {
                        /* TODO arms: t_solve stays 0 */
                    } else { /* MI1:392-436 (every other type falls here; only CUBIC evaluates) */
                        float t0 = t_min, t1 = t1_ms;
                        float vt0 = interp_general(type, t0, cv, 0.0f);
                        t_solve = t0;
                        if (vt0 != cst) {
                            float raw_t0 = t0, last_vtm = 0.f;
                            for (int j = 0; j < CUBIC_ITERATION_NUMBER; ++j) {
                                float tm = (t0 + t1) * 0.5f;
                                float vtm = interp_general(type, tm, cv, 0.0f);
                                t_solve = tm; last_vtm = vtm;
                                if ((int32_t)(f2u(canon(vtm - cst)) ^ f2u(canon(vt0 - cst))) >= 0) { t0 = tm; vt0 = vtm; }
                                else t1 = tm;
                            }
                            if (fabsf(last_vtm - cst) > 1.f) t_solve = raw_t0;
                        }
                    }
                }
                P[8 + side] = u2f((f2u(canon(t_solve)) & 0xFFFFFFFCu) | (uint32_t)side); /* MI1:440 */
            }
            t0_ms = t1_ms; p0x = p1x; p0y = p1y; /* MI1:442-443 */
        }
    }
}

/* ------------------------------------------------------------------ GF:90-246 */
static inline void curve_interp2(uint32_t type, float t, const float *cx, const float *cy, float *ox, float *oy) {
    /* GF:59-87: default result is cv0 */
    if (type == T_LINE) { *ox = lerpf(cx[0], cx[1], t); *oy = lerpf(cy[0], cy[1], t); return; }
    if (type == T_CUBIC) {
        *ox = interp_general(T_CUBIC, t, cx, 0.f);
        *oy = interp_general(T_CUBIC, t, cy, 0.f);
        return;
    }
    if (g_full_rvg && (type == T_QUADRIC || type == T_ARC)) { /* f-1 */
        *ox = interp_general(type, t, cx, 0.f);
        *oy = interp_general(type, t, cy, 0.f);
        return;
    }
    *ox = cx[0]; *oy = cy[0];
}

void orc_gen_fragment(int32_t nf, uint32_t n_paths, const int32_t *inter,
                      const uint32_t *curve_path, const uint32_t *curve_pos_map,
                      const uint32_t *curve_type, const float *tpos,
                      int width, int height,
                      int32_t *key, int32_t *idx, int32_t *path, int32_t *wind, int32_t *seg) {
    if (nf <= 0) { /* no thread runs in the reference; define the table as all-zero */
        for (uint32_t j = 0; j <= n_paths; ++j) seg[j] = 0;
        return;
    }
#pragma omp parallel for schedule(static)
    for (int32_t f = 0; f < nf; ++f) {
        int32_t c0 = inter[2 * f], b0 = inter[2 * f + 1];
        int32_t c1, b1;
        if (f + 1 != nf) { c1 = inter[2 * (f + 1)]; b1 = inter[2 * (f + 1) + 1]; }
        else { c1 = -1; b1 = 0x3f800000; } /* GF:99 */
        float t0 = u2f((uint32_t)b0 & 0xFFFFFFFCu);
        float t1 = u2f((uint32_t)b1 & 0xFFFFFFFCu);
        t0 = (t0 < 0.0f) ? 0.0f : t0; /* GF:103-104 max(t,0) */
        t1 = (t1 < 0.0f) ? 0.0f : t1;
        uint32_t pidx = curve_path[c0];
        int32_t yx = (int32_t)0xFFFEFFFEu; /* GF:111 */
        if (c0 != c1) t1 = 1.0f;
        int wn = 0;
        if (t0 < t1) {
            uint32_t type = curve_type[c0];
            uint32_t po = curve_pos_map[c0];
            float cx[4] = {0, 0, 0, 0}, cy[4] = {0, 0, 0, 0};
            uint32_t np = (type == T_LINE) ? 2 : (type == T_QUADRIC) ? 3 : (type == T_CUBIC) ? 4 : 0; /* GF:132-156 */
            if (g_full_rvg && type == T_ARC) np = 3;
            for (uint32_t k = 0; k < np; ++k) { cx[k] = tpos[2 * (po + k)]; cy[k] = tpos[2 * (po + k) + 1]; }
            if (g_full_rvg && type == T_ARC) cx[3] = cy[3] = g_curve_weight ? g_curve_weight[c0] : 1.0f;
            float pfx, pfy, plx, ply;
            curve_interp2(type, t0, cx, cy, &pfx, &pfy);
            curve_interp2(type, t1, cx, cy, &plx, &ply);
            int raw_x = float2int_rd(((pfx + plx) * 0.5f) / FRAG_SIZE) * FRAG_SIZE; /* GF:169-170 */
            int raw_y = float2int_rd(((pfy + ply) * 0.5f) / FRAG_SIZE) * FRAG_SIZE;
            int pos_x = imin(imax(raw_x, -FRAG_SIZE), (int)(((uint32_t)width & 0xFFFFFFFEu) + FRAG_SIZE));
            int pos_y = imin(imax(raw_y, -FRAG_SIZE), (int)(((uint32_t)height & 0xFFFFFFFEu) + FRAG_SIZE));
            int32_t y_shift = (int32_t)((uint32_t)(pos_y + 0x7FFF) << 16);
            int32_t x_shift = ((pos_x + 0x7FFF) & 0xFFFF);
            if ((uint32_t)raw_y < (uint32_t)height) yx = y_shift | x_shift; /* GF:185-187 */
            int wn_y = pos_y + 1;
            if (pfy == ply) wn = 0; /* GF:190-199 */
            else {
                if (pfy < (float)wn_y && (float)wn_y <= ply) wn = -1;
                else if (ply < (float)wn_y && (float)wn_y <= pfy) wn = 1;
            }
        }
        key[f] = yx; idx[f] = f; path[f] = (int32_t)pidx; wind[f] = wn; /* GF:221-224 */
    }
    /* segment table GF:226-244 (single writer per slot; done serially here) */
    for (int32_t f = 0; f < nf; ++f) {
        int32_t c0 = inter[2 * f];
        int32_t c1 = (f + 1 != nf) ? inter[2 * (f + 1)] : -1;
        if (c0 != c1) {
            uint32_t p0 = curve_path[c0];
            uint32_t p1 = (c1 >= 0) ? curve_path[c1] : n_paths;
            for (uint32_t j = p0 + 1; j <= p1; ++j) seg[j] = f + 1;
        }
        if (f == 0) {
            uint32_t p0 = curve_path[c0];
            for (uint32_t j = 0; j <= p0; ++j) seg[j] = 0;
        }
    }
}

/* ------------------------------------------------------------------ SORT */
typedef struct { int32_t k, v; } kv_t;
static int kv_cmp(const void *a, const void *b) { /* SORT:69, signed */
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    return 0;
}
void orc_seg_sort(uint32_t n_paths, const int32_t *seg, int32_t *key, int32_t *idx) {
#pragma omp parallel
    {
        kv_t *buf = NULL; size_t cap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t p = 0; p < (int64_t)n_paths; ++p) {
            int32_t b = seg[p], e = seg[p + 1];
            int32_t n = e - b;
            if (n <= 1) continue; /* SORT:41-43 */
            if ((size_t)n > cap) { free(buf); cap = (size_t)n * 2; buf = (kv_t *)malloc(cap * sizeof(kv_t)); }
            for (int32_t i = 0; i < n; ++i) { buf[i].k = key[b + i]; buf[i].v = idx[b + i]; }
            qsort(buf, (size_t)n, sizeof(kv_t), kv_cmp);
            for (int32_t i = 0; i < n; ++i) { key[b + i] = buf[i].k; idx[b + i] = buf[i].v; }
        }
        free(buf);
    }
}

void orc_seg_sort_literal(uint32_t n_paths, const int32_t *seg, int32_t *key, int32_t *idx) {
    for (uint32_t p = 0; p < n_paths; ++p) { /* SORT:53-97: segment_size rounds of odd-even exchange */
        int32_t begin = seg[p], n = seg[p + 1] - seg[p];
        if (n <= 1) continue;
        int flag = 1;
        for (int32_t r = 0; r < n; ++r) {
            flag = 1 - flag;
            for (int32_t l = flag; l + 1 < n; l += 2) {
                int32_t li = begin + l, ri = li + 1;
                int32_t kl = key[li], kr = key[ri], vl = idx[li], vr = idx[ri];
                if (kl > kr || (kl == kr && vl > vr)) { key[li] = kr; key[ri] = kl; idx[li] = vr; idx[ri] = vl; }
            }
        }
    }
}

/* ------------------------------------------------------------------ SHUF:17-27 */
void orc_shuffle(int32_t nf, const int32_t *idx, const int32_t *wind, int32_t *out) {
    for (int32_t i = 0; i < nf; ++i) out[i] = wind[idx[i]];
}

/* ------------------------------------------------------------------ MARK:22-94 */
void orc_mark(int32_t nf, const int32_t *key, const int32_t *path, const int32_t *wn,
              const uint32_t *fill_rule, int width, int height, int32_t *flags) {
#pragma omp parallel for schedule(static)
    for (int32_t f = 0; f < nf; ++f) {
        int frag = 0, span = 0;
        int32_t yx1 = key[f];
        int x1 = (yx1 & 0xFFFF) - 0x7FFF;
        int y1 = ((yx1 >> 16) & 0xFFFF) - 0x7FFF;
        int oob = (x1 < 0 || y1 < 0 || x1 >= width || y1 >= height);
        if (f == 0) {
            frag = oob ? 0 : 1; span = 0; /* MARK:44-51 */
        } else {
            int32_t p0 = path[f - 1] & 0x3FFFFFFF, p1 = path[f] & 0x3FFFFFFF;
            uint32_t rule = fill_rule[p1];
            int32_t yx0 = key[f - 1];
            int x0 = (yx0 & 0xFFFF) - 0x7FFF;
            int y0 = ((yx0 >> 16) & 0xFFFF) - 0x7FFF;
            if (oob) frag = 0;
            else if (p0 != p1 || yx0 != yx1) frag = 1;
            else frag = 0;
            int32_t w = wn[f];
            int wn_flag = ((rule == 0) && (w != 0)) || ((rule == 1) && ((w & 1) != 0));
            span = (y0 == y1 && ((x0 + FRAG_SIZE) < x1) && p0 == p1 && wn_flag) ? 1 : 0;
        }
        flags[f] = frag;
        flags[nf + f] = span;
    }
}

/* ------------------------------------------------------------------ GEN:33-103 */
void orc_emit(int32_t nf, const int32_t *key, const int32_t *path, const int32_t *flags,
              const int32_t *scan3, const uint32_t *fill_info, int32_t n_out_frag, int32_t *out) {
#pragma omp parallel for schedule(static)
    for (int32_t f = 0; f < nf; ++f) {
        int frag_flag = flags[f];
        int frag_index = scan3[f + 1];
        int span_flag = flags[nf + f];
        int span_index = scan3[nf + f + 1] - n_out_frag;
        int nfb = frag_index - frag_flag, nsb = span_index - span_flag;
        if (frag_flag != 0) {
            int oi = nfb + nsb;
            int32_t raw = key[f];
            int rx = (raw & 0xFFFF) - 0x7FFF;
            int ry = ((raw >> 16) & 0xFFFF) - 0x7FFF;
            int32_t p = path[f] & 0x3FFFFFFF;
            out[4 * oi + 0] = (int32_t)(((uint32_t)ry << 16) | (uint32_t)rx);
            out[4 * oi + 1] = 2;
            out[4 * oi + 2] = (int32_t)fill_info[p];
            out[4 * oi + 3] = frag_index;
        }
        if (span_flag != 0) {
            int oi = nfb + nsb + frag_flag;
            int32_t k0 = key[f - 1], k1 = key[f];
            int32_t p = path[f] & 0x3FFFFFFF;
            int y0 = ((k0 >> 16) & 0xFFFF) - 0x7FFF;
            int x0 = (k0 & 0xFFFF) - 0x7FFF + FRAG_SIZE;
            int x1 = (k1 & 0xFFFF) - 0x7FFF;
            x0 = imax(0, x0);
            out[4 * oi + 0] = (int32_t)(((uint32_t)y0 << 16) | (uint32_t)x0);
            out[4 * oi + 1] = x1 - x0;
            out[4 * oi + 2] = (int32_t)fill_info[p];
            out[4 * oi + 3] = 0;
        }
    }
}

/* ------------------------------------------------------------------ stage 5 */
/* SURVEY section 8 f-3 (NOT in the reference, which composites opaque: blendEnable = VK_FALSE, SR:893-895): with orc_set_blend(1)
 * the records are composited "source over" in their order — which is path order — onto the white clear colour:
 * alpha 255 overwrites (as the reference does), alpha 0 leaves the pixel alone, anything between blends each colour
 * channel as (src * a + dst * (255 - a) + 127) / 255 in integers; the frame stays opaque (alpha 255). Mirrored bit for
 * bit by SLPR_FLAG_BLEND (csrc/raster.cuh). Off by default. */
static int g_blend = 0;
void orc_set_blend(int on) { g_blend = on; }
static inline uint32_t blend_over(uint32_t dst, uint32_t src) {
    uint32_t a = src >> 24;
    if (a == 255u) return src;
    if (a == 0u) return dst;
    uint32_t out = 0xFF000000u;
    for (int ch = 0; ch < 3; ++ch) {
        uint32_t s = (src >> (8 * ch)) & 0xFFu, d = (dst >> (8 * ch)) & 0xFFu;
        out |= ((s * a + d * (255u - a) + 127u) / 255u) << (8 * ch);
    }
    return out;
}

void orc_fill(int64_t n_records, const int32_t *rec, int width, int height, uint8_t *rgba) {
    memset(rgba, 0xFF, (size_t)width * (size_t)height * 4); /* SR:622 clear white */
    if (g_blend) {
        for (int64_t r = 0; r < n_records; ++r) {
            int32_t yx = rec[4 * r], w = rec[4 * r + 1];
            uint32_t col = (uint32_t)rec[4 * r + 2];
            int X = yx & 0xFFFF, Y = yx >> 16;
            int xa = imax(X, 0), xb = imin(X + w, width);
            for (int row = Y; row <= Y + 1; ++row) {
                if (row < 0 || row >= height) continue;
                uint8_t *dst = rgba + ((size_t)(height - 1 - row) * (size_t)width) * 4;
                for (int x = xa; x < xb; ++x) {
                    uint32_t d;
                    memcpy(&d, dst + 4 * (size_t)x, 4);
                    d = blend_over(d, col);
                    memcpy(dst + 4 * (size_t)x, &d, 4);
                }
            }
        }
        return;
    }
    for (int64_t r = 0; r < n_records; ++r) {
        int32_t yx = rec[4 * r], w = rec[4 * r + 1];
        uint32_t col = (uint32_t)rec[4 * r + 2];
        int X = yx & 0xFFFF;  /* VERT:27 */
        int Y = yx >> 16;     /* VERT:27 arithmetic shift */
        int xa = imax(X, 0), xb = imin(X + w, width);
        for (int row = Y; row <= Y + 1; ++row) { /* line at y=Y+1, width 2 -> rows Y, Y+1 */
            if (row < 0 || row >= height) continue;
            uint8_t *dst = rgba + ((size_t)(height - 1 - row) * (size_t)width) * 4; /* VERT:44 y flip */
            for (int x = xa; x < xb; ++x) memcpy(dst + 4 * (size_t)x, &col, 4); /* R in low byte, VERT:8-10 */
        }
    }
}

/* ------------------------------------------------------------------ host glue */
uint32_t orc_quantise_colour(const float *rgba, float opacity) { /* SR:107-118 */
    float c[4] = {rgba[0], rgba[1], rgba[2], rgba[3]};
    c[3] = c[3] * opacity;
    uint8_t col[4];
    for (int i = 0; i < 4; ++i) { float v = c[i] * 255.0f; col[i] = (uint8_t)v; }
    uint32_t w = (uint32_t)col[0] | ((uint32_t)col[1] << 8) | ((uint32_t)col[2] << 16) | ((uint32_t)col[3] << 24);
    return (w & 0xFF000000u) ? w : 0;
}

static double now_ms(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

orc_frame *orc_render(uint32_t n_points, const float *pos, const uint32_t *pos_path,
                      uint32_t n_curves, const uint32_t *curve_pos_map,
                      const uint32_t *curve_type, const uint32_t *curve_path,
                      uint32_t n_paths, const uint32_t *fill_rule, const uint32_t *fill_info,
                      const float *rows, int width, int height, int do_fill) {
    orc_frame *F = (orc_frame *)calloc(1, sizeof(orc_frame));
    F->n_points = n_points; F->n_curves = n_curves; F->n_paths = n_paths;
    F->width = width; F->height = height;
    double t = now_ms(), t2;
    F->tpos = (float *)malloc(sizeof(float) * 2 * (size_t)(n_points + 1));
    F->path_visible = (int32_t *)calloc((size_t)n_paths + 1, 4);
    orc_transform(n_points, pos, pos_path, rows, width, height, F->tpos, F->path_visible);
    t2 = now_ms(); F->ms[0] = t2 - t; t = t2;
    F->cut_cache = (float *)malloc(sizeof(float) * 5 * (size_t)(n_curves + 1));
    F->curve_count = (int32_t *)malloc(4 * (size_t)(n_curves + 1));
    F->curve_offset = (int32_t *)malloc(4 * (size_t)(n_curves + 1));
    orc_monotonize_count(n_curves, curve_type, curve_pos_map, curve_path, F->tpos, F->path_visible,
                         width, height, F->cut_cache, F->curve_count);
    t2 = now_ms(); F->ms[1] = t2 - t; t = t2;
    orc_exclusive_scan(n_curves, F->curve_count, F->curve_offset);
    int32_t nf = F->curve_offset[n_curves];
    F->n_fragments = nf;
    t2 = now_ms(); F->ms[2] = t2 - t; t = t2;
    size_t n1 = (size_t)nf + 1;
    F->inter = (int32_t *)malloc(8 * n1);
    orc_intersect(n_curves, curve_type, curve_pos_map, curve_path, F->tpos, F->path_visible,
                  F->cut_cache, F->curve_offset, width, height, F->inter);
    t2 = now_ms(); F->ms[3] = t2 - t; t = t2;
    F->key = (int32_t *)malloc(4 * n1); F->idx = (int32_t *)malloc(4 * n1);
    F->path = (int32_t *)malloc(4 * n1); F->wind = (int32_t *)malloc(4 * n1);
    F->seg = (int32_t *)calloc((size_t)n_paths + 1, 4);
    orc_gen_fragment(nf, n_paths, F->inter, curve_path, curve_pos_map, curve_type, F->tpos,
                     width, height, F->key, F->idx, F->path, F->wind, F->seg);
    t2 = now_ms(); F->ms[4] = t2 - t; t = t2;
    F->skey = (int32_t *)malloc(4 * n1); F->sidx = (int32_t *)malloc(4 * n1);
    memcpy(F->skey, F->key, 4 * (size_t)nf); memcpy(F->sidx, F->idx, 4 * (size_t)nf);
    orc_seg_sort(n_paths, F->seg, F->skey, F->sidx);
    t2 = now_ms(); F->ms[5] = t2 - t; t = t2;
    F->swind = (int32_t *)malloc(4 * n1);
    F->wn = (int32_t *)malloc(4 * n1);
    orc_shuffle(nf, F->sidx, F->wind, F->swind);
    orc_exclusive_scan(nf, F->swind, F->wn);
    F->flags = (int32_t *)malloc(8 * n1);
    F->scan3 = (int32_t *)malloc(8 * n1 + 4);
    orc_mark(nf, F->skey, F->path, F->wn, fill_rule, width, height, F->flags);
    orc_exclusive_scan(2 * (int64_t)nf, F->flags, F->scan3);
    F->n_out_frag = F->scan3[nf];                       /* SR:578 */
    F->n_span = F->scan3[2 * (size_t)nf] - F->n_out_frag; /* SR:579-580 */
    size_t no = (size_t)F->n_out_frag + (size_t)F->n_span;
    F->records = (int32_t *)malloc(16 * (no + 1));
    orc_emit(nf, F->skey, F->path, F->flags, F->scan3, fill_info, F->n_out_frag, F->records);
    t2 = now_ms(); F->ms[6] = t2 - t; t = t2;
    if (do_fill) {
        F->rgba = (uint8_t *)malloc((size_t)width * (size_t)height * 4);
        orc_fill((int64_t)no, F->records, width, height, F->rgba);
    }
    t2 = now_ms(); F->ms[7] = t2 - t;
    return F;
}

void orc_frame_free(orc_frame *F) {
    if (!F) return;
    free(F->tpos); free(F->path_visible); free(F->cut_cache); free(F->curve_count); free(F->curve_offset);
    free(F->inter); free(F->key); free(F->idx); free(F->path); free(F->wind); free(F->seg);
    free(F->skey); free(F->sidx); free(F->swind); free(F->wn); free(F->flags); free(F->scan3);
    free(F->records); free(F->rgba); free(F);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
