// geom.cuh — curve-space kernels: transform + visibility, monotonic subdivision + counting,
// scanline/grid intersection walk, fragment generation. All fp32 arithmetic is written with
// explicit round-to-nearest intrinsics (no FMA contraction) so that results are bit-identical
// to the reference shaders evaluated without contraction.
#pragma once
#include "common.cuh"

namespace slpr {

// ------------------------------------------------------------------------------------------------
// K1: transform_pos.comp:33-85. One thread per point, grid-stride. The reference ORs the region
// nibble into path_visible non-atomically (a data race, TP:72-80); here the OR is first reduced
// over the lanes of the warp that hit the same path (points of a path are contiguous), then one
// atomicOr per distinct path per warp.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_transform(const FrameParams *__restrict__ P, uint32_t n_points,
                                                   const float2 *__restrict__ pos,
                                                   const uint32_t *__restrict__ pos_path,
                                                   float2 *__restrict__ tpos, int *__restrict__ path_visible) {
    const float m0x = P->rows[0], m0y = P->rows[1], m0z = P->rows[2], m0w = P->rows[3];
    const float m1x = P->rows[4], m1y = P->rows[5], m1z = P->rows[6], m1w = P->rows[7];
    const float m3x = P->rows[12], m3y = P->rows[13], m3z = P->rows[14], m3w = P->rows[15];
    const float w = (float)P->width, h = (float)P->height;  // TP:8 (floats), SR.cpp:1153-1154
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n_points + 31u) & ~31u;  // keep whole warps in the loop for the shuffles
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool live = i < n_points;
        uint32_t flag = 0, pidx = 0xFFFFFFFFu;
        if (live) {
            const float2 p = pos[i];
            // dot(vec4(x,y,0,1), m) evaluated left to right (TP:41-46)
            float ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, m0x), __fmul_rn(p.y, m0y)), __fmul_rn(0.0f, m0z)),
                                 __fmul_rn(1.0f, m0w));
            float oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, m1x), __fmul_rn(p.y, m1y)), __fmul_rn(0.0f, m1z)),
                                 __fmul_rn(1.0f, m1w));
            float ow = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, m3x), __fmul_rn(p.y, m3y)), __fmul_rn(0.0f, m3z)),
                                 __fmul_rn(1.0f, m3w));
            ox = __fdiv_rn(ox, ow);  // TP:53-54
            oy = __fdiv_rn(oy, ow);
            const int xf = ox < 0 ? 0 : (ox < w ? 1 : 2);  // TP:67-68
            const int yf = oy < 0 ? 0 : (oy < h ? 1 : 2);
            switch ((yf << 4) | xf) {  // TP:71-82
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
            pidx = pos_path[i];
            tpos[i] = make_float2(ox, oy);
        }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, pidx);
        const uint32_t red = __reduce_or_sync(peers, flag);
        if (live && (lane_id() == (uint32_t)(__ffs(peers) - 1))) atomicOr(&path_visible[pidx], (int)red);
    }
}

// ------------------------------------------------------------------------------------------------
// helpers shared by K2 and K4
// ------------------------------------------------------------------------------------------------
// make_intersection_0.comp:76-129
__device__ __forceinline__ void solve_quad(float a, float b, float c, float &r0, float &r1) {
    if (a == 0) {
        const float x = __fdiv_rn(-c, b);
        r0 = x; r1 = x;
        return;
    }
    const float A = a, B = __fmul_rn(b, 0.5f), C = c;
    float tx = 0.f, ty = 0.f;
    const float R = __fsub_rn(__fmul_rn(B, B), __fmul_rn(A, C));
    if (R > 0.0f) {
        const float SR = __fsqrt_rn(R);
        if (B > 0.0f) {
            const float TB = __fadd_rn(B, SR);
            tx = __fdiv_rn(-C, TB); ty = __fdiv_rn(-TB, A);
        } else {
            const float TB = __fadd_rn(-B, SR);
            tx = __fdiv_rn(TB, A); ty = __fdiv_rn(C, TB);
        }
    }
    r0 = tx; r1 = ty;
}

// make_intersection_0.comp:186-221 == make_intersection_1.comp:174-213 (one axis)
__device__ __forceinline__ int cut_range(int dim, int &b, int &e) {
    const int cmin = 0, cmax = (int)((uint32_t)dim & 0xFFFFFFFEu) + FRAG_SIZE;
    if ((b < cmin && e < cmin) || (b > cmax && e > cmax) || (b > e)) return 0;
    b = min(max(b, cmin), cmax);
    e = min(max(e, cmin), cmax);
    return max((e - b) / FRAG_SIZE + 1, 0);
}

struct CurvePts {
    float x[4], y[4];
};

__device__ __forceinline__ void load_points(uint32_t type, uint32_t po, const float2 *__restrict__ tpos, CurvePts &c) {
    const uint32_t n = type & 7u;  // MI0:252
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        float2 p = make_float2(0.f, 0.f);  // uninitialised shared memory in the reference; never consumed
        if (i < n) p = tpos[po + i];
        c.x[i] = p.x; c.y[i] = p.y;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: make_intersection_0.comp:226-410. One thread per curve: monotonic cut parameters (<=4),
// literal partial insertion sort (including the `float t2 = q3;` slip at MI0:340), and the number
// of 2-px grid crossings per monotone piece. Band mode (new): a curve whose control-point box
// misses the band by more than one pixel emits nothing (exact).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_monotonize_count(const FrameParams *__restrict__ P, uint32_t n_curves,
                                                          const uint32_t *__restrict__ curve_type,
                                                          const uint32_t *__restrict__ curve_pos_map,
                                                          const uint32_t *__restrict__ curve_path,
                                                          const float2 *__restrict__ tpos,
                                                          const int *__restrict__ path_visible,
                                                          float *__restrict__ cut_cache, int *__restrict__ count) {
    const int width = P->width, height = P->height;
    const bool cull = P->cull != 0;
    // Only interior band edges cull: beyond the frame edges the reference still emits (invalid-key)
    // fragments whose winding deltas pair up across curves of a path (sampling rows -1 and H'+3).
    const float band_lo = (P->band_y0 > 0) ? (float)(P->band_y0 - 1) : -3.0e38f;
    const float band_hi = (P->band_y1 < P->height) ? (float)(P->band_y1 + 1) : 3.0e38f;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_curves; c += gridDim.x * blockDim.x) {
        const uint32_t type = curve_type[c];
        CurvePts cp;
        load_points(type, curve_pos_map[c], tpos, cp);
        uint32_t n_cuts = 0;
        const bool visible = !path_invisible(path_visible[curve_path[c]]);  // MI0:260-261
        float tq[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        if (visible) {
            if (type == T_CUBIC) {  // MI0:267-308 (LINE has no cuts; QUADRIC/ARC are TODO arms)
#pragma unroll
                for (int ax = 0; ax < 2; ++ax) {
                    const float x0 = ax ? cp.y[0] : cp.x[0], x1 = ax ? cp.y[1] : cp.x[1];
                    const float x2 = ax ? cp.y[2] : cp.x[2], x3 = ax ? cp.y[3] : cp.x[3];
                    float r0 = 0.f, r1 = 0.f;
                    const float a = __fadd_rn(__fmul_rn(3.0f, __fsub_rn(x1, x2)), __fsub_rn(x3, x0));
                    const float b = __fmul_rn(2.0f, __fadd_rn(__fsub_rn(x0, x1), __fsub_rn(x2, x1)));
                    const float cc = __fsub_rn(x1, x0);
                    solve_quad(a, b, cc, r0, r1);
                    if (r0 > 0.0f && r0 < 1.0f) { tq[n_cuts] = r0; ++n_cuts; }
                    if (r1 > 0.0f && r1 < 1.0f && r1 != r0) { tq[n_cuts] = r1; ++n_cuts; }
                }
            }
            q0 = tq[0]; q1 = tq[1]; q2 = tq[2]; q3 = tq[3];  // MI0:310-313
            if (n_cuts >= 2) {                                // MI0:315-322
                const float t1 = q1, t0 = q0;
                if (t1 < t0) { q1 = t0; q0 = t1; }
            }
            if (n_cuts >= 3) {  // MI0:323-337
                const float t2 = q2, t1 = q1;
                if (t2 < t1) {
                    q2 = t1;
                    const float t0 = q0;
                    if (t2 < t0) { q1 = t0; q0 = t2; } else { q1 = t2; }
                }
            }
            // MI0:338-359: the reference compares q3 with itself (`float t2 = q3;`), so a fourth cut
            // is never inserted. Kept: with t3 == t2 the branch `t3 < t2` is dead.
        }
        tq[0] = q0; tq[1] = q1; tq[2] = q2; tq[3] = q3;  // MI0:363-366
        cut_cache[5 * c + 0] = q0;                      // MI0:368-372
        cut_cache[5 * c + 1] = q1;
        cut_cache[5 * c + 2] = q2;
        cut_cache[5 * c + 3] = q3;
        cut_cache[5 * c + 4] = u2f(n_cuts);
        if (visible) { tq[n_cuts] = 1.f; ++n_cuts; }  // MI0:374-377

        bool culled = false;
        if (cull) {
            const uint32_t np = type & 7u;
            float ymin = cp.y[0], ymax = cp.y[0];
            for (uint32_t i = 1; i < 4; ++i)
                if (i < np) { ymin = fminf(ymin, cp.y[i]); ymax = fmaxf(ymax, cp.y[i]); }
            culled = (ymax < band_lo) || (ymin >= band_hi);
        }

        float p0x = cp.x[0], p0y = cp.y[0];
        int pcnt = 0;
        for (uint32_t i = 0; i < n_cuts; ++i) {  // MI0:383-408
            const float t1 = tq[i];
            const float p1x = interp_general(type, t1, cp.x[0], cp.x[1], cp.x[2], cp.x[3], 1.0f);
            const float p1y = interp_general(type, t1, cp.y[0], cp.y[1], cp.y[2], cp.y[3], 1.0f);
            // get_xy_begin_end, MI0:167-183 (floor)
            const float xlo = (p0x <= p1x) ? p0x : p1x, xhi = (p0x <= p1x) ? p1x : p0x;
            const float ylo = (p0y <= p1y) ? p0y : p1y, yhi = (p0y <= p1y) ? p1y : p0y;
            int xb = f2i(__fmul_rn(floorf(__fdiv_rn(xlo, 2.0f)), 2.0f)) + FRAG_SIZE;
            int xe = f2i(__fmul_rn(floorf(__fdiv_rn(xhi, 2.0f)), 2.0f));
            int yb = f2i(__fmul_rn(floorf(__fdiv_rn(ylo, 2.0f)), 2.0f)) + FRAG_SIZE;
            int ye = f2i(__fmul_rn(floorf(__fdiv_rn(yhi, 2.0f)), 2.0f));
            const int nx = cut_range(width, xb, xe);
            const int ny = cut_range(height, yb, ye);
            pcnt += 1 + nx + ny;
            p0x = p1x; p0y = p1y;
        }
        count[c] = culled ? 0 : pcnt;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: make_intersection_1.comp:217-447. The reference walks one curve per thread: per monotone piece
// it merges the x- and y-grid crossings in parameter order; lines in closed form, cubics by a
// 24-step bisection whose bracket starts at the previously emitted crossing on the same axis — so
// the walk along a piece is inherently sequential if the emitted t must be bit-identical.
//
// B200 formulation (the arithmetic per step is the reference's, operation for operation):
//   * the nested loops (pieces x crossings) are flattened into a per-lane state machine;
//   * curves are handed out dynamically: each warp owns a chunk of curve indices (one global
//     atomicAdd per WALK_CHUNK curves) and a lane that finishes its curve takes the next index by
//     ballot, so no lane idles while work remains;
//   * each outer iteration first advances every lane to its next crossing solve (emission, side
//     selection, piece set-up and curve fetch are the cheap, divergent part) and then runs the
//     24-step bisection — 70 % of all instructions — once, with the whole warp converged.
// Writes (curve, tbits) records at the curve's scanned offset.
// ------------------------------------------------------------------------------------------------
constexpr int WALK_THREADS = 128;
constexpr int WALK_CHUNK = 32;

__global__ void __launch_bounds__(WALK_THREADS) k_intersect(const FrameParams *__restrict__ P, uint32_t n_curves,
                                                            const uint32_t *__restrict__ curve_type,
                                                            const uint32_t *__restrict__ curve_pos_map,
                                                            const float2 *__restrict__ tpos,
                                                            const float *__restrict__ cut_cache,
                                                            const int *__restrict__ offsets,
                                                            const FrameCounters *__restrict__ ctr, int capacity,
                                                            int2 *__restrict__ inter, int *__restrict__ work_counter) {
    if (ctr->n_fragments > capacity) return;  // overflow: the host re-renders with larger buffers
    const int width = P->width, height = P->height;
    const uint32_t lane = lane_id(), lt = lanemask_lt();
    enum { NEED_CURVE = 0, NEED_PIECE = 1, WALK = 2, DONE = 3 };
    int state = NEED_CURVE;
    uint32_t wl_next = 0, wl_end = 0;  // this warp's chunk of curve indices (warp-uniform)
    bool exhausted = false;

    // per-curve state
    uint32_t c = 0, type = 0, n_cuts = 0, piece = 0;
    CurvePts cp;
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    float t0_ms = 0.f, p0x = 0.f, p0y = 0.f;
    int pcnt = 0;
    // per-piece state
    float t1_ms = 0.f, p1x = 0.f, p1y = 0.f, x = 0.f, y = 0.f, dx = 0.f, dy = 0.f, tx = 0.f, ty = 0.f;
    int n_x = 0, n_y = 0, n_loop = 0, it = 0, i_inte_last = 0;
    // pending solve
    bool pending = false;
    int side = 0;
    float cst = 0.f, t_min = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { cp.x[i] = 0.f; cp.y[i] = 0.f; }

    // MI1:440-443: store the tagged parameter on its side; the piece ends after n_loop+1 steps
#define WALK_FINISH_STEP(T_SOLVE)                                                              \
    do {                                                                                       \
        const float tagged_ = u2f((f2u(T_SOLVE) & 0xFFFFFFFCu) | (uint32_t)side);              \
        if (side) ty = tagged_; else tx = tagged_;                                             \
        ++it;                                                                                  \
        if (it == n_loop) {                                                                    \
            t0_ms = t1_ms; p0x = p1x; p0y = p1y;                                               \
            ++piece;                                                                           \
            state = (piece < n_cuts) ? NEED_PIECE : NEED_CURVE;                                \
        }                                                                                      \
    } while (0)

    while (true) {
        // =========== advance every lane to its next solve ===========
        while (true) {
            const uint32_t busy = __ballot_sync(0xFFFFFFFFu, !pending && state != DONE);
            if (!busy) break;
            // ---- hand out curves to the lanes that need one
            while (true) {
                const uint32_t need = __ballot_sync(0xFFFFFFFFu, state == NEED_CURVE);
                if (!need) break;
                if (wl_next >= wl_end) {
                    if (exhausted) {
                        if (state == NEED_CURVE) state = DONE;
                        break;
                    }
                    uint32_t base = 0;
                    if (lane == 0) base = (uint32_t)atomicAdd(work_counter, WALK_CHUNK);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n_curves) { exhausted = true; continue; }
                    wl_next = base;
                    wl_end = min(base + (uint32_t)WALK_CHUNK, n_curves);
                }
                const uint32_t avail = wl_end - wl_next;
                const uint32_t rank = __popc(need & lt);
                if (state == NEED_CURVE && rank < avail) {
                    c = wl_next + rank;
                    pcnt = offsets[c];
                    if (offsets[c + 1] != pcnt) {  // count 0 = invisible or band-culled curve: take another one
                        type = curve_type[c];
                        load_points(type, curve_pos_map[c], tpos, cp);
                        q0 = cut_cache[5 * c + 0]; q1 = cut_cache[5 * c + 1];
                        q2 = cut_cache[5 * c + 2]; q3 = cut_cache[5 * c + 3];
                        // count > 0 implies the path is visible (MI0:374-377), so MI1:257-260 appends t = 1
                        n_cuts = f2u(cut_cache[5 * c + 4]) + 1u;  // MI1:255
                        piece = 0;
                        t0_ms = 0.f; p0x = cp.x[0]; p0y = cp.y[0];
                        state = NEED_PIECE;
                    }
                }
                wl_next += min((uint32_t)__popc(need), avail);
            }
            // ---- set up the next monotone piece (MI1:266-308)
            if (state == NEED_PIECE) {
                const bool last = piece + 1 == n_cuts;  // the appended t = 1
                t1_ms = last ? 1.f : (piece == 0) ? q0 : (piece == 1) ? q1 : (piece == 2) ? q2 : q3;
                p1x = interp_general(type, t1_ms, cp.x[0], cp.x[1], cp.x[2], cp.x[3], 0.0f);
                p1y = interp_general(type, t1_ms, cp.y[0], cp.y[1], cp.y[2], cp.y[3], 0.0f);
                // MI1:271-276: tag t1 in its two mantissa LSBs
                if (floorf(p1x) == p1x) t1_ms = u2f((f2u(t1_ms) & 0xFFFFFFFCu) | 2u);
                else t1_ms = u2f(f2u(t1_ms) | 3u);
                // get_xy_begin_end_delta, MI1:150-170 (float2int_rd)
                const bool xfwd = p0x <= p1x, yfwd = p0y <= p1y;
                int xb = float2int_rd(__fdiv_rn(xfwd ? p0x : p1x, 2.0f)) * FRAG_SIZE + FRAG_SIZE;
                int xe = float2int_rd(__fdiv_rn(xfwd ? p1x : p0x, 2.0f)) * FRAG_SIZE;
                int yb = float2int_rd(__fdiv_rn(yfwd ? p0y : p1y, 2.0f)) * FRAG_SIZE + FRAG_SIZE;
                int ye = float2int_rd(__fdiv_rn(yfwd ? p1y : p0y, 2.0f)) * FRAG_SIZE;
                dx = xfwd ? 2.0f : -2.0f; dy = yfwd ? 2.0f : -2.0f;
                n_x = cut_range(width, xb, xe);
                n_y = cut_range(height, yb, ye);
                n_loop = n_x + n_y + 1;
                x = (float)(dx < 0 ? xe : xb);  // MI1:301-302
                y = (float)(dy < 0 ? ye : yb);
                tx = t0_ms; ty = t0_ms;  // point_coords slots 8, 9
                i_inte_last = (int)f2u(-1.0f);
                it = -1;
                state = WALK;
            }
            // ---- first half of one walk step (MI1:310-375): pick the side, emit the record
            if (state == WALK && !pending) {
                const bool first = it == -1;
                // MI1:321-359 written without branches: side, whether that side is exhausted, t_min
                side = first ? ((n_x != 0 && n_y == 0) ? 1 : 0) : ((tx <= ty) ? 0 : 1);
                const bool park = first ? (n_x == 0 || n_y == 0) : ((side ? n_y : n_x) <= 0);
                t_min = first ? tx : (side ? ty : tx);
                if (it >= 0) {  // MI1:361-375
                    int i_out = (int)f2u(t_min);
                    if ((f2u(t_min) & 0xFFFFFFFCu) == ((uint32_t)i_inte_last & 0xFFFFFFFCu)) {
                        i_out |= i_inte_last;
                        inter[pcnt - 1] = make_int2((int)c, i_out);
                    }
                    inter[pcnt] = make_int2((int)c, i_out);
                    i_inte_last = i_out;
                    ++pcnt;
                }
                if (park) {
                    WALK_FINISH_STEP(2.0f);  // t_solve = 2: the side is exhausted
                } else {
                    if (side) { --n_y; cst = y; y = __fadd_rn(y, dy); }
                    else { --n_x; cst = x; x = __fadd_rn(x, dx); }
                    pending = true;
                }
            }
        }
        if (__all_sync(0xFFFFFFFFu, state == DONE)) break;

        // =========== solve the pending crossing, whole warp converged (MI1:377-437) ===========
        if (pending) {
            float t_solve = 0.0f;
            const float c0 = side ? cp.y[0] : cp.x[0], c1 = side ? cp.y[1] : cp.x[1];
            if (type == T_CUBIC) {  // MI1:392-436
                const float c2 = side ? cp.y[2] : cp.x[2], c3 = side ? cp.y[3] : cp.x[3];
                // LERP(a,b,t) = a + t*(b-a): the first-level differences do not depend on t
                const float d01 = __fsub_rn(c1, c0), d12 = __fsub_rn(c2, c1), d23 = __fsub_rn(c3, c2);
                float t0 = t_min, t1 = t1_ms;
                float vt0;
                {
                    const float a0 = __fadd_rn(c0, __fmul_rn(t0, d01)), a1 = __fadd_rn(c1, __fmul_rn(t0, d12)),
                                a2 = __fadd_rn(c2, __fmul_rn(t0, d23));
                    const float b0 = lerpf(a0, a1, t0), b1 = lerpf(a1, a2, t0);
                    vt0 = lerpf(b0, b1, t0);
                }
                t_solve = t0;
                if (vt0 != cst) {
                    const float raw_t0 = t0;
                    uint32_t s0 = f2u(__fsub_rn(vt0, cst)), s_last = 0;
#pragma unroll 4
                    for (int j = 0; j < CUBIC_ITERATION_NUMBER; ++j) {
                        const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
                        const float a0 = __fadd_rn(c0, __fmul_rn(tm, d01)), a1 = __fadd_rn(c1, __fmul_rn(tm, d12)),
                                    a2 = __fadd_rn(c2, __fmul_rn(tm, d23));
                        const float b0 = lerpf(a0, a1, tm), b1 = lerpf(a1, a2, tm);
                        const float vtm = lerpf(b0, b1, tm);
                        t_solve = tm;
                        s_last = f2u(__fsub_rn(vtm, cst));
                        if ((int)(s_last ^ s0) >= 0) { t0 = tm; s0 = s_last; }  // vt0 = vtm
                        else t1 = tm;
                    }
                    if (fabsf(u2f(s_last)) > 1.f) t_solve = raw_t0;  // MI1:430-433
                }
            } else if (type == T_LINE) {  // MI1:379-385
                float a = __fsub_rn(c1, c0);
                a = (a != 0.0f) ? __fdiv_rn(1.0f, a) : 0.0f;
                float v = __fmul_rn(__fsub_rn(cst, c0), a);
                v = (v < t_min) ? t_min : v;        // GLSL max(x,y) = x<y ? y : x
                t_solve = (t1_ms < v) ? t1_ms : v;  // GLSL min(x,y) = y<x ? y : x
            } else if (type == T_QUADRIC || type == T_ARC) {
                // TODO arms in the reference: t_solve stays 0
            } else {  // any other type value: interpolateGeneralCurve returns 0 (MI1:81-83,144)
                t_solve = t_min;
                if (0.0f != cst) {
                    float t0 = t_min;
                    for (int j = 0; j < CUBIC_ITERATION_NUMBER; ++j) {  // vtm - c == vt0 - c: t0 always moves
                        const float tm = __fmul_rn(__fadd_rn(t0, t1_ms), 0.5f);
                        t_solve = tm; t0 = tm;
                    }
                    if (fabsf(__fsub_rn(0.0f, cst)) > 1.f) t_solve = t_min;
                }
            }
            WALK_FINISH_STEP(t_solve);
            pending = false;
        }
    }
#undef WALK_FINISH_STEP
}

// ------------------------------------------------------------------------------------------------
// K5: gen_fragment.comp:90-246. One thread per intersection record i: the curve piece between
// record i and i+1 lies in one 2x2 cell; emit its cell key and winding delta. Output is the sort
// input: a compact, order-preserving 64-bit key (path | row rank | cell x) and a 32-bit value
// (fragment index | (winding delta + 1) << 30), so that the gather of shuffle_fragment.comp is not
// needed after the sort. With taps the reference planes 0, 2, 4 are written as well.
// ------------------------------------------------------------------------------------------------
struct FragTaps {
    int *key32;  // plane 0 (gen_fragment.comp:221), [nf] = -1 at index nf (GF:240)
    int *path;   // plane 2
    int *wind;   // plane 4
};

__device__ __forceinline__ uint64_t pack_key(const KeyLayout &L, uint32_t path, bool valid, int pos_x, int pos_y) {
    // Signed int32 order of the reference key (SORT:69): rows y = 2,4,.. ascending, then
    // the invalid key 0xFFFEFFFE, then row y = 0; x ascending inside a row.
    uint32_t yk, xk;
    if (!valid) { yk = (uint32_t)(L.ny - 1); xk = 0; }
    else {
        yk = (pos_y == 0) ? (uint32_t)L.ny : (uint32_t)(pos_y / 2 - 1);
        xk = (uint32_t)((pos_x + FRAG_SIZE) / 2);
    }
    return ((uint64_t)path << (L.bits_x + L.bits_y)) | ((uint64_t)yk << L.bits_x) | (uint64_t)xk;
}

// Inverse of pack_key: the reference's 32-bit yx key and the path id.
__device__ __forceinline__ int unpack_key32(const KeyLayout &L, uint64_t k, uint32_t &path) {
    const uint32_t xk = (uint32_t)(k & ((1ull << L.bits_x) - 1));
    const uint32_t yk = (uint32_t)((k >> L.bits_x) & ((1ull << L.bits_y) - 1));
    path = (uint32_t)(k >> (L.bits_x + L.bits_y));
    if (yk == (uint32_t)(L.ny - 1)) return (int)0xFFFEFFFEu;
    const int pos_y = (yk == (uint32_t)L.ny) ? 0 : (int)(yk + 1) * 2;
    const int pos_x = (int)xk * 2 - FRAG_SIZE;
    return (int)(((uint32_t)(pos_y + 0x7FFF) << 16) | ((uint32_t)(pos_x + 0x7FFF) & 0xFFFFu));
}

__global__ void __launch_bounds__(256) k_gen_fragment(const FrameParams *__restrict__ P,
                                                      const FrameCounters *__restrict__ ctr, int capacity,
                                                      KeyLayout L, const int2 *__restrict__ inter,
                                                      const uint32_t *__restrict__ curve_path,
                                                      const uint32_t *__restrict__ curve_pos_map,
                                                      const uint32_t *__restrict__ curve_type,
                                                      const float2 *__restrict__ tpos,
                                                      uint64_t *__restrict__ key64, uint32_t *__restrict__ val,
                                                      FragTaps taps) {
    const int nf = ctr->n_fragments;
    if (nf > capacity) return;
    const int width = P->width, height = P->height;
    const bool cull = P->cull != 0;
    const int by0 = P->band_y0, by1 = P->band_y1;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
        const int2 r0 = inter[f];
        const int2 r1 = (f + 1 != nf) ? inter[f + 1] : make_int2(-1, 0x3f800000);  // GF:99
        float t0 = u2f((uint32_t)r0.y & 0xFFFFFFFCu);
        float t1 = u2f((uint32_t)r1.y & 0xFFFFFFFCu);
        t0 = (t0 < 0.0f) ? 0.0f : t0;  // GF:103-104
        t1 = (t1 < 0.0f) ? 0.0f : t1;
        const int cidx = r0.x;
        const uint32_t pidx = curve_path[cidx];
        if (r0.x != r1.x) t1 = 1.0f;  // GF:113-115
        bool valid = false;
        int pos_x = 0, pos_y = 0, wn = 0;
        if (t0 < t1) {
            const uint32_t type = curve_type[cidx];
            const uint32_t po = curve_pos_map[cidx];
            float2 cv0 = make_float2(0.f, 0.f), cv1 = cv0, cv2 = cv0, cv3 = cv0;
            if (type == T_LINE || type == T_QUADRIC || type == T_CUBIC) { cv0 = tpos[po]; cv1 = tpos[po + 1]; }  // GF:132-156
            if (type == T_QUADRIC || type == T_CUBIC) cv2 = tpos[po + 2];
            if (type == T_CUBIC) cv3 = tpos[po + 3];
            float pfx = cv0.x, pfy = cv0.y, plx = cv0.x, ply = cv0.y;  // GF:60: default result is cv0
            if (type == T_LINE) {
                pfx = lerpf(cv0.x, cv1.x, t0); pfy = lerpf(cv0.y, cv1.y, t0);
                plx = lerpf(cv0.x, cv1.x, t1); ply = lerpf(cv0.y, cv1.y, t1);
            } else if (type == T_CUBIC) {
                pfx = cubic_eval(cv0.x, cv1.x, cv2.x, cv3.x, t0); pfy = cubic_eval(cv0.y, cv1.y, cv2.y, cv3.y, t0);
                plx = cubic_eval(cv0.x, cv1.x, cv2.x, cv3.x, t1); ply = cubic_eval(cv0.y, cv1.y, cv2.y, cv3.y, t1);
            }
            const int raw_x = float2int_rd(__fdiv_rn(__fmul_rn(__fadd_rn(pfx, plx), 0.5f), 2.0f)) * FRAG_SIZE;  // GF:169-170
            const int raw_y = float2int_rd(__fdiv_rn(__fmul_rn(__fadd_rn(pfy, ply), 0.5f), 2.0f)) * FRAG_SIZE;
            pos_x = min(max(raw_x, -FRAG_SIZE), (int)(((uint32_t)width & 0xFFFFFFFEu) + FRAG_SIZE));  // GF:177-178
            pos_y = min(max(raw_y, -FRAG_SIZE), (int)(((uint32_t)height & 0xFFFFFFFEu) + FRAG_SIZE));
            valid = (uint32_t)raw_y < (uint32_t)height;  // GF:185-187
            const float wn_y = (float)(pos_y + 1);
            if (pfy == ply) wn = 0;  // GF:190-199
            else if (pfy < wn_y && wn_y <= ply) wn = -1;
            else if (ply < wn_y && wn_y <= pfy) wn = 1;
            if (cull && valid && (pos_y < by0 || pos_y >= by1)) { valid = false; wn = 0; }  // band mode (new)
        }
        key64[f] = pack_key(L, pidx, valid, pos_x, pos_y);
        val[f] = (uint32_t)f | ((uint32_t)(wn + 1) << 30);
        if (taps.key32) {
            taps.key32[f] = valid ? (int)(((uint32_t)(pos_y + 0x7FFF) << 16) | ((uint32_t)(pos_x + 0x7FFF) & 0xFFFFu))
                                  : (int)0xFFFEFFFEu;
            taps.path[f] = (int)pidx;
            taps.wind[f] = wn;
            if (f == 0) taps.key32[nf] = -1;  // GF:240
        }
    }
}

// Sort segment table (gen_fragment.comp:226-244) for the tap: seg[j] = first record of the first
// curve whose path is >= j; seg[n_paths] = nf. Derived from the scanned curve offsets.
__global__ void k_segments_tap(uint32_t n_curves, uint32_t n_paths, const uint32_t *__restrict__ curve_path,
                               const int *__restrict__ offsets, int *__restrict__ seg) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c <= n_curves; c += gridDim.x * blockDim.x) {
        const uint32_t lo = (c == 0) ? 0u : curve_path[c - 1] + 1u;
        const uint32_t hi = (c == n_curves) ? n_paths : curve_path[c];
        const int v = offsets[c];
        for (uint32_t j = lo; j <= hi && j <= n_paths; ++j) seg[j] = v;
    }
}

}  // namespace slpr
