// walk.cuh — make_intersection_1.comp:217-447 + gen_fragment.comp:90-246, re-formulated per
// monotone PIECE instead of per curve.
//
// The reference walks one curve per thread: per monotone piece it merges the x- and y-grid
// crossings in parameter order; lines in closed form, cubics by a 24-step bisection whose bracket
// starts at the previously emitted crossing on the same axis, so the walk along a piece is
// inherently sequential if the emitted t is to be bit-identical. Pieces, however, are independent:
// a piece starts from (t0, p0) = the previous piece's tagged end parameter and end point, both
// functions of the cut parameters alone. So:
//   k_piece_emit   one thread per curve: set up every piece (MI1:266-308) and store it as a 64-byte
//                  record at its slot in LENGTH-SORTED order (bucket, rank) from k_monotonize_count;
//   k_walk         one lane per piece, 32 consecutive records per warp = 32 pieces of (nearly) equal
//                  length, so the whole warp runs the same number of merge steps and the 24-step
//                  bisection — 70 % of all instructions — executes converged; the next group's
//                  ticket and its 2 KB of records are fetched (cp.async into shared memory) while
//                  the current group is walked; every emitted record
//                  also closes the fragment that began at the previous record (gen_fragment fused,
//                  the (curve, tbits) intersection records are only a debug tap now);
//   k_piece_fix    re-emits the boundary fragment of the rare pieces whose successor did not start exactly
//                  at its t0 (only possible after the MI0:340 cut-order slip); k_walk lists them.
// The arithmetic of every step is the reference's, operation for operation.
#pragma once
#include "geom.cuh"

namespace slpr {

struct __align__(16) PieceRec {
    float4 px, py;  // control points of the curve (x[0..3], y[0..3])
    float4 tt;      // t0_ms (tagged start), t1_ms (tagged end), bits: first x grid line | first y grid line << 16,
                    // bits: path | even-odd rule << 31
    uint4 m;        // n_x | n_y << 15 | dx<0 << 30 | dy<0 << 31;  curve;  first record index;  type | piece << 8
};

struct WalkTemp {
    int *group_counter;           // zeroed per frame
};

// Where the pieces were ranked: k_monotonize_count's launch shape (PieceLayout) and the exclusive per-(bucket,
// block) prefix inside the block's window from k_bucket_scan.
struct PieceRanks {
    const uint32_t *block_base;  // [WALK_BUCKETS][n_blocks]
    const uint32_t *vhist;       // [(n_windows + 1) * WALK_BUCKETS] pieces per virtual bucket (PieceLayout)
    PieceLayout lay;
};

// descending exclusive scan of the (at most WALK_VBUCKETS_MAX) virtual bucket counts by one block of 256 threads
__device__ __forceinline__ void vbucket_bases(const PieceRanks &pr, uint32_t *s_vbase, uint32_t *s_total) {
    __shared__ uint32_t s_part[8];
    const uint32_t nv = pr.lay.n_vbuckets();                   // a multiple of 64
    const uint32_t per = (nv + 255u) / 256u;                   // consecutive entries (in descending order) per thread
    const uint32_t first = threadIdx.x * per;                  // rank of this thread's first entry, 0 = highest index
    uint32_t sum = 0;
    for (uint32_t k = 0; k < per; ++k)
        if (first + k < nv) sum += pr.vhist[nv - 1u - (first + k)];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    uint32_t run = incl - sum;
    for (uint32_t k = 0; k < warp; ++k) run += s_part[k];
    for (uint32_t k = 0; k < per; ++k) {
        if (first + k < nv) {
            const uint32_t idx = nv - 1u - (first + k);
            s_vbase[idx] = run;
            run += pr.vhist[idx];
        }
    }
    if (threadIdx.x == 255) *s_total = run;
    __syncthreads();
}

#ifndef SLPR_LONG_MIN_RECORDS
#define SLPR_LONG_MIN_RECORDS 20
#endif
constexpr uint32_t LONG_MIN_RECORDS = SLPR_LONG_MIN_RECORDS;  // (> WALK_LONG, so that these buckets are never windowed)
constexpr uint32_t LONG_PIECES_MAX = 4096;

// ------------------------------------------------------------------------------------------------
#ifndef SLPR_PE_MIN_BLOCKS
#define SLPR_PE_MIN_BLOCKS 1
#endif
template <bool FMA>
__global__ void __launch_bounds__(256, SLPR_PE_MIN_BLOCKS) k_piece_emit(const FrameParams *__restrict__ P, uint32_t n_curves,
                                                    const uint32_t *__restrict__ curve_type,
                                                    const uint32_t *__restrict__ curve_pos_map,
                                                    const uint32_t *__restrict__ curve_path,
                                                    const uint32_t *__restrict__ fill_rule,
                                                    const float2 *__restrict__ tpos, const float *__restrict__ cut_cache,
                                                    const int *__restrict__ offsets, const uint32_t *__restrict__ slots,
                                                    FrameCounters *__restrict__ ctr, int capacity, PieceRanks ranks,
                                                    LiveCurves live, PieceRec *__restrict__ pieces, FullRvg full) {
    __shared__ uint32_t s_vbase[WALK_VBUCKETS_MAX];
    __shared__ uint32_t s_pos[WALK_BUCKETS];
    __shared__ uint32_t s_total;
    if (ctr->n_fragments > capacity) return;
    vbucket_bases(ranks, s_vbase, &s_total);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctr->n_pieces = (int)s_total;
        // The longest pieces are laid out first: [0, n_long) of the record array are the ones the long-piece kernels
        // walk when long_mode is on — the top length bucket (62 records or more) and as many of the next buckets,
        // down to LONG_MIN_RECORDS, as keep the count within LONG_PIECES_MAX (few pieces: a lane of k_walk then
        // never walks more than a couple of dozen crossings on its own; many pieces: nothing changes).
        const uint32_t *top = ranks.vhist + ranks.lay.n_windows * WALK_BUCKETS;
        uint32_t n_long = top[WALK_BUCKETS - 1];
        for (uint32_t b = WALK_BUCKETS - 2; b >= LONG_MIN_RECORDS && b > ranks.lay.long_min; --b) {
            if (n_long + top[b] > LONG_PIECES_MAX) break;
            n_long += top[b];
        }
        ctr->n_long = (int)n_long;
        ctr->n_top = (int)top[WALK_BUCKETS - 1];
    }
    // This block walks the work items the block of the same index ranked in k_monotonize_count (same launch
    // shape), so one 64-entry table — where this block's pieces of every length start — places all its pieces
    // without a global look-up per piece.
    if (threadIdx.x < WALK_BUCKETS)
        s_pos[threadIdx.x] = s_vbase[ranks.lay.vbucket(blockIdx.x, threadIdx.x)] + ranks.block_base[threadIdx.x * ranks.lay.n_blocks + blockIdx.x];
    __syncthreads();
    const int width = P->width, height = P->height;
    const uint32_t n_work = live.count(n_curves);
    const uint32_t ipb = ranks.lay.items_per_block(n_work);
    const uint32_t w_end = min(n_work, (blockIdx.x + 1u) * ipb);
    for (uint32_t w = blockIdx.x * ipb + threadIdx.x; w < w_end; w += blockDim.x) {
        const uint32_t c = live.curve(w);
        int pcnt = offsets[c];
        if (offsets[c + 1] == pcnt) continue;  // invisible or band-culled
        uint32_t slot[5];  // all five up front (those past the curve's last piece are never used)
#pragma unroll
        for (int k = 0; k < 5; ++k) slot[k] = slots[5 * c + k];
        const uint32_t type = curve_type[c];
        const uint32_t pidx = curve_path[c];
        // MARK:82 only distinguishes rule 1 (even-odd) from the rest
        const uint32_t path_rule = pidx | (fill_rule[pidx] == 1u ? 0x80000000u : 0u);
        CurvePts cp;
        load_points(type, curve_pos_map[c], tpos, cp);
        full.stash_weight(type, c, cp);  // (an ARC's weight then travels in the piece record's fourth x slot)
        const float q0 = cut_cache[5 * c + 0], q1 = cut_cache[5 * c + 1], q2 = cut_cache[5 * c + 2], q3 = cut_cache[5 * c + 3];
        // count > 0 implies the path is visible (MI0:374-377), so MI1:257-260 appends t = 1
        const uint32_t n_cuts = f2u(cut_cache[5 * c + 4]) + 1u;  // MI1:255
        float t0_ms = 0.f, p0x = cp.x[0], p0y = cp.y[0];
        for (uint32_t piece = 0; piece < n_cuts; ++piece) {  // MI1:266-308
            float t1_ms = (piece + 1 == n_cuts) ? 1.f : (piece == 0) ? q0 : (piece == 1) ? q1 : (piece == 2) ? q2 : q3;
            const float p1x = interp_full<FMA>(type, t1_ms, cp.x[0], cp.x[1], cp.x[2], cp.x[3], 0.0f, full.on());
            const float p1y = interp_full<FMA>(type, t1_ms, cp.y[0], cp.y[1], cp.y[2], cp.y[3], 0.0f, full.on());
            // MI1:271-276: tag t1 in its two mantissa LSBs
            if (floorf(p1x) == p1x) t1_ms = u2f((f2u(t1_ms) & 0xFFFFFFFCu) | 2u);
            else t1_ms = u2f(f2u(t1_ms) | 3u);
            // get_xy_begin_end_delta, MI1:150-170 (float2int_rd)
            const bool xfwd = p0x <= p1x, yfwd = p0y <= p1y;
            int xb = float2int_rd(__fmul_rn(xfwd ? p0x : p1x, 0.5f)) * FRAG_SIZE + FRAG_SIZE;
            int xe = float2int_rd(__fmul_rn(xfwd ? p1x : p0x, 0.5f)) * FRAG_SIZE;
            int yb = float2int_rd(__fmul_rn(yfwd ? p0y : p1y, 0.5f)) * FRAG_SIZE + FRAG_SIZE;
            int ye = float2int_rd(__fmul_rn(yfwd ? p1y : p0y, 0.5f)) * FRAG_SIZE;
            const int n_x = cut_range(width, xb, xe);
            const int n_y = cut_range(height, yb, ye);
            PieceRec r;
            r.px = make_float4(cp.x[0], cp.x[1], cp.x[2], cp.x[3]);
            r.py = make_float4(cp.y[0], cp.y[1], cp.y[2], cp.y[3]);
            // MI1:301-302; the grid lines are clamped to [0, dim + 2] whenever they are used (count > 0)
            r.tt = make_float4(t0_ms, t1_ms, u2f(((uint32_t)(xfwd ? xb : xe) & 0xFFFFu) | ((uint32_t)(yfwd ? yb : ye) << 16)),
                               u2f(path_rule));
            // Does the NEXT piece of the curve end below its start (cuts left out of order by the MI0:340 slip)? Only
            // then can its first record differ from its start parameter, and only then does k_piece_fix need this
            // piece's last parameter: the piece is told to leave it (bit 16). Conservative in the two tag bits.
            uint32_t next_unordered = 0;
            if (piece + 1 < n_cuts) {
                const float next_raw = (piece + 2 == n_cuts) ? 1.f : (piece == 0) ? q1 : (piece == 1) ? q2 : q3;
                next_unordered = ((f2u(next_raw) & 0xFFFFFFFCu) <= (f2u(t1_ms) | 3u)) ? 1u : 0u;
            }
            r.m = make_uint4((uint32_t)n_x | ((uint32_t)n_y << 15) | (xfwd ? 0u : 1u << 30) | (yfwd ? 0u : 1u << 31), c,
                             (uint32_t)pcnt, (type & 0xFFu) | (piece << 8) | ((type > 0xFFu) ? 0x80u : 0u) | (next_unordered << 16));
            const uint32_t sl = (piece == 0) ? slot[0] : (piece == 1) ? slot[1] : (piece == 2) ? slot[2] : (piece == 3) ? slot[3] : slot[4];
            pieces[s_pos[sl >> 26] + (sl & 0x03FFFFFFu)] = r;
            pcnt += n_x + n_y + 1;
            t0_ms = t1_ms; p0x = p1x; p0y = p1y;  // MI1:442-443
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Plain PTX atomic: atomicAdd() by one lane is compiled into the warp-aggregated form whose result
// is shuffled — and therefore waited for — right away; here the ticket is only read a group later.
__device__ __forceinline__ uint32_t take_ticket(int *counter) {
    uint32_t t;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(counter) : "memory");
    return t;
}

// A piece's fragments have consecutive indices, so a lane holds keys back until it has an aligned
// pair (one 16-byte store instead of two 8-byte ones) and values until it has an aligned quad: the
// scattered stores of the walk are its second cost after the bisection.
#ifndef SLPR_WALK_PAIR
#define SLPR_WALK_PAIR 2 /* 1: held-back pairs / quads with one arm per f mod 4 (divergent); 2: sliding window, predicated stores: walk 0.4604 -> 0.4510 ms */
#endif
struct FragStore {
    uint64_t pk;
    uint32_t pv0, pv1, pv2;
    int f_first;
    bool direct;  // (kernel-uniform) few groups per warp: the frame is bound by its longest pieces, not by stores —
                  // plain stores keep the per-crossing instruction count down (line-heavy scenes: -20 %)
    __device__ __forceinline__ void put(int f, uint64_t k, uint32_t v, uint64_t *__restrict__ key64, uint32_t *__restrict__ val) {
        if (!SLPR_WALK_PAIR || direct) {
            key64[f] = k; val[f] = v;
            return;
        }
#if SLPR_WALK_PAIR == 2
        // Without divergent arms (the lanes of a warp are at different f mod 4): the held-back key is simply the last
        // one, the held-back values a sliding window (pv0, pv1, pv2 = values of f-3, f-2, f-1); the stores are predicated.
        if (f & 1) {
            if (f > f_first) *reinterpret_cast<ulonglong2 *>(key64 + f - 1) = make_ulonglong2(pk, k);
            else key64[f] = k;
        }
        pk = k;
        if ((f & 3) == 3) {
            if (f - 3 >= f_first) *reinterpret_cast<uint4 *>(val + f - 3) = make_uint4(pv0, pv1, pv2, v);
            else {
                if (f - 2 >= f_first) val[f - 2] = pv1;
                if (f - 1 >= f_first) val[f - 1] = pv2;
                val[f] = v;
            }
        }
        pv0 = pv1; pv1 = pv2; pv2 = v;
#else
        if (f & 1) {
            if (f > f_first) *reinterpret_cast<ulonglong2 *>(key64 + f - 1) = make_ulonglong2(pk, k);
            else key64[f] = k;
        } else
            pk = k;
        const int q = f & 3;
        if (q == 3) {
            if (f - 3 >= f_first) *reinterpret_cast<uint4 *>(val + f - 3) = make_uint4(pv0, pv1, pv2, v);
            else {
                if (f - 2 >= f_first) val[f - 2] = pv1;
                if (f - 1 >= f_first) val[f - 1] = pv2;
                val[f] = v;
            }
        } else if (q == 0) pv0 = v;
        else if (q == 1) pv1 = v;
        else pv2 = v;
#endif
    }
    // after the piece's last fragment
    __device__ __forceinline__ void flush(int f_last, uint64_t *__restrict__ key64, uint32_t *__restrict__ val) {
        if (!SLPR_WALK_PAIR || direct) return;
        if (!(f_last & 1)) key64[f_last] = pk;
        const int q = f_last & 3, base = f_last - q;
#if SLPR_WALK_PAIR == 2
        if (q != 3) {  // window: pv2 = value of f_last, pv1 of f_last - 1, pv0 of f_last - 2
            val[f_last] = pv2;
            if (q >= 1 && f_last - 1 >= f_first) val[f_last - 1] = pv1;
            if (q >= 2 && base >= f_first) val[base] = pv0;
        }
#else
        if (q != 3) {
            if (base >= f_first) val[base] = pv0;
            if (q >= 1 && base + 1 >= f_first) val[base + 1] = pv1;
            if (q >= 2) val[base + 2] = pv2;
        }
#endif
    }
};

constexpr int WALK_THREADS = 128;
#ifndef SLPR_WALK_UNROLL
#define SLPR_WALK_UNROLL 4
#endif
constexpr int WALK_UNROLL = SLPR_WALK_UNROLL;  // bisection steps per loop trip
#ifndef SLPR_WALK_MIN_BLOCKS
#define SLPR_WALK_MIN_BLOCKS 1
#endif

// One crossing of the piece with the grid line `cst` on axis `side` (0: x, 1: y), searched from t_min (the previous
// crossing on this axis) up to the piece's end t1_ms: make_intersection_1.comp:377-437, shared by k_walk and the
// long-piece chains (k_long_chains) so that both produce the same bits.
template <bool FULL, bool FMA = false>
__device__ __forceinline__ float solve_crossing(uint32_t type, const CurvePts &cp, int side, float t_min, float t1_ms, float cst) {
    float t_solve = 0.0f;
    const float c0 = side ? cp.y[0] : cp.x[0], c1 = side ? cp.y[1] : cp.x[1];
    if (type == T_CUBIC) {  // MI1:392-436
        const float c2 = side ? cp.y[2] : cp.x[2], c3 = side ? cp.y[3] : cp.x[3];
        // LERP(a,b,t) = a + t*(b-a): the first-level differences do not depend on t
        const float d01 = __fsub_rn(c1, c0), d12 = __fsub_rn(c2, c1), d23 = __fsub_rn(c3, c2);
        float t0 = t_min, t1 = t1_ms;
        float vt0;
        {
            const float a0 = madd_t<FMA>(t0, d01, c0), a1 = madd_t<FMA>(t0, d12, c1), a2 = madd_t<FMA>(t0, d23, c2);
            const float b0 = lerp_t<FMA>(a0, a1, t0), b1 = lerp_t<FMA>(a1, a2, t0);
            vt0 = lerp_t<FMA>(b0, b1, t0);
        }
        t_solve = t0;
        if (vt0 != cst) {
            const float raw_t0 = t0;
            // the sign of (vt0 - c) never changes: t0 only moves to points of the same sign
            const bool neg0 = (int)f2u(__fsub_rn(vt0, cst)) < 0;
            uint32_t s_last = 0;
#pragma unroll WALK_UNROLL
            for (int j = 0; j < CUBIC_ITERATION_NUMBER; ++j) {
                const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
                const float a0 = madd_t<FMA>(tm, d01, c0), a1 = madd_t<FMA>(tm, d12, c1), a2 = madd_t<FMA>(tm, d23, c2);
                const float b0 = lerp_t<FMA>(a0, a1, tm), b1 = lerp_t<FMA>(a1, a2, tm);
                const float vtm = lerp_t<FMA>(b0, b1, tm);
                t_solve = tm;
                s_last = f2u(__fsub_rn(vtm, cst));
                // same sign as at t0: t0 = tm (vt0 = vtm, MI1:421-424), else t1 = tm. One
                // predicate instruction (sign test XOR the loop-invariant sign) + two selects.
                asm("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %3, 0;\n\tsetp.lt.xor.s32 p, %2, 0, q;\n\t"
                    "selp.f32 %0, %0, %4, p;\n\tselp.f32 %1, %4, %1, p;\n\t}"
                    : "+f"(t0), "+f"(t1)
                    : "r"(s_last), "r"((int)neg0), "f"(tm));
            }
            if (fabsf(u2f(s_last)) > 1.f) t_solve = raw_t0;  // MI1:430-433
        }
    } else if (type == T_LINE) {  // MI1:379-385
        float a = __fsub_rn(c1, c0);
        a = (a != 0.0f) ? __fdiv_rn(1.0f, a) : 0.0f;
        float v = __fmul_rn(__fsub_rn(cst, c0), a);
        v = (v < t_min) ? t_min : v;        // GLSL max(x,y) = x<y ? y : x
        t_solve = (t1_ms < v) ? t1_ms : v;  // GLSL min(x,y) = y<x ? y : x
    } else if (FULL && (type == T_QUADRIC || type == T_ARC)) {  // f-1: the bisection of MI1:392-436 on this curve's evaluator
        const float c2 = side ? cp.y[2] : cp.x[2], w = cp.x[3];
        float t0 = t_min, t1 = t1_ms;
        float vt0 = (type == T_ARC) ? eval_arc<FMA>(c0, c1, c2, w, t0) : eval_quadric<FMA>(c0, c1, c2, t0);
        t_solve = t0;
        if (vt0 != cst) {
            const float raw_t0 = t0;
            float last_vtm = 0.f;
            for (int j = 0; j < CUBIC_ITERATION_NUMBER; ++j) {
                const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
                const float vtm = (type == T_ARC) ? eval_arc<FMA>(c0, c1, c2, w, tm) : eval_quadric<FMA>(c0, c1, c2, tm);
                t_solve = tm; last_vtm = vtm;
                if ((int)(f2u(__fsub_rn(vtm, cst)) ^ f2u(__fsub_rn(vt0, cst))) >= 0) { t0 = tm; vt0 = vtm; }
                else t1 = tm;
            }
            if (fabsf(__fsub_rn(last_vtm, cst)) > 1.f) t_solve = raw_t0;
        }
    } else if (type == T_QUADRIC || type == T_ARC) {
        t_solve = 0.0f;  // TODO arms in the reference: t_solve stays 0
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
    return t_solve;
}

// FULL (SLPR_FLAG_FULL_RVG, f-1): QUADRIC / ARC pieces are walked with the reference's bisection on their own evaluators
// (common.cuh) instead of the reference's TODO arms; a separate instantiation, so the default kernel is untouched.
template <bool FULL, bool FMA>
__global__ void __launch_bounds__(WALK_THREADS, SLPR_WALK_MIN_BLOCKS) k_walk(const FrameParams *__restrict__ P,
                                                       const PieceRec *__restrict__ pieces,
                                                       FrameCounters *__restrict__ ctr, int capacity, WalkTemp tmp,
                                                       KeyLayout L, uint64_t *__restrict__ key64,
                                                       uint32_t *__restrict__ val, FragTaps taps,
                                                       int2 *__restrict__ inter, float2 *__restrict__ boundary,
                                                       uint4 *__restrict__ fixlist, int skip_long) {
    const int nf_total = ctr->n_fragments;
    if (nf_total > capacity) return;
    if (taps.key32 && blockIdx.x == 0 && threadIdx.x == 0 && nf_total > 0) taps.key32[nf_total] = -1;  // GF:240
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    __shared__ uint4 s_stage[WALK_THREADS / 32][2][128];  // per warp: two groups of 32 records (2 KB each)
    const uint32_t n_pieces = (uint32_t)ctr->n_pieces;  // k_piece_emit
    const uint32_t n_long = skip_long ? (uint32_t)ctr->n_long : 0u;  // long pieces are walked by k_long_chains / k_long_emit
    const FragEnv env = load_frag_env(P);
    const uint32_t n_chunks = n_pieces * 4u;  // 16-byte chunks in the record array
    // group g's records -> stage st, one commit group per call (empty past the end)
    auto stage_group = [&](uint32_t g, int st) {
        const unsigned long long first = (unsigned long long)g * 128ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t q = j * 32 + lane;
            if (first + q < n_chunks) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_stage[warp][st][q]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst),
                             "l"(reinterpret_cast<const uint4 *>(pieces) + first + q));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Groups are taken two ahead so that neither the ticket atomic nor the record fetch is waited for.
    // The first two rounds are dealt statically (warp w: groups w and w + #warps): groups are sorted
    // longest first, and a warp holding two CONSECUTIVE tickets would walk the two longest groups one
    // after the other — on a scene bound by its longest pieces that doubles the frame time.
    const uint32_t n_warps = gridDim.x * (WALK_THREADS / 32);
    uint32_t g_cur = blockIdx.x * (WALK_THREADS / 32) + warp, g_next = g_cur + n_warps;
    stage_group(g_cur, 0);
    for (int st = 0;; st ^= 1) {
        if ((unsigned long long)g_cur * 32ull >= n_pieces) break;  // tickets only grow: nothing left for this warp
        uint32_t g_after = 0;
        if (lane == 0) g_after = take_ticket(tmp.group_counter) + 2u * n_warps;
        stage_group(g_next, st ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const uint32_t idx = g_cur * 32u + lane;
        const bool active = idx < n_pieces && idx >= n_long;

        // ---- the piece, from the staged copy
        CurvePts cp;
        float t0_ms = 0.f, t1_ms = 0.f, x = 0.f, y = 0.f, dx = 2.f, dy = 2.f;
        int n_x = 0, n_y = 0, n_loop = -2, pcnt = 0;
        uint32_t c = 0, type = T_LINE, piece = 0, pidx = 0, rule_bit = 0, keep_last = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { cp.x[i] = 0.f; cp.y[i] = 0.f; }
        if (active) {
            const uint4 *r = &s_stage[warp][st][4 * lane];
            const uint4 a = r[0], b = r[1], t = r[2], m = r[3];
            cp.x[0] = u2f(a.x); cp.x[1] = u2f(a.y); cp.x[2] = u2f(a.z); cp.x[3] = u2f(a.w);
            cp.y[0] = u2f(b.x); cp.y[1] = u2f(b.y); cp.y[2] = u2f(b.z); cp.y[3] = u2f(b.w);
            t0_ms = u2f(t.x); t1_ms = u2f(t.y);
            x = (float)(t.z & 0xFFFFu); y = (float)(t.z >> 16);
            pidx = t.w & 0x7FFFFFFFu; rule_bit = t.w >> 31;
            n_x = (int)(m.x & 0x7FFFu); n_y = (int)((m.x >> 15) & 0x7FFFu);
            dx = (m.x & (1u << 30)) ? -2.f : 2.f; dy = (m.x & (1u << 31)) ? -2.f : 2.f;
            c = m.y; pcnt = (int)m.z;
            type = (m.w & 0x80u) ? 0xFFFFu : (m.w & 0x7Fu);  // types above 0xFF only need to be "other"
            piece = (m.w >> 8) & 0xFFu;
            keep_last = (m.w >> 16) & 1u;
            n_loop = n_x + n_y + 1;
        }
        __syncwarp();  // every lane has read its record: the stage may be refilled two groups from now
        g_cur = g_next;
        g_next = __shfl_sync(0xFFFFFFFFu, g_after, 0);
        FragStore fs;
        fs.pk = 0; fs.pv0 = fs.pv1 = fs.pv2 = 0; fs.f_first = pcnt;
        fs.direct = (unsigned long long)n_pieces <= 64ull * n_warps;  // at most two groups per warp
        float tx = t0_ms, ty = t0_ms;  // point_coords slots 8, 9 (MI1:304-305)
        int i_inte_last = (int)f2u(-1.0f);
        bool have_prev = false;
        float prev_t = 0.f, prev_x = 0.f, prev_y = 0.f;
        uint32_t first_bits = 0, last_bits = 0;

        for (int it = -1;; ++it) {  // MI1:310-441, the whole warp in lock step
            const bool live = active && it < n_loop;
            if (!__any_sync(0xFFFFFFFFu, live)) break;
            // MI1:321-359 without branches: side, whether that side is exhausted, t_min
            const bool first = it == -1;
            const int side = first ? ((n_x != 0 && n_y == 0) ? 1 : 0) : ((tx <= ty) ? 0 : 1);
            const bool park = first ? (n_x == 0 || n_y == 0) : ((side ? n_y : n_x) <= 0);
            const float t_min = first ? tx : (side ? ty : tx);
            float cst = 0.f;
            if (live && !park) {
                if (side) { --n_y; cst = y; y = __fadd_rn(y, dy); }
                else { --n_x; cst = x; x = __fadd_rn(x, dx); }
            }
            if (live && it >= 0) {  // MI1:361-375 + gen_fragment for the fragment that ends here
                if (inter) {
                    int i_out = (int)f2u(t_min);
                    if ((f2u(t_min) & 0xFFFFFFFCu) == ((uint32_t)i_inte_last & 0xFFFFFFFCu)) {
                        i_out |= i_inte_last;
                        inter[pcnt - 1] = make_int2((int)c, i_out);
                    }
                    inter[pcnt] = make_int2((int)c, i_out);
                    i_inte_last = i_out;
                }
                if (it == 0) first_bits = f2u(t_min);
                last_bits = f2u(t_min);
                // GF:101-104: the record's parameter without its tag bits, clamped at 0
                float tc = u2f(f2u(t_min) & 0xFFFFFFFCu);
                tc = (tc < 0.0f) ? 0.0f : tc;
                float ex, ey;
                eval_point<FULL, FMA>(type, cp, tc, ex, ey);
                if (have_prev) {
                    uint64_t k; uint32_t v;
                    make_fragment(env, L, pcnt - 1, pidx, rule_bit, prev_t, tc, prev_x, prev_y, ex, ey, k, v, taps);
                    fs.put(pcnt - 1, k, v, key64, val);
                }
                prev_t = tc; prev_x = ex; prev_y = ey; have_prev = true;
                ++pcnt;
            }
            // ---- solve the next crossing on `side` (MI1:377-437); converged: every lane is at the same step
            float t_solve = park ? 2.0f : 0.0f;
            const bool solve = live && !park;
            if (__any_sync(0xFFFFFFFFu, solve)) {
                if (solve) t_solve = solve_crossing<FULL, FMA>(type, cp, side, t_min, t1_ms, cst);
            }
            if (live) {  // MI1:440
                const float tagged = u2f((f2u(t_solve) & 0xFFFFFFFCu) | (uint32_t)side);
                if (side) ty = tagged; else tx = tagged;
            }
        }
        // ---- the fragment across the piece boundary: last record of this piece -> first record of the
        //      next piece of the curve, or t = 1 at the curve's end (GF:99,113-115). The next piece's first
        //      record is its start parameter t0 = this piece's tagged t1 (MI1:304-305,442) in every case
        //      but one: when the MI0:340 slip leaves the cuts out of order, a bisection can land below t0.
        //      So the boundary fragment is emitted here with t1 = (t1_ms without tag bits), and a piece
        //      whose first record does not match its t0 is listed for k_piece_fix, which redoes its
        //      predecessor's boundary fragment from the recorded parameters.
        if (active) {
            float tcl = u2f(f2u(t1_ms) & 0xFFFFFFFCu);
            tcl = (tcl < 0.0f) ? 0.0f : tcl;
            float ex, ey;
            eval_point<FULL, FMA>(type, cp, tcl, ex, ey);
            uint64_t k; uint32_t v;
            make_fragment(env, L, pcnt - 1, pidx, rule_bit, prev_t, tcl, prev_x, prev_y, ex, ey, k, v, taps);
            fs.put(pcnt - 1, k, v, key64, val);
            fs.flush(pcnt - 1, key64, val);
            // The first / last parameters are only left behind where k_piece_fix can need them: a piece that ends
            // below its start (every parameter it emits otherwise lies at or above its start: midpoints of a bracket
            // [t0, t1 >= t0], closed-form line crossings clamped into it) and the piece before such a piece.
            const bool unordered = piece > 0 && (f2u(t1_ms) & 0xFFFFFFFCu) <= (f2u(t0_ms) | 3u);
            if (unordered || keep_last) boundary[5 * c + piece] = make_float2(u2f(first_bits), u2f(last_bits));
            if (piece > 0 && (first_bits & 0xFFFFFFFCu) != (f2u(t0_ms) & 0xFFFFFFFCu)) {  // rare: a few dozen per million curves
                if (unordered) fixlist[atomicAdd(&ctr->n_fix, 1)] = make_uint4(c, piece, (uint32_t)fs.f_first, 0u);
                else ctr->fix_missed = 1;  // cannot happen (see above); the host refuses the frame if it does
            }
        }
    }
}

// One crossing, solved by a whole warp: the bisection of MI1:392-436 is a chain of 24 dependent evaluations (about 57
// cycles each on one lane — what a frame of few, long curves waits for). A warp runs it five steps at a time: lane L
// in 1..31 is node L of the binary tree of possible outcomes of the next five steps (depth d = floor(log2 L); the bits
// of L below the leading one are the decisions taken on the way down, 1 = "t0 moves"), derives that node's bracket with
// the same (t0 + t1) * 0.5 roundings the sequential loop would perform along that path, and evaluates the curve at
// its midpoint. One ballot collects the 31 sign tests; the leaf whose ancestors all decided the way its index says
// holds the bracket after the five steps and hands it to the warp. Every value is produced by the same operations, in
// the same order, as in solve_crossing — only the evaluations that the sequential loop would not have reached are
// thrown away. 24 steps = 4 rounds of five and one of four: ~1.7 x faster than one lane (profiles/README.md).
// ev(t) must be solve_crossing's evaluator for the curve type; returns the parameter for all lanes.
#ifndef SLPR_LONG_WARP
#define SLPR_LONG_WARP 1
#endif
// What a lane needs to know about its node, computed once per kernel; selections by lane-constant conditions are
// bitwise (one LOP3 on a mask register) so that no predicate has to be kept or recomputed inside the rounds.
__device__ __forceinline__ float bsel(uint32_t m, float a, float b) { return u2f((f2u(a) & m) | (f2u(b) & ~m)); }
struct TreeLane {
    uint32_t right[4];  // all ones: the path to this node goes right (t0 moves) at level k; 0: left, or below the node
    uint32_t at[5];     // all ones at this node's depth (lane 0 evaluates the root a second time, unused)
    uint32_t amask, aexp;  // the ballot bits of this node's ancestors, and what they must read for the bisection to come here
    bool leaf5, leaf4;     // last level of a five-step / four-step round
    __device__ __forceinline__ explicit TreeLane(uint32_t lane) {
        const int d = lane ? 31 - __clz((int)lane) : 0;
        amask = aexp = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool r = lane && k < d && ((lane >> (d - 1 - k)) & 1u);
            right[k] = r ? 0xFFFFFFFFu : 0u;
            if (lane && k < d) { amask |= 1u << (lane >> (d - k)); if (r) aexp |= 1u << (lane >> (d - k)); }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) at[k] = (k == d) ? 0xFFFFFFFFu : 0u;
        leaf5 = lane && d == 4; leaf4 = lane && d == 3;
    }
};

template <class Eval>
__device__ __forceinline__ float warp_bisect(Eval ev, float t_min, float t1_ms, float cst, const TreeLane &tl, uint32_t lane) {
    float t0 = t_min, t1 = t1_ms;
    uint32_t flip = 0;  // all ones when v(t0) - cst is negative: the ballots of the sign bits, flipped, read "same sign as at t0"
    float t_last = 0.f;
    uint32_t s_last = 0;
#pragma unroll
    for (int done = 0; done < CUBIC_ITERATION_NUMBER; done += 5) {
        const bool five = done + 5 <= CUBIC_ITERATION_NUMBER;  // else the last, four-step round
        // Four levels down this lane's path (lanes of shallower nodes just go on to the left; they keep what they
        // passed). A level's midpoint is (previous midpoint + the end that stayed) * 0.5 — the sum the sequential
        // loop forms, fp addition commutes — so the dependent chain is one add and one multiply per level.
        float lo = t0, hi = t1, tm = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
        float my_tm = tm, lo3 = lo, hi3 = hi;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float other = bsel(tl.right[k], hi, lo);
            lo = bsel(tl.right[k], tm, lo);
            hi = bsel(tl.right[k], hi, tm);
            tm = __fmul_rn(__fadd_rn(tm, other), 0.5f);
            my_tm = bsel(tl.at[k + 1], tm, my_tm);
            if (k == 2) { lo3 = lo; hi3 = hi; }
        }
        if (done == 0 && lane == 0) my_tm = t0;  // the first round's spare lane evaluates v(t0) (MI1:396-404) alongside
        const float v = ev(my_tm);
        const uint32_t sb = f2u(__fsub_rn(v, cst));
        uint32_t mask = __ballot_sync(0xFFFFFFFFu, (int)sb < 0);
        if (done == 0) {
            if (__any_sync(0xFFFFFFFFu, lane == 0 && v == cst)) return t0;  // the curve is on the grid line at t0
            flip = (mask & 1u) ? 0u : 0xFFFFFFFFu;  // the sign at t0 never changes: t0 only moves to points of the same sign
        }
        mask ^= flip;  // bit L: node L has the sign of t0, so t0 = tm there (MI1:421-424), else t1 = tm
        const bool same = ((mask >> lane) & 1u) != 0u;
        // exactly one leaf of the round's last level has all its ancestors deciding the way that leads to it: it holds
        // the bracket after the round and hands it to the warp
        const bool on = (five ? tl.leaf5 : tl.leaf4) && (mask & tl.amask) == tl.aexp;
        const float nlo = five ? lo : lo3, nhi = five ? hi : hi3;
        const int src = 31 - __clz((int)__ballot_sync(0xFFFFFFFFu, on));
        t0 = __shfl_sync(0xFFFFFFFFu, same ? my_tm : nlo, src);
        t1 = __shfl_sync(0xFFFFFFFFu, same ? nhi : my_tm, src);
        if (!five) {
            t_last = __shfl_sync(0xFFFFFFFFu, my_tm, src);
            s_last = __shfl_sync(0xFFFFFFFFu, sb, src);
        }
    }
    return (fabsf(u2f(s_last)) > 1.f) ? t_min : t_last;  // MI1:430-433
}

// ------------------------------------------------------------------------------------------------
// Long pieces (round 2). A piece's crossings with the x grid lines form one chain — every bisection bracket starts
// at the previous x crossing (MI1:392-436) — and its crossings with the y grid lines another; only the order in
// which the reference EMITS them couples the two (MI1:321-359: repeatedly the smaller of the two chain heads, x on
// ties). k_walk gives a piece to one lane, which is right for a million short pieces but leaves a frame of a few
// thousand curves waiting for the one lane that walks its longest piece: the shipped scenes at 4K spend 85-95 % of
// the frame in k_walk (tiger: 1.28 of 1.42 ms, one cubic of ~2000 crossings; test.rvg: a full-height line). With
// long_mode on (few long pieces; chosen by the host from the last frame's count), pieces of 62 or more crossings are
// instead walked in three steps, bit-identical to the sequential walk:
//   k_long_chains  one warp per (piece, axis): a LINE's crossings are closed-form and independent but for a clamp
//                  against the previous one — all lanes evaluate them at once and check that no clamp was active
//                  (else lane 0 redoes the chain in order); a curve's chain is walked crossing by crossing, each
//                  crossing by the whole warp (warp_bisect: five bisection steps per round);
//   k_long_emit    one block per piece: the merged emission order by rank (binary search; valid when both chains
//                  ascend, which is checked — else lane 0 merges head by head), then all lanes form the fragments
//                  between consecutive records exactly as k_walk does (make_fragment), plus the boundary fragment
//                  and the rare boundary-repair bookkeeping.
// Scratch: two 32-bit words per record, in the sort's output buffers (free until the sort).
// ------------------------------------------------------------------------------------------------
struct LongScratch {
    uint32_t *chain;   // [capacity] tagged crossing parameters: x chain at [first record, +n_x), y chain behind it
    uint32_t *merged;  // [capacity] the emitted record parameters in order
};

constexpr int LONG_EMIT_THREADS = 128;   // k_long_emit
constexpr int LONG_WARP_RECORDS = 256;   // pieces of up to this many records are emitted by one warp, longer ones by a block
constexpr int LONG_STAGE_WORDS = 5632;   // records of a piece whose chains and merged order it keeps in shared memory (2 x 22 KB)

struct LongPiece {
    CurvePts cp;
    float t0_ms, t1_ms, x0, y0, dx, dy;
    int n_x, n_y, pcnt;
    uint32_t c, type, piece, pidx, rule_bit, keep_last;
};
__device__ __forceinline__ LongPiece load_long_piece(const PieceRec *__restrict__ pieces, uint32_t idx) {
    const uint4 *r = reinterpret_cast<const uint4 *>(pieces + idx);
    const uint4 a = r[0], b = r[1], t = r[2], m = r[3];
    LongPiece p;
    p.cp.x[0] = u2f(a.x); p.cp.x[1] = u2f(a.y); p.cp.x[2] = u2f(a.z); p.cp.x[3] = u2f(a.w);
    p.cp.y[0] = u2f(b.x); p.cp.y[1] = u2f(b.y); p.cp.y[2] = u2f(b.z); p.cp.y[3] = u2f(b.w);
    p.t0_ms = u2f(t.x); p.t1_ms = u2f(t.y);
    p.x0 = (float)(t.z & 0xFFFFu); p.y0 = (float)(t.z >> 16);
    p.pidx = t.w & 0x7FFFFFFFu; p.rule_bit = t.w >> 31;
    p.n_x = (int)(m.x & 0x7FFFu); p.n_y = (int)((m.x >> 15) & 0x7FFFu);
    p.dx = (m.x & (1u << 30)) ? -2.f : 2.f; p.dy = (m.x & (1u << 31)) ? -2.f : 2.f;
    p.c = m.y; p.pcnt = (int)m.z;
    p.type = (m.w & 0x80u) ? 0xFFFFu : (m.w & 0x7Fu);
    p.piece = (m.w >> 8) & 0xFFu;
    p.keep_last = (m.w >> 16) & 1u;
    return p;
}

template <bool FULL, bool FMA>
__global__ void __launch_bounds__(128) k_long_chains(const PieceRec *__restrict__ pieces, const FrameCounters *__restrict__ ctr,
                                                     int capacity, LongScratch sc) {
    if (ctr->n_fragments > capacity) return;
    const uint32_t n_items = 2u * (uint32_t)ctr->n_long;
    const uint32_t lane = lane_id();
    const TreeLane tl(lane);
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += n_warps) {
        const int side = (int)(item & 1u);
        const LongPiece p = load_long_piece(pieces, item >> 1);
        const int n = side ? p.n_y : p.n_x;
        if (n == 0) continue;
        const float g0 = side ? p.y0 : p.x0, d = side ? p.dy : p.dx;
        uint32_t *out = sc.chain + p.pcnt + (side ? p.n_x : 0);
        bool done = false;
        if (p.type == T_LINE) {  // MI1:379-385, all crossings at once
            const float c0 = side ? p.cp.y[0] : p.cp.x[0], c1 = side ? p.cp.y[1] : p.cp.x[1];
            float a = __fsub_rn(c1, c0);
            a = (a != 0.0f) ? __fdiv_rn(1.0f, a) : 0.0f;
            bool ok = true;
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + (int)lane;
                if (k < n) {
                    // grid lines are small integers: g0 + k * d is exact, like the reference's repeated additions
                    const float v = __fmul_rn(__fsub_rn(__fadd_rn(g0, __fmul_rn((float)k, d)), c0), a);
                    float t_before = p.t0_ms;  // what the sequential walk would have clamped against
                    if (k > 0) {
                        const float vp = __fmul_rn(__fsub_rn(__fadd_rn(g0, __fmul_rn((float)(k - 1), d)), c0), a);
                        const float tp = (p.t1_ms < vp) ? p.t1_ms : vp;
                        t_before = u2f((f2u(tp) & 0xFFFFFFFCu) | (uint32_t)side);
                    }
                    if (v < t_before) ok = false;  // the clamp max(v, t_min) would have been active: order matters
                    const float t = (p.t1_ms < v) ? p.t1_ms : v;
                    out[k] = (f2u(t) & 0xFFFFFFFCu) | (uint32_t)side;
                }
            }
            done = __all_sync(0xFFFFFFFFu, ok);
        }
        if (SLPR_LONG_WARP && !done && (p.type == T_CUBIC || (FULL && (p.type == T_QUADRIC || p.type == T_ARC)))) {
            // in order, each crossing by the whole warp (warp_bisect): the evaluators are solve_crossing's
            const float c0 = side ? p.cp.y[0] : p.cp.x[0], c1 = side ? p.cp.y[1] : p.cp.x[1], c2 = side ? p.cp.y[2] : p.cp.x[2];
            const float c3 = side ? p.cp.y[3] : p.cp.x[3], w = p.cp.x[3];
            const float d01 = __fsub_rn(c1, c0), d12 = __fsub_rn(c2, c1), d23 = __fsub_rn(c3, c2);
            auto cubic = [&](float t) {
                const float a0 = madd_t<FMA>(t, d01, c0), a1 = madd_t<FMA>(t, d12, c1), a2 = madd_t<FMA>(t, d23, c2);
                const float b0 = lerp_t<FMA>(a0, a1, t), b1 = lerp_t<FMA>(a1, a2, t);
                return lerp_t<FMA>(b0, b1, t);
            };
            auto quadric = [&](float t) { return eval_quadric<FMA>(c0, c1, c2, t); };
            auto arc = [&](float t) { return eval_arc<FMA>(c0, c1, c2, w, t); };
            float t_prev = p.t0_ms, g = g0;
            for (int k = 0; k < n; ++k) {
                const float cst = g;
                g = __fadd_rn(g, d);
                float ts;
                if (FULL && p.type == T_ARC) ts = warp_bisect(arc, t_prev, p.t1_ms, cst, tl, lane);
                else if (FULL && p.type == T_QUADRIC) ts = warp_bisect(quadric, t_prev, p.t1_ms, cst, tl, lane);
                else ts = warp_bisect(cubic, t_prev, p.t1_ms, cst, tl, lane);
                const uint32_t tg = (f2u(ts) & 0xFFFFFFFCu) | (uint32_t)side;
                if (lane == 0) out[k] = tg;
                t_prev = u2f(tg);
            }
            done = true;
        }
        if (!done && lane == 0) {  // in order by one lane (lines when a clamp was active; other types)
            float t_prev = p.t0_ms, g = g0;
            for (int k = 0; k < n; ++k) {
                const float cst = g;
                g = __fadd_rn(g, d);
                const float ts = solve_crossing<FULL, FMA>(p.type, p.cp, side, t_prev, p.t1_ms, cst);
                const uint32_t tg = (f2u(ts) & 0xFFFFFFFCu) | (uint32_t)side;
                out[k] = tg;
                t_prev = u2f(tg);
            }
        }
    }
}

// One piece, by a warp (BLOCKWIDE false; `tid` = lane) or by a whole block (true; `tid` = thread): merged order, then fragments.
template <bool FULL, bool FMA, bool BLOCKWIDE>
__device__ __forceinline__ void emit_long_piece(const LongPiece &p, int tid, uint32_t *st_chain, uint32_t *st_merged, int stage_words,
                                                const FragEnv &env, FrameCounters *__restrict__ ctr, const LongScratch &sc, const KeyLayout &L,
                                                uint64_t *__restrict__ key64, uint32_t *__restrict__ val, const FragTaps &taps,
                                                int2 *__restrict__ inter, float2 *__restrict__ boundary, uint4 *__restrict__ fixlist) {
    constexpr int G = BLOCKWIDE ? LONG_EMIT_THREADS : 32;
    auto group_sync = [] { if (BLOCKWIDE) __syncthreads(); else __syncwarp(); };
    const int n_loop = p.n_x + p.n_y + 1;
    // the two chains as the reference's merge sees them: the piece's start record leads the y chain, or the x chain
    // when there is no y crossing (MI1:321-333); an exhausted chain reads 2.0 tagged with its side (MI1:340-359)
    const bool x_leads = p.n_x > 0 && p.n_y == 0;
    const int len_x = p.n_x + (x_leads ? 1 : 0), len_y = p.n_y + (x_leads ? 0 : 1);
    const uint32_t *cx = sc.chain + p.pcnt, *cy = sc.chain + p.pcnt + p.n_x;
    uint32_t *mg = sc.merged + p.pcnt;
    // The ranks below are binary searches, one per record, and every record is read twice more when the fragments
    // are formed: chains and merged order live in shared memory when the piece fits (any piece of a 4K frame does).
    group_sync();  // the previous piece is done with the stage
    if (n_loop <= stage_words) {
        for (int i = tid; i < p.n_x + p.n_y; i += G) st_chain[i] = cx[i];
        cx = st_chain; cy = st_chain + p.n_x; mg = st_merged;
    }
    group_sync();
    const uint32_t t0_bits = f2u(p.t0_ms);
    auto get_x = [&](int i) { return x_leads ? (i == 0 ? t0_bits : cx[i - 1]) : cx[i]; };
    auto get_y = [&](int j) { return x_leads ? cy[j] : (j == 0 ? t0_bits : cy[j - 1]); };
    // Head by head (MI1:321-359), an element is emitted once its predecessors in its chain are out and it is the smaller
    // head: that is a merge by the running maximum of each chain. The chains ascend — every bracket starts at the previous
    // crossing — except that the first crossing of the leading chain can read below the piece's start record by its tag
    // bits (a piece that starts on a grid line: the tiger's full-height line at 4K), so the leading chain is ranked by
    // max(element, start record). Anything else out of order (not observed) goes the sequential way below.
    const float head = u2f(t0_bits);
    auto eff_x = [&](int i) { const float v = u2f(get_x(i)); return (x_leads && v < head) ? head : v; };
    auto eff_y = [&](int j) { const float v = u2f(get_y(j)); return (!x_leads && v < head) ? head : v; };
    bool mono = true;
    for (int i = tid + 1; i < len_x; i += G) mono = mono && !(eff_x(i) < eff_x(i - 1));
    for (int j = tid + 1; j < len_y; j += G) mono = mono && !(eff_y(j) < eff_y(j - 1));
    if (BLOCKWIDE ? __syncthreads_and(mono) : __all_sync(0xFFFFFFFFu, mono)) {  // a record's place is its index plus its rank in the other chain
        for (int i = tid; i < len_x; i += G) {
            const float v = eff_x(i);
            int lo = 0, hi = len_y;  // y records strictly before v (x goes first on ties: `tx <= ty`)
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (eff_y(mid) < v) lo = mid + 1; else hi = mid; }
            mg[i + lo] = get_x(i);
        }
        for (int j = tid; j < len_y; j += G) {
            const float v = eff_y(j);
            int lo = 0, hi = len_x;  // x records at or before v
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (eff_x(mid) <= v) lo = mid + 1; else hi = mid; }
            mg[j + lo] = get_y(j);
        }
    } else if (tid == 0) {  // head by head, as the reference does
#ifdef SLPR_DEBUG_MONO
        printf("non-monotone piece: n_x %d n_y %d blockwide %d\n", p.n_x, p.n_y, (int)BLOCKWIDE);
#endif
        int i = 0, j = 0;
        for (int k = 0; k < n_loop; ++k) {
            const uint32_t hx = i < len_x ? get_x(i) : (f2u(2.0f) | 0u), hy = j < len_y ? get_y(j) : (f2u(2.0f) | 1u);
            if (u2f(hx) <= u2f(hy)) { mg[k] = hx; ++i; } else { mg[k] = hy; ++j; }
        }
    }
    group_sync();
    if (inter && tid == 0) {  // debug tap (MI1:361-375): the records with their tag bits merged across equal parameters
        int i_inte_last = (int)f2u(-1.0f);
        for (int k = 0; k < n_loop; ++k) {
            int i_out = (int)mg[k];
            if ((mg[k] & 0xFFFFFFFCu) == ((uint32_t)i_inte_last & 0xFFFFFFFCu)) {
                i_out |= i_inte_last;
                inter[p.pcnt + k - 1] = make_int2((int)p.c, i_out);
            }
            inter[p.pcnt + k] = make_int2((int)p.c, i_out);
            i_inte_last = i_out;
        }
    }
    // ---- fragments: record k closes the fragment that record k - 1 opened (gen_fragment fused, as in k_walk)
    for (int k = 1 + tid; k <= n_loop; k += G) {
        {
            float ta = u2f(mg[k - 1] & 0xFFFFFFFCu);
            ta = (ta < 0.0f) ? 0.0f : ta;
            float tb = u2f((k < n_loop ? mg[k] : f2u(p.t1_ms)) & 0xFFFFFFFCu);  // k == n_loop: the fragment across the piece boundary
            tb = (tb < 0.0f) ? 0.0f : tb;
            float ax, ay, bx, by;
            eval_point<FULL, FMA>(p.type, p.cp, ta, ax, ay);
            eval_point<FULL, FMA>(p.type, p.cp, tb, bx, by);
            uint64_t kk; uint32_t vv;
            make_fragment(env, L, p.pcnt + k - 1, p.pidx, p.rule_bit, ta, tb, ax, ay, bx, by, kk, vv, taps);
            key64[p.pcnt + k - 1] = kk;
            val[p.pcnt + k - 1] = vv;
        }
    }
    if (tid == 0) {  // the boundary-repair bookkeeping of k_walk
        const uint32_t first_bits = mg[0], last_bits = mg[n_loop - 1];
        const bool unordered = p.piece > 0 && (f2u(p.t1_ms) & 0xFFFFFFFCu) <= (f2u(p.t0_ms) | 3u);
        if (unordered || p.keep_last) boundary[5 * p.c + p.piece] = make_float2(u2f(first_bits), u2f(last_bits));
        if (p.piece > 0 && (first_bits & 0xFFFFFFFCu) != (f2u(p.t0_ms) & 0xFFFFFFFCu)) {
            if (unordered) fixlist[atomicAdd(&ctr->n_fix, 1)] = make_uint4(p.c, p.piece, (uint32_t)p.pcnt, 0u);
            else ctr->fix_missed = 1;
        }
    }
}

// Pieces of more than LONG_WARP_RECORDS records — all in the top length bucket, the first n_top pieces — are emitted by
// a whole block each; the others by a warp each (a block would idle on their 20 to 250 records and its barriers would
// only add latency: tiger at 4K has ~4000 of them).
template <bool FULL, bool FMA>
__global__ void __launch_bounds__(LONG_EMIT_THREADS) k_long_emit(const FrameParams *__restrict__ P, const PieceRec *__restrict__ pieces,
                                                   FrameCounters *__restrict__ ctr, int capacity, LongScratch sc, KeyLayout L,
                                                   uint64_t *__restrict__ key64, uint32_t *__restrict__ val, FragTaps taps,
                                                   int2 *__restrict__ inter, float2 *__restrict__ boundary, uint4 *__restrict__ fixlist) {
    __shared__ uint32_t s_chain[LONG_STAGE_WORDS], s_merged[LONG_STAGE_WORDS];
    if (ctr->n_fragments > capacity) return;
    const uint32_t n_long = (uint32_t)ctr->n_long, n_top = (uint32_t)ctr->n_top;
    const FragEnv env = load_frag_env(P);
    for (uint32_t idx = blockIdx.x; idx < n_top; idx += gridDim.x) {
        const LongPiece p = load_long_piece(pieces, idx);
        if (p.n_x + p.n_y + 1 <= LONG_WARP_RECORDS) continue;  // (block-uniform)
        emit_long_piece<FULL, FMA, true>(p, (int)threadIdx.x, s_chain, s_merged, LONG_STAGE_WORDS, env, ctr, sc, L, key64, val, taps, inter, boundary, fixlist);
    }
    __syncthreads();
    constexpr int WARPS = LONG_EMIT_THREADS / 32, SLICE = LONG_STAGE_WORDS / WARPS;
    static_assert(SLICE >= LONG_WARP_RECORDS, "a warp's slice of the stage holds its piece");
    const uint32_t warp = threadIdx.x >> 5, n_warps = gridDim.x * WARPS;
    for (uint32_t idx = blockIdx.x * WARPS + warp; idx < n_long; idx += n_warps) {
        const LongPiece p = load_long_piece(pieces, idx);
        if (idx < n_top && p.n_x + p.n_y + 1 > LONG_WARP_RECORDS) continue;
        emit_long_piece<FULL, FMA, false>(p, (int)lane_id(), s_chain + warp * SLICE, s_merged + warp * SLICE, SLICE, env, ctr, sc, L, key64, val, taps, inter, boundary, fixlist);
    }
}

// ------------------------------------------------------------------------------------------------
// One thread per listed piece (curve, piece >= 1, its first record): the fragment between the last
// record of the piece before it and its own first record, from the parameters both pieces recorded. Runs after
// k_walk, when every piece of the frame has left its boundary parameters.
template <bool FMA>
__global__ void __launch_bounds__(256) k_piece_fix(const FrameParams *__restrict__ P, const uint32_t *__restrict__ curve_type,
                                                   const uint32_t *__restrict__ curve_pos_map,
                                                   const uint32_t *__restrict__ curve_path, const uint32_t *__restrict__ fill_rule,
                                                   const float2 *__restrict__ tpos, const FrameCounters *__restrict__ ctr, int capacity,
                                                   const float2 *__restrict__ boundary, const uint4 *__restrict__ fixlist, KeyLayout L,
                                                   uint64_t *__restrict__ key64, uint32_t *__restrict__ val, FragTaps taps, FullRvg full) {
    if (ctr->n_fragments > capacity) return;
    const int n_fix = ctr->n_fix;
    if (n_fix == 0) return;  // the common case
    const FragEnv env = load_frag_env(P);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_fix; i += gridDim.x * blockDim.x) {
        const uint4 e = fixlist[i];
        const uint32_t c = e.x, piece = e.y;
        const uint32_t type = curve_type[c];
        const uint32_t pidx = curve_path[c];
        const uint32_t rule_bit = fill_rule[pidx] == 1u ? 1u : 0u;
        CurvePts cp;
        load_points(type, curve_pos_map[c], tpos, cp);
        full.stash_weight(type, c, cp);
        const int f = (int)e.z - 1;  // last record of the piece before: a curve's records are consecutive
        // GF:99-104: t0 from that record, t1 from the next record of the curve
        float t0 = u2f(f2u(boundary[5 * c + piece - 1].y) & 0xFFFFFFFCu);
        float t1 = u2f(f2u(boundary[5 * c + piece].x) & 0xFFFFFFFCu);
        t0 = (t0 < 0.0f) ? 0.0f : t0;
        t1 = (t1 < 0.0f) ? 0.0f : t1;
        float ax, ay, bx, by;
        if (full.on()) { eval_point<true, FMA>(type, cp, t0, ax, ay); eval_point<true, FMA>(type, cp, t1, bx, by); }
        else { eval_point<false, FMA>(type, cp, t0, ax, ay); eval_point<false, FMA>(type, cp, t1, bx, by); }
        emit_fragment(env, L, f, pidx, rule_bit, t0, t1, ax, ay, bx, by, key64, val, taps);
    }
}

}  // namespace slpr
