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
// Monotone map float -> uint32 (a < b  <=>  float_order(a) < float_order(b); NaNs land at the two ends).
__device__ __forceinline__ uint32_t float_order(float f) {
    const uint32_t b = f2u(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// transform_pos.comp:41-82 for one point: the transformed position and the point's region bit for its path's mask
struct PointXform {
    float m0x, m0y, m0z, m0w, m1x, m1y, m1z, m1w, m3x, m3y, m3z, m3w, w, h;
    __device__ __forceinline__ explicit PointXform(const FrameParams *__restrict__ P)
        : m0x(P->rows[0]), m0y(P->rows[1]), m0z(P->rows[2]), m0w(P->rows[3]), m1x(P->rows[4]), m1y(P->rows[5]), m1z(P->rows[6]),
          m1w(P->rows[7]), m3x(P->rows[12]), m3y(P->rows[13]), m3z(P->rows[14]), m3w(P->rows[15]),
          w((float)P->width), h((float)P->height) {}  // TP:8 (floats), SR.cpp:1153-1154
    template <bool FMA = false>
    __device__ __forceinline__ uint32_t apply(float2 p, float2 &out) const {
        // dot(vec4(x,y,0,1), m) evaluated left to right (TP:41-46)
        float ox = madd_t<FMA>(1.0f, m0w, madd_t<FMA>(0.0f, m0z, madd_t<FMA>(p.y, m0y, __fmul_rn(p.x, m0x))));
        float oy = madd_t<FMA>(1.0f, m1w, madd_t<FMA>(0.0f, m1z, madd_t<FMA>(p.y, m1y, __fmul_rn(p.x, m1x))));
        const float ow = madd_t<FMA>(1.0f, m3w, madd_t<FMA>(0.0f, m3z, madd_t<FMA>(p.y, m3y, __fmul_rn(p.x, m3x))));
        ox = __fdiv_rn(ox, ow);  // TP:53-54
        oy = __fdiv_rn(oy, ow);
        out = make_float2(ox, oy);
        const int xf = ox < 0 ? 0 : (ox < w ? 1 : 2);  // TP:67-68
        const int yf = oy < 0 ? 0 : (oy < h ? 1 : 2);
        switch ((yf << 4) | xf) {  // TP:71-82
            case 0x00: return 0x10000000u;
            case 0x01: return 0x01000000u;
            case 0x02: return 0x00100000u;
            case 0x10: return 0x00010000u;
            case 0x11: return 0x10000001u;
            case 0x12: return 0x00001000u;
            case 0x20: return 0x00000100u;
            case 0x21: return 0x00000010u;
            case 0x22: return 0x00000001u;
            default: return 0u;
        }
    }
};

#ifndef SLPR_XF_PREFETCH
#define SLPR_XF_PREFETCH 1
#endif
template <bool FMA>
__global__ void __launch_bounds__(256) k_transform(const FrameParams *__restrict__ P, uint32_t n_points,
                                                   const float2 *__restrict__ pos,
                                                   const uint32_t *__restrict__ pos_path,
                                                   float2 *__restrict__ tpos, int *__restrict__ path_visible,
                                                   const uint8_t *__restrict__ path_live) {
    const PointXform xf(P);
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n_points + 31u) & ~31u;  // keep whole warps in the loop for the shuffles
#if SLPR_XF_PREFETCH
    // The point and its path index are fetched together, one trip ahead: the kernel used to wait twice per point —
    // for the point, and again, behind the divisions, for the path index the match needs (ncu: 2 x 30 % of its stalls).
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float2 p_nx = make_float2(0.f, 0.f);
    uint32_t pid_nx = 0xFFFFFFFFu;
    if (i < n_points) { pid_nx = pos_path[i]; p_nx = pos[i]; }
    for (; i < n_round; i += stride) {
        const float2 p = p_nx;
        uint32_t pidx = pid_nx;
        bool live = i < n_points;
        const uint32_t i_nx = i + stride;
        pid_nx = 0xFFFFFFFFu;
        if (i_nx < n_points) { pid_nx = pos_path[i_nx]; p_nx = pos[i_nx]; }
        uint32_t flag = 0;
        if (live && path_live) {  // band mode: the points of a path that cannot reach the band are not needed
            live = path_live[pidx] != 0;
            if (!live) pidx = 0xFFFFFFFFu;
        }
        if (live) {
            float2 o;
            flag = xf.apply<FMA>(p, o);
            tpos[i] = o;
        }
#else
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool live = i < n_points;
        uint32_t flag = 0, pidx = 0xFFFFFFFFu;
        if (live && path_live) {  // band mode: the points of a path that cannot reach the band are not needed
            pidx = pos_path[i];
            live = path_live[pidx] != 0;
            if (!live) pidx = 0xFFFFFFFFu;
        }
        if (live) {
            float2 o;
            flag = xf.apply<FMA>(pos[i], o);
            pidx = pos_path[i];
            tpos[i] = o;
        }
#endif
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, pidx);
        const uint32_t red = __reduce_or_sync(peers, flag);
        const bool leader = live && (lane_id() == (uint32_t)(__ffs(peers) - 1));
        if (leader) atomicOr(&path_visible[pidx], (int)red);
    }
}

// Note: x / FRAG_SIZE in the shaders is a division by 2.0: written as x * 0.5f below (bit-identical, the
// scaling is exact for every float) to keep IEEE division sequences out of the hot loops.
// ------------------------------------------------------------------------------------------------
// helpers shared by K2 and K4
// ------------------------------------------------------------------------------------------------
// make_intersection_0.comp:76-129
template <bool FMA = false>
__device__ __forceinline__ void solve_quad(float a, float b, float c, float &r0, float &r1) {
    if (a == 0) {
        const float x = __fdiv_rn(-c, b);
        r0 = x; r1 = x;
        return;
    }
    const float A = a, B = __fmul_rn(b, 0.5f), C = c;
    float tx = 0.f, ty = 0.f;
    const float R = FMA ? __fmaf_rn(B, B, -__fmul_rn(A, C)) : __fsub_rn(__fmul_rn(B, B), __fmul_rn(A, C));
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

// f-1 (SLPR_FLAG_FULL_RVG): what the real QUADRIC / ARC arithmetic needs beyond the reference's inputs — the weight of
// every ARC's middle control point (it rides in the unused fourth slot of the curve's points).
struct FullRvg {
    const float *curve_weight;  // nullptr: mode off (reference behaviour: the TODO arms)
    __device__ __forceinline__ bool on() const { return curve_weight != nullptr; }
    __device__ __forceinline__ void stash_weight(uint32_t type, uint32_t curve, CurvePts &c) const {
        if (curve_weight && type == T_ARC) { const float w = curve_weight[curve]; c.x[3] = w; c.y[3] = w; }
    }
};

// ------------------------------------------------------------------------------------------------
// K2: make_intersection_0.comp:226-410. One thread per curve: monotonic cut parameters (<=4),
// literal partial insertion sort (including the `float t2 = q3;` slip at MI0:340), and the number
// of 2-px grid crossings per monotone piece. Band mode (new): a curve whose control-point box
// misses the band by more than one pixel emits nothing (exact).
// New: every monotone piece is also filed under a length bucket (its record count, capped) and gets
// a rank inside the bucket (block-aggregated atomics), so that the walk kernel can process pieces
// of equal length in the same warp: slots[5c+k] = bucket << 26 | rank.
// ------------------------------------------------------------------------------------------------
constexpr int WALK_BUCKETS = 64;
// Pieces are laid out for the walk longest first. On big scenes the short pieces (the bulk) are in addition grouped
// by window of consecutive curves: first every piece of more than `long_min` records, by length; then window
// after window, inside a window by length again. A window's pieces — and so its fragments, which neighbour each
// other in the fragment arrays — are then walked within a few tens of microseconds of each other, so the 32-byte
// sectors two pieces share are completed in L2 instead of being written to HBM half empty and read back.
// Measured (profiles/README.md): k_walk's DRAM traffic 666 -> 455 MB per 4K frame with every piece windowed, but
// the kernel is bound by instruction issue, and groups of long pieces that start with their window instead of
// with the frame leave a tail: +2..8 % at 1 M curves, -2 % at 4 M curves (where a frame's fragments are 13 x
// the L2). So the host windows the pieces of up to WALK_LONG records on scenes of more than WALK_WINDOWED_CURVES
// curves and keeps one global longest-first order (long_min = 0) otherwise.
constexpr int WALK_MAX_WINDOWS = 64;
#ifndef SLPR_WALK_WINDOW
#define SLPR_WALK_WINDOW 131072
#endif
constexpr int WALK_WINDOW_CURVES = SLPR_WALK_WINDOW;
#ifndef SLPR_WALK_LONG
#define SLPR_WALK_LONG 16
#endif
constexpr uint32_t WALK_LONG = SLPR_WALK_LONG;
#ifndef SLPR_WALK_WINDOWED_CURVES
#define SLPR_WALK_WINDOWED_CURVES 2000000
#endif
constexpr long long WALK_WINDOWED_CURVES = SLPR_WALK_WINDOWED_CURVES;
constexpr int WALK_VBUCKETS_MAX = (WALK_MAX_WINDOWS + 1) * WALK_BUCKETS;
struct PieceLayout {
    uint32_t n_blocks;       // blocks of k_monotonize_count; block b ranks the work items [b * ipb, (b + 1) * ipb)
    uint32_t n_windows;      // windows of blocks_per_window consecutive blocks
    uint32_t blocks_per_window;
    uint32_t long_min;       // pieces of more than this many records are laid out before all windows (0: every piece)
    __device__ __forceinline__ uint32_t items_per_block(uint32_t n_work) const {
        return (((n_work + n_blocks - 1) / n_blocks) + 31u) & ~31u;  // whole warps; every block gets work
    }
    __device__ __forceinline__ uint32_t n_vbuckets() const { return (n_windows + 1u) * WALK_BUCKETS; }
    // virtual bucket of (window, length bucket); pieces are laid out in DESCENDING virtual bucket order
    __device__ __forceinline__ uint32_t vbucket_of_window(uint32_t win, uint32_t bucket) const {
        return bucket > long_min ? n_windows * WALK_BUCKETS + bucket : (n_windows - 1u - win) * WALK_BUCKETS + bucket;
    }
    __device__ __forceinline__ uint32_t vbucket(uint32_t block, uint32_t bucket) const {
        return vbucket_of_window(block / blocks_per_window, bucket);
    }
};

// Band mode: the per-curve kernels (k_monotonize_count, k_piece_emit, k_piece_fix) walk a compacted list of
// the curves whose path comes near the band instead of all curves — with 8 bands 7 of 8 curves are dead, and
// scattered among live ones they would leave the warps of those kernels 1/8 full. Work item w -> curve.
struct LiveCurves {
    const uint32_t *list;         // nullptr: every curve is live (full frame)
    const FrameCounters *ctr;
    __device__ __forceinline__ uint32_t count(uint32_t n_curves) const { return list ? (uint32_t)ctr->n_live : n_curves; }
    __device__ __forceinline__ uint32_t curve(uint32_t w) const { return list ? list[w] : w; }
};

// Band mode, per path: can any control point of the path come near the band? Decided from the path's box in
// OBJECT space (static, computed by slpr_load_scene): the four corners go through the frame's matrix exactly
// like points do; a projective map with w > 0 keeps the box convex, so every transformed control point lies
// between the smallest and largest corner row (up to rounding, covered by the margin). Conservative: a corner
// with w <= 0 or a NaN keeps the path alive. Dead paths skip k_transform and every per-curve kernel.
struct BandCull {
    float m1x, m1y, m1z, m1w, m3x, m3y, m3z, m3w, lo, hi;
    __device__ __forceinline__ explicit BandCull(const FrameParams *__restrict__ P)
        : m1x(P->rows[4]), m1y(P->rows[5]), m1z(P->rows[6]), m1w(P->rows[7]), m3x(P->rows[12]), m3y(P->rows[13]), m3z(P->rows[14]),
          m3w(P->rows[15]),
          lo((P->band_y0 > 0) ? (float)(P->band_y0 - 3) : -3.0e38f),       // per-curve margin 1 px + 2 px for rounding
          hi((P->band_y1 < P->height) ? (float)(P->band_y1 + 3) : 3.0e38f) {}
    __device__ __forceinline__ bool alive(float4 b) const {  // xmin, ymin, xmax, ymax; an empty path has xmin > xmax
        if (!(b.x <= b.z)) return false;  // no points, hence no curves
        float ymin = 3.0e38f, ymax = -3.0e38f;
        bool safe = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = (k & 1) ? b.z : b.x, y = (k & 2) ? b.w : b.y;
            const float oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, m1x), __fmul_rn(y, m1y)), __fmul_rn(0.0f, m1z)), __fmul_rn(1.0f, m1w));
            const float ow = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, m3x), __fmul_rn(y, m3y)), __fmul_rn(0.0f, m3z)), __fmul_rn(1.0f, m3w));
            const float v = __fdiv_rn(oy, ow);
            safe = safe && (ow > 0.0f) && (v == v);
            ymin = fminf(ymin, v);
            ymax = fmaxf(ymax, v);
        }
        return !safe || !(ymax < lo || ymin >= hi);
    }
};

// (both band front ends also list the live paths — in ascending order inside a warp's 32 — for the segmented sort,
// which then visits 1/n_bands of the paths instead of all of them)
__global__ void __launch_bounds__(256) k_path_cull(const FrameParams *__restrict__ P, uint32_t n_paths,
                                                   const float4 *__restrict__ path_obj_box, uint8_t *__restrict__ path_live,
                                                   uint32_t *__restrict__ live_paths, FrameCounters *__restrict__ ctr) {
    const BandCull cull(P);
    const uint32_t n_round = (n_paths + 31u) & ~31u;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += gridDim.x * blockDim.x) {
        const bool alive = p < n_paths && cull.alive(path_obj_box[p]);
        if (p < n_paths) path_live[p] = alive ? 1 : 0;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, alive);
        if (m) {
            uint32_t base = 0;
            if (lane_id() == 0) base = (uint32_t)atomicAdd(&ctr->n_live_paths, __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (alive) live_paths[base + __popc(m & lanemask_lt())] = p;
        }
    }
}

// Band mode when the scene's points are grouped by path (they are whenever the scene comes from loadVG): cull,
// transform and live-curve list in one pass that only touches the paths that can reach the band — with 8 bands 7
// of 8 paths are dead, and reading a path index per point and per curve of the whole scene to find that out cost
// more than the band's own points and curves (0.16 of 1.55 ms per band at 16K). One thread tests one path; the
// warp then takes its live paths one after the other: transforms the path's points (k_transform's arithmetic),
// stores the path's visibility mask (the warp is its only writer) and appends its curves to the live list.
template <bool FMA>
__global__ void __launch_bounds__(256) k_band_paths(const FrameParams *__restrict__ P, uint32_t n_paths,
                                                    const float4 *__restrict__ path_obj_box,
                                                    const uint32_t *__restrict__ path_first_point,
                                                    const uint32_t *__restrict__ path_first_curve, const float2 *__restrict__ pos,
                                                    float2 *__restrict__ tpos, int *__restrict__ path_visible,
                                                    uint32_t *__restrict__ list, FrameCounters *__restrict__ ctr,
                                                    uint32_t *__restrict__ live_paths) {
    const BandCull cull(P);
    const PointXform xf(P);
    const uint32_t lane = lane_id();
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_groups = (n_paths + 31u) >> 5;
    for (uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
        const uint32_t p = g * 32u + lane;
        uint32_t pt0 = 0, npt = 0, c0 = 0, ncv = 0;
        bool alive = false;
        if (p < n_paths && cull.alive(path_obj_box[p])) {
            alive = true;
            pt0 = path_first_point[p]; npt = path_first_point[p + 1] - pt0;
            c0 = path_first_curve[p]; ncv = path_first_curve[p + 1] - c0;
        }
        uint32_t live_mask = __ballot_sync(0xFFFFFFFFu, alive);
        if (!live_mask) continue;
        {   // the live paths, for the sort
            uint32_t pbase = 0;
            if (lane == 0) pbase = (uint32_t)atomicAdd(&ctr->n_live_paths, __popc(live_mask));
            pbase = __shfl_sync(0xFFFFFFFFu, pbase, 0);
            if (alive) live_paths[pbase + __popc(live_mask & lanemask_lt())] = p;
        }
        // the warp's curves take one run of the list
        uint32_t incl = ncv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t base = 0;
        if (lane == 0 && total) base = (uint32_t)atomicAdd(&ctr->n_live, (int)total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        const uint32_t my_off = base + incl - ncv;
        while (live_mask) {
            const int src = __ffs(live_mask) - 1;
            live_mask &= live_mask - 1;
            const uint32_t s_pt0 = __shfl_sync(0xFFFFFFFFu, pt0, src), s_npt = __shfl_sync(0xFFFFFFFFu, npt, src);
            const uint32_t s_c0 = __shfl_sync(0xFFFFFFFFu, c0, src), s_ncv = __shfl_sync(0xFFFFFFFFu, ncv, src);
            const uint32_t s_off = __shfl_sync(0xFFFFFFFFu, my_off, src);
            for (uint32_t i = lane; i < s_ncv; i += 32) list[s_off + i] = s_c0 + i;
            uint32_t flags = 0;
            for (uint32_t i = lane; i < s_npt; i += 32) {
                float2 o;
                flags |= xf.apply<FMA>(pos[s_pt0 + i], o);
                tpos[s_pt0 + i] = o;
            }
            flags = __reduce_or_sync(0xFFFFFFFFu, flags);
            if (lane == 0) path_visible[g * 32u + (uint32_t)src] = (int)flags;
        }
    }
}

// Compacts the curves of live paths into `list` (arbitrary order) and zeroes the count of the others. A warp
// classifies 32 x LIVE_CHUNK consecutive curves and reserves its part of the list with ONE atomic.
constexpr int LIVE_CHUNK = 16;
__global__ void __launch_bounds__(256) k_band_live(uint32_t n_curves, const uint32_t *__restrict__ curve_path,
                                                   const uint8_t *__restrict__ path_live, int *__restrict__ count,
                                                   uint32_t *__restrict__ list, FrameCounters *__restrict__ ctr) {
    const uint32_t lane = lane_id();
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_chunks = (n_curves + 32 * LIVE_CHUNK - 1) / (32 * LIVE_CHUNK);
    for (uint32_t ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < n_chunks; ch += warps) {
        const uint32_t c0 = ch * 32 * LIVE_CHUNK;
        uint32_t mask[LIVE_CHUNK];
        uint32_t total = 0;
#pragma unroll
        for (int j = 0; j < LIVE_CHUNK; ++j) {
            const uint32_t c = c0 + j * 32 + lane;
            bool alive = false;
            if (c < n_curves) {
                alive = path_live[curve_path[c]] != 0;
                if (!alive) count[c] = 0;
            }
            mask[j] = __ballot_sync(0xFFFFFFFFu, alive);
            total += __popc(mask[j]);
        }
        if (total == 0) continue;
        uint32_t base = 0;
        if (lane == 0) base = (uint32_t)atomicAdd(&ctr->n_live, (int)total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
        for (int j = 0; j < LIVE_CHUNK; ++j) {
            if ((mask[j] >> lane) & 1u) list[base + __popc(mask[j] & ((1u << lane) - 1))] = c0 + j * 32 + lane;
            base += __popc(mask[j]);
        }
    }
}

#ifndef SLPR_MONO_PREFETCH
#define SLPR_MONO_PREFETCH 1
#endif
template <bool FMA>
__global__ void __launch_bounds__(256) k_monotonize_count(const FrameParams *__restrict__ P, uint32_t n_curves,
                                                          const uint32_t *__restrict__ curve_type,
                                                          const uint32_t *__restrict__ curve_pos_map,
                                                          const uint32_t *__restrict__ curve_path,
                                                          const float2 *__restrict__ tpos,
                                                          const int *__restrict__ path_visible,
                                                          float *__restrict__ cut_cache, int *__restrict__ count,
                                                          uint32_t *__restrict__ slots, uint32_t *__restrict__ block_cnt,
                                                          LiveCurves live, PieceLayout lay, FullRvg full) {
    // Pieces are ranked inside (block, length bucket) with shared-memory atomics only; the block's 64 counts
    // go to block_cnt[bucket][block] at the end and k_bucket_scan turns them into positions. (One global
    // atomicAdd per block iteration and bucket on 64 addresses serialised in L2: 0.1 ms at 4 M curves.)
    __shared__ uint32_t s_cnt[WALK_BUCKETS];
    if (threadIdx.x < WALK_BUCKETS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int width = P->width, height = P->height;
    const bool cull = P->cull != 0;
    // Only interior band edges cull: beyond the frame edges the reference still emits (invalid-key)
    // fragments whose winding deltas pair up across curves of a path (sampling rows -1 and H'+3).
    const float band_lo = (P->band_y0 > 0) ? (float)(P->band_y0 - 1) : -3.0e38f;
    const float band_hi = (P->band_y1 < P->height) ? (float)(P->band_y1 + 1) : 3.0e38f;
    const uint32_t n_work = live.count(n_curves);
    const uint32_t ipb = lay.items_per_block(n_work);  // a contiguous run of work items per block (PieceLayout)
    const uint32_t w_end = min(n_work, (blockIdx.x + 1u) * ipb);
#if SLPR_MONO_PREFETCH
    // Two levels of loads instead of four: what depends on the curve index alone (type, first point, path) is
    // fetched one trip ahead, and the path's visibility mask together with the points (it used to be asked for only
    // after the band test, behind the points: ncu showed a third of the kernel's stalls on these chains).
    uint32_t w = blockIdx.x * ipb + threadIdx.x;
    uint32_t c_nx = 0, type_nx = 0, po_nx = 0, path_nx = 0;
    if (w < w_end) { c_nx = live.curve(w); type_nx = curve_type[c_nx]; po_nx = curve_pos_map[c_nx]; path_nx = curve_path[c_nx]; }
    for (; w < w_end; w += blockDim.x) {
        const uint32_t c = c_nx, type = type_nx, cpath = path_nx;
        CurvePts cp;
        load_points(type, po_nx, tpos, cp);
        const int pvis = path_visible[cpath];
        if (w + blockDim.x < w_end) {
            c_nx = live.curve(w + blockDim.x); type_nx = curve_type[c_nx]; po_nx = curve_pos_map[c_nx]; path_nx = curve_path[c_nx];
        }
        full.stash_weight(type, c, cp);
#else
    for (uint32_t w = blockIdx.x * ipb + threadIdx.x; w < w_end; w += blockDim.x) {
        const uint32_t c = live.curve(w);
        const uint32_t type = curve_type[c];
        CurvePts cp;
        load_points(type, curve_pos_map[c], tpos, cp);
        full.stash_weight(type, c, cp);
        const uint32_t cpath = curve_path[c];
#endif
        uint32_t n_cuts = 0;
        bool culled = false;  // band mode: a curve whose control-point box misses the band is skipped like an invisible one
        // (only for curves that stay inside their control points' box: a type without a shader arm — and ARC without the
        // full-RVG arithmetic — is evaluated as the point (0, 0), a rational arc with a weight <= 0 leaves the hull)
        const bool boxed = type == T_LINE || type == T_CUBIC || type == T_QUADRIC || (full.on() && type == T_ARC && cp.x[3] > 0.0f);
        if (cull && boxed) {
            const uint32_t np = type & 7u;
            float ymin = cp.y[0], ymax = cp.y[0];
            for (uint32_t i = 1; i < 4; ++i)
                if (i < np) { ymin = fminf(ymin, cp.y[i]); ymax = fmaxf(ymax, cp.y[i]); }
            culled = (ymax < band_lo) || (ymin >= band_hi);
        }
#if SLPR_MONO_PREFETCH
        const bool visible = !culled && !path_invisible(pvis);  // MI0:260-261
#else
        const bool visible = !culled && !path_invisible(path_visible[cpath]);  // MI0:260-261
#endif
        float tq[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        if (visible) {
            if (type == T_CUBIC) {  // MI0:267-308 (LINE has no cuts; QUADRIC/ARC are TODO arms)
#pragma unroll
                for (int ax = 0; ax < 2; ++ax) {
                    const float x0 = ax ? cp.y[0] : cp.x[0], x1 = ax ? cp.y[1] : cp.x[1];
                    const float x2 = ax ? cp.y[2] : cp.x[2], x3 = ax ? cp.y[3] : cp.x[3];
                    float r0 = 0.f, r1 = 0.f;
                    const float a = madd_t<FMA>(3.0f, __fsub_rn(x1, x2), __fsub_rn(x3, x0));
                    const float b = __fmul_rn(2.0f, __fadd_rn(__fsub_rn(x0, x1), __fsub_rn(x2, x1)));
                    const float cc = __fsub_rn(x1, x0);
                    solve_quad<FMA>(a, b, cc, r0, r1);
                    if (r0 > 0.0f && r0 < 1.0f) { tq[n_cuts] = r0; ++n_cuts; }
                    if (r1 > 0.0f && r1 < 1.0f && r1 != r0) { tq[n_cuts] = r1; ++n_cuts; }
                }
            } else if (full.on() && (type == T_QUADRIC || type == T_ARC)) {  // f-1: zero of the derivative's numerator (oracle.c)
#pragma unroll
                for (int ax = 0; ax < 2; ++ax) {
                    const float p0 = ax ? cp.y[0] : cp.x[0], p1 = ax ? cp.y[1] : cp.x[1], p2 = ax ? cp.y[2] : cp.x[2];
                    const float w = (type == T_ARC) ? cp.x[3] : 1.0f;
                    const float A = __fmul_rn(w, __fsub_rn(p1, p0)), B = __fsub_rn(p2, p0), C = __fmul_rn(w, __fsub_rn(p2, p1));
                    const float a = __fadd_rn(__fsub_rn(A, B), C);
                    const float b = __fsub_rn(B, __fmul_rn(2.0f, A));
                    float r0 = 0.f, r1 = 0.f;
                    solve_quad<FMA>(a, b, A, r0, r1);
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

        float p0x = cp.x[0], p0y = cp.y[0];
        int pcnt = 0;
        for (uint32_t i = 0; i < n_cuts; ++i) {  // MI0:383-408
            const float t1 = tq[i];
            const float p1x = interp_full<FMA>(type, t1, cp.x[0], cp.x[1], cp.x[2], cp.x[3], 1.0f, full.on());
            const float p1y = interp_full<FMA>(type, t1, cp.y[0], cp.y[1], cp.y[2], cp.y[3], 1.0f, full.on());
            // get_xy_begin_end, MI0:167-183 (floor)
            const float xlo = (p0x <= p1x) ? p0x : p1x, xhi = (p0x <= p1x) ? p1x : p0x;
            const float ylo = (p0y <= p1y) ? p0y : p1y, yhi = (p0y <= p1y) ? p1y : p0y;
            int xb = f2i(__fmul_rn(floorf(__fmul_rn(xlo, 0.5f)), 2.0f)) + FRAG_SIZE;
            int xe = f2i(__fmul_rn(floorf(__fmul_rn(xhi, 0.5f)), 2.0f));
            int yb = f2i(__fmul_rn(floorf(__fmul_rn(ylo, 0.5f)), 2.0f)) + FRAG_SIZE;
            int ye = f2i(__fmul_rn(floorf(__fmul_rn(yhi, 0.5f)), 2.0f));
            const int nx = cut_range(width, xb, xe);
            const int ny = cut_range(height, yb, ye);
            pcnt += 1 + nx + ny;
            p0x = p1x; p0y = p1y;
            if (!culled) {  // file the piece under its length bucket (a scheduling hint only)
                const uint32_t b = (uint32_t)min(1 + nx + ny, WALK_BUCKETS - 1);
                slots[5 * c + i] = (b << 26) | atomicAdd(&s_cnt[b], 1u);  // rank inside (this block, bucket b)
            }
        }
        count[c] = culled ? 0 : pcnt;
    }
    __syncthreads();
    if (threadIdx.x < WALK_BUCKETS) block_cnt[threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x];
}

// block_cnt[bucket][block] -> exclusive prefix over the blocks of the same window — over all blocks for the long
// buckets — (in place) and vhist[virtual bucket] = number of pieces in it. Grid (WALK_BUCKETS, n_windows), 128 or
// 1024 threads.
__global__ void __launch_bounds__(1024) k_bucket_scan(uint32_t *__restrict__ block_cnt, PieceLayout lay, uint32_t *__restrict__ vhist) {
    __shared__ uint32_t s_w[32];
    const uint32_t n_warps = blockDim.x >> 5;  // 4 with many windows, 32 when one run covers every block
    const uint32_t bucket = blockIdx.x, win = blockIdx.y;
    const bool is_long = bucket > lay.long_min;
    uint32_t b0 = win * lay.blocks_per_window, b1 = min(lay.n_blocks, b0 + lay.blocks_per_window);
    if (is_long) {  // one run over all blocks, done by the block of window 0; the per-window slots stay empty
        if (threadIdx.x == 0) vhist[(lay.n_windows - 1u - win) * WALK_BUCKETS + bucket] = 0u;
        if (win != 0) return;
        b0 = 0; b1 = lay.n_blocks;
    } else if (win == 0 && threadIdx.x == 0) {
        vhist[lay.n_windows * WALK_BUCKETS + bucket] = 0u;  // the long slot of a short bucket
    }
    uint32_t *row = block_cnt + (size_t)bucket * lay.n_blocks;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (uint32_t base = b0; base < b1; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < b1) ? row[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += o;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t k = 0; k < n_warps; ++k) {
            if (k < warp) before += s_w[k];
            total += s_w[k];
        }
        if (i < b1) row[i] = carry + before + incl - v;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) vhist[lay.vbucket_of_window(win, bucket)] = carry;
}

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


// gen_fragment.comp:117-224 for one fragment: the curve piece between two consecutive intersection
// records (parameters t0 < t1, end points pf / pl already evaluated) lies in one 2x2 cell. Emits the
// sort input: compact key (path | row rank | cell x) and value
// (fragment index | fill rule of the path << 29 | (delta+1) << 30): the span kernel then needs no gather.
constexpr uint32_t VAL_INDEX_MASK = 0x1FFFFFFFu;

// the frame constants gen_fragment needs, read from FrameParams once per thread
struct FragEnv {
    int width, height, cull, band_y0, band_y1;
};
__device__ __forceinline__ FragEnv load_frag_env(const FrameParams *__restrict__ P) {
    return FragEnv{P->width, P->height, P->cull, P->band_y0, P->band_y1};
}

// make_fragment computes the pair; emit_fragment also stores it.
__device__ __forceinline__ void make_fragment(const FragEnv &P, const KeyLayout &L, int f, uint32_t pidx,
                                              uint32_t rule_bit,
                                              float t0, float t1, float pfx, float pfy, float plx, float ply,
                                              uint64_t &key_out, uint32_t &val_out, const FragTaps &taps) {
    bool valid = false;
    int pos_x = 0, pos_y = 0, wn = 0;
    if (t0 < t1) {  // GF:119
        const int width = P.width, height = P.height;
        const int raw_x = float2int_rd(__fmul_rn(__fmul_rn(__fadd_rn(pfx, plx), 0.5f), 0.5f)) * FRAG_SIZE;  // GF:169-170
        const int raw_y = float2int_rd(__fmul_rn(__fmul_rn(__fadd_rn(pfy, ply), 0.5f), 0.5f)) * FRAG_SIZE;
        pos_x = min(max(raw_x, -FRAG_SIZE), (int)(((uint32_t)width & 0xFFFFFFFEu) + FRAG_SIZE));  // GF:177-178
        pos_y = min(max(raw_y, -FRAG_SIZE), (int)(((uint32_t)height & 0xFFFFFFFEu) + FRAG_SIZE));
        valid = (uint32_t)raw_y < (uint32_t)height;  // GF:185-187
        const float wn_y = (float)(pos_y + 1);
        if (pfy == ply) wn = 0;  // GF:190-199
        else if (pfy < wn_y && wn_y <= ply) wn = -1;
        else if (ply < wn_y && wn_y <= pfy) wn = 1;
        if (P.cull) {  // band mode (new): every fragment belongs to exactly one band; rows outside the frame go to the
                       // band at that frame edge. A fragment of another band keeps its slot but carries nothing.
            const bool below = raw_y < P.band_y0 && P.band_y0 > 0;
            const bool above = raw_y >= P.band_y1 && P.band_y1 < height;
            if (below || above) { valid = false; wn = 0; }
        }
    }
    key_out = pack_key(L, pidx, valid, pos_x, pos_y);
    val_out = (uint32_t)f | (rule_bit << 29) | ((uint32_t)(wn + 1) << 30);
    if (taps.key32) {
        taps.key32[f] = valid ? (int)(((uint32_t)(pos_y + 0x7FFF) << 16) | ((uint32_t)(pos_x + 0x7FFF) & 0xFFFFu))
                              : (int)0xFFFEFFFEu;
        taps.path[f] = (int)pidx;
        taps.wind[f] = wn;
    }
}

__device__ __forceinline__ void emit_fragment(const FragEnv &P, const KeyLayout &L, int f, uint32_t pidx,
                                              uint32_t rule_bit,
                                              float t0, float t1, float pfx, float pfy, float plx, float ply,
                                              uint64_t *__restrict__ key64, uint32_t *__restrict__ val,
                                              const FragTaps &taps) {
    uint64_t k; uint32_t v;
    make_fragment(P, L, f, pidx, rule_bit, t0, t1, pfx, pfy, plx, ply, k, v, taps);
    key64[f] = k;
    val[f] = v;
}

// curve_interpolate of gen_fragment.comp:59-87 (default result cv0; only LINE and CUBIC evaluate)
template <bool FULL = false, bool FMA = false>
__device__ __forceinline__ void eval_point(uint32_t type, const CurvePts &cp, float t, float &ox, float &oy) {
    if (FULL && type == T_QUADRIC) {
        ox = eval_quadric<FMA>(cp.x[0], cp.x[1], cp.x[2], t); oy = eval_quadric<FMA>(cp.y[0], cp.y[1], cp.y[2], t);
    } else if (FULL && type == T_ARC) {
        ox = eval_arc<FMA>(cp.x[0], cp.x[1], cp.x[2], cp.x[3], t); oy = eval_arc<FMA>(cp.y[0], cp.y[1], cp.y[2], cp.y[3], t);
    } else if (type == T_CUBIC) {
        ox = cubic_eval<FMA>(cp.x[0], cp.x[1], cp.x[2], cp.x[3], t);
        oy = cubic_eval<FMA>(cp.y[0], cp.y[1], cp.y[2], cp.y[3], t);
    } else if (type == T_LINE) {
        ox = lerp_t<FMA>(cp.x[0], cp.x[1], t); oy = lerp_t<FMA>(cp.y[0], cp.y[1], t);
    } else if (type == T_QUADRIC) {
        ox = cp.x[0]; oy = cp.y[0];
    } else {
        ox = 0.f; oy = 0.f;  // cv0 is never loaded for other types (GF:132-156)
    }
}

}  // namespace slpr
