// raster.cuh — stage 5: draw records -> RGBA8 framebuffer. Replaces scanlinepr.vert/.frag and the
// fixed-function LINE_LIST raster (scanlinepr.vert:19-46; SR.cpp:611-656,883-895): every record
// covers pixels x in [X, X+width) on scanline rows {Y, Y+1}; writes are opaque and the later
// record wins; the clear colour is white (SR.cpp:622); image row = H-1-scanline row (VERT:44).
//
// All record coordinates are even, so coverage is resolved on the 2x2-pixel cell grid:
//   k_fill_cells    atomicMax(cell, record index + 1)  — deterministic "later record wins" (small frames); on big
//                   frames k_spans marks the cells itself, with path + 1 (spans.cuh)
//   k_resolve       cell -> colour of that record / path (or white), 2x2 pixels, 128-bit stores; also
//                   re-zeroes the cell grid for the next frame.
// The cell grid of a 4K frame is 8 MB and stays L2-resident between the two kernels.
#pragma once
#include "common.cuh"

namespace slpr {

#ifndef SLPR_FILL_NARROW
#define SLPR_FILL_NARROW 1 /* records up to this many cells are filled by their own thread (round 1, whole-warp spans: 4 -> 0.162, 8 -> 0.147, 16 -> 0.199 ms; with spans by groups of 8 lanes, k_spans at 512-thread tiles: 1 -> 0.2283, 2 -> 0.2296, 4 -> 0.2308, 8 -> 0.2334; at 128-thread tiles: 1 -> 0.2058, 2 -> 0.2091) */
#endif

#ifndef SLPR_FILL_WARP
#define SLPR_FILL_WARP 64 /* spans of more cells than this are filled by the whole warp (the big shapes of small scenes: test.rvg at 4K 0.141 -> 0.131 ms) */
#endif
#ifndef SLPR_FILL_GROUP
#define SLPR_FILL_GROUP 8 /* lanes that fill one wide span together (measured on the 1 M-curve 4K frame, k_spans, narrow = 8: 32 -> 0.2555, 16 -> 0.2357, 8 -> 0.2334, 4 -> 0.2506 ms) */
#endif
// The coverage marks of 32 records, one per lane (ncell cells from cell (cx0, cy), priority prio; alpha = the
// record's alpha byte, looked at only when BLEND): opaque records atomicMax their priority into the cells, narrow
// ones by their own lane, wide ones by the whole warp. BLEND: translucent records append list nodes instead; the
// warp allocates the nodes of its narrow records with one atomicAdd, those of a wide span with another.
template <bool BLEND>
__device__ __forceinline__ void mark_cells32(uint32_t *__restrict__ cells, int cw, int cx0, int ncell, int cy, uint32_t prio,
                                             uint32_t alpha, const BlendList &bl, FrameCounters *ctr, uint32_t lane) {
    bool soft = false;  // translucent
    if (BLEND) {
        if (alpha == 0u) ncell = 0;  // fully transparent: leaves no trace
        soft = ncell > 0 && alpha != 255u;
    }
    if (ncell > 0 && ncell <= SLPR_FILL_NARROW && !soft) {
        uint32_t *row = cells + (size_t)cy * cw + cx0;
        for (int c = 0; c < ncell; ++c) atomicMax(row + c, prio);
    }
    if (BLEND && __any_sync(0xFFFFFFFFu, soft && ncell <= SLPR_FILL_NARROW)) {
        const int need = (soft && ncell <= SLPR_FILL_NARROW) ? ncell : 0;
        int incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += o;
        }
        uint32_t base = 0;
        if (lane == 31) base = (uint32_t)atomicAdd(&ctr->n_blend_nodes, incl);
        base = __shfl_sync(0xFFFFFFFFu, base, 31) + (uint32_t)(incl - need);
        const size_t cell0 = (size_t)cy * cw + cx0;
        for (int c = 0; c < need; ++c) blend_append(bl, cell0 + c, base + (uint32_t)c, prio);
    }
    // wide spans: groups of lanes fill them, 32 / SLPR_FILL_GROUP at a time (with blending: the whole warp, one after the other).
    // ncu had the one-span-per-warp loop at 26 % of k_spans' instructions: 25 per span, most lanes idle on a 9-cell span.
    uint32_t wide = __ballot_sync(0xFFFFFFFFu, ncell > SLPR_FILL_NARROW);
    const uint32_t wide_soft = BLEND ? __ballot_sync(0xFFFFFFFFu, soft) : 0u;
    if constexpr (!BLEND) {
        // SLPR_FILL_GROUP lanes per span, 32 / SLPR_FILL_GROUP spans at a time (a group without a span of its own
        // repeats the first one: harmless); spans of more than SLPR_FILL_WARP cells by the whole warp
        constexpr int GROUPS = 32 / SLPR_FILL_GROUP;
        const int grp = (int)lane / SLPR_FILL_GROUP;
        uint32_t huge = __ballot_sync(0xFFFFFFFFu, ncell > SLPR_FILL_WARP);
        wide &= ~huge;
        while (huge) {
            const int src = __ffs(huge) - 1;
            huge &= huge - 1;
            const int s_cx0 = __shfl_sync(0xFFFFFFFFu, cx0, src);
            const int s_n = __shfl_sync(0xFFFFFFFFu, ncell, src);
            const int s_cy = __shfl_sync(0xFFFFFFFFu, cy, src);
            const uint32_t s_prio = __shfl_sync(0xFFFFFFFFu, prio, src);
            uint32_t *row = cells + (size_t)s_cy * cw + s_cx0;
            for (int c = (int)lane; c < s_n; c += 32) atomicMax(row + c, s_prio);
        }
        while (wide) {
            int src = __ffs(wide) - 1;
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                const int s = wide ? __ffs(wide) - 1 : src;
                wide &= wide - 1;
                if (g == grp) src = s;
            }
            const int s_cx0 = __shfl_sync(0xFFFFFFFFu, cx0, src);
            const int s_n = __shfl_sync(0xFFFFFFFFu, ncell, src);
            const int s_cy = __shfl_sync(0xFFFFFFFFu, cy, src);
            const uint32_t s_prio = __shfl_sync(0xFFFFFFFFu, prio, src);
            uint32_t *row = cells + (size_t)s_cy * cw + s_cx0;
            for (int c = (int)lane % SLPR_FILL_GROUP; c < s_n; c += SLPR_FILL_GROUP) atomicMax(row + c, s_prio);
        }
        return;
    }
    while (wide) {
        const int src = __ffs(wide) - 1;
        wide &= wide - 1;
        const int s_cx0 = __shfl_sync(0xFFFFFFFFu, cx0, src);
        const int s_n = __shfl_sync(0xFFFFFFFFu, ncell, src);
        const int s_cy = __shfl_sync(0xFFFFFFFFu, cy, src);
        const uint32_t s_prio = __shfl_sync(0xFFFFFFFFu, prio, src);
        if (BLEND && ((wide_soft >> src) & 1u)) {
            uint32_t base = 0;
            if (lane == 0) base = (uint32_t)atomicAdd(&ctr->n_blend_nodes, s_n);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            const size_t cell0 = (size_t)s_cy * cw + s_cx0;
            for (int c = (int)lane; c < s_n; c += 32) blend_append(bl, cell0 + c, base + (uint32_t)c, s_prio);
            continue;
        }
        uint32_t *row = cells + (size_t)s_cy * cw + s_cx0;
        for (int c = (int)lane; c < s_n; c += 32) atomicMax(row + c, s_prio);
    }
}

template <bool BLEND>
__global__ void __launch_bounds__(256) k_fill_cells(const FrameParams *__restrict__ P,
                                                    FrameCounters *__restrict__ ctr, int capacity,
                                                    const int4 *__restrict__ records, uint32_t *__restrict__ cells,
                                                    int cw, BlendList bl) {
    if (frame_void(ctr, capacity)) return;
    const int nrec = ctr->n_records;
    const int nround = (nrec + 31) & ~31;
    const int height = P->height;
    const uint32_t lane = lane_id();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nround; r += gridDim.x * blockDim.x) {
        int cx0 = 0, ncell = 0, cy = 0;
        uint32_t alpha = 255u;
        if (r < nrec) {
            const int4 rec = records[r];
            if (BLEND) alpha = (uint32_t)rec.z >> 24;
            const int X = rec.x & 0xFFFF, Y = rec.x >> 16;  // VERT:27
            if (Y >= 0 && Y < height) {
                cx0 = X >> 1;
                ncell = min((X + rec.y) >> 1, cw) - cx0;
                cy = Y >> 1;
            }
        }
        mark_cells32<BLEND>(cells, cw, cx0, ncell, cy, (uint32_t)r + 1u, alpha, bl, ctr, lane);
    }
}

// One thread resolves two horizontally adjacent cells = 4 pixels on 2 image rows. A cell holds 0 (empty) or, + 1,
// the index of the last record that covers it (BY_PATH false: marked by k_fill_cells) or the highest path that
// covers it (BY_PATH true: marked by k_spans<true, .>); its colour is that record's / that path's.
// SLPR_FLAG_BLEND: the colour of one cell = white, or its top-most opaque record / path, with the translucent ones above
// it composited in ascending priority (blend_over). The list is unordered (atomicExch pushes); it is walked once per
// entry composited, each walk picking the smallest priority not yet done — lists are a handful of nodes long.
template <bool BY_PATH>
__device__ __forceinline__ uint32_t composite_cell(uint32_t top, uint32_t head, const uint2 *__restrict__ nodes,
                                                   const int4 *__restrict__ records, const uint32_t *__restrict__ fill_info) {
    uint32_t dst = top ? (BY_PATH ? fill_info[top - 1] : (uint32_t)records[top - 1].z) : 0xFFFFFFFFu;
    uint32_t last = top;
    while (head) {
        uint32_t best = 0xFFFFFFFFu;
        for (uint32_t n = head; n; ) {
            const uint2 nd = nodes[n - 1];
            if (nd.x > last && nd.x < best) best = nd.x;
            n = nd.y;
        }
        if (best == 0xFFFFFFFFu) break;
        dst = blend_over(dst, BY_PATH ? fill_info[best - 1] : (uint32_t)records[best - 1].z);
        last = best;
    }
    return dst;
}

template <bool BY_PATH, bool BLEND>
__global__ void __launch_bounds__(256) k_resolve(const FrameParams *__restrict__ P, const int4 *__restrict__ records,
                                                 const uint32_t *__restrict__ fill_info,
                                                 uint32_t *__restrict__ cells, int cw, uint8_t *__restrict__ fb,
                                                 size_t stride_bytes, BlendList bl) {
    const int width = P->width, height = P->height;
    const int cy0 = P->band_y0 >> 1, cy1 = (P->band_y1 + 1) >> 1;
    const int pairs = (cw + 1) >> 1;
    const long long total = (long long)pairs * (cy1 - cy0);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int cy = cy0 + (int)(t / pairs);
        const int cx = (int)(t % pairs) * 2;
        uint32_t *cp = cells + (size_t)cy * cw + cx;
        uint32_t col[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            col[k] = 0xFFFFFFFFu;  // clear colour (1,1,1,1), SR.cpp:622
            if (cx + k < cw) {
                const uint32_t v = cp[k];
                if (BLEND) {
                    uint32_t *hp = bl.heads + (size_t)cy * cw + cx + k;
                    const uint32_t head = *hp;
                    if (head) *hp = 0;
                    if (v) cp[k] = 0;
                    col[k] = composite_cell<BY_PATH>(v, head, bl.nodes, records, fill_info);
                } else if (v) { col[k] = BY_PATH ? fill_info[v - 1] : (uint32_t)records[v - 1].z; cp[k] = 0; }  // colour bytes R,G,B,A = fill_info (VERT:8-10)
            }
        }
        const int px = cx * 2;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int row = cy * 2 + dy;  // scanline row
            if (row >= height) continue;
            uint8_t *dst = fb + (size_t)(height - 1 - row) * stride_bytes + (size_t)px * 4;  // VERT:44 y flip
            if (px + 3 < width && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(col[0], col[0], col[1], col[1]);
            } else {
                uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
                for (int k = 0; k < 4; ++k)
                    if (px + k < width) d32[k] = col[k >> 1];
            }
        }
    }
}

// SLPR_FLAG_AA4 (beyond the reference, which has no antialiasing: README.md:5 "comb-like multisampling (TODO)"):
// the pipeline ran at four times the frame's size, so every pixel of the frame is 2 x 2 coverage cells = four samples on
// a regular grid, each holding the top-most path that covers it (ordered, opaque compositing per sample); the pixel is
// their box-filtered average, (sum + 2) >> 2 per channel. One thread per pixel; the cells are re-zeroed.
template <bool BY_PATH, bool BLEND>
__global__ void __launch_bounds__(256) k_resolve_aa4(const FrameParams *__restrict__ P, const int4 *__restrict__ records,
                                                     const uint32_t *__restrict__ fill_info, uint32_t *__restrict__ cells, int cw,
                                                     uint8_t *__restrict__ fb, size_t stride_bytes, int out_w, int out_h, BlendList bl) {
    const int y_lo = P->band_y0 >> 2, y_hi = (P->band_y1 + 3) >> 2;  // output scanline rows of the band
    const long long total = (long long)out_w * (y_hi - y_lo);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int y = y_lo + (int)(t / out_w), x = (int)(t % out_w);
        if (y >= out_h) continue;
        uint32_t sum[4] = {2u, 2u, 2u, 2u};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            uint2 *cp = reinterpret_cast<uint2 *>(cells + (size_t)(2 * y + dy) * cw + 2 * x);  // cw is even (4 x width / 2)
            const uint2 v = *cp;
            if (v.x | v.y) *cp = make_uint2(0u, 0u);
            uint2 hd = make_uint2(0u, 0u);
            if (BLEND) {
                uint2 *hp = reinterpret_cast<uint2 *>(bl.heads + (size_t)(2 * y + dy) * cw + 2 * x);
                hd = *hp;
                if (hd.x | hd.y) *hp = make_uint2(0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t id = k ? v.y : v.x;
                const uint32_t col = BLEND ? composite_cell<BY_PATH>(id, k ? hd.y : hd.x, bl.nodes, records, fill_info)
                                           : id ? (BY_PATH ? fill_info[id - 1] : (uint32_t)records[id - 1].z) : 0xFFFFFFFFu;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) sum[ch] += (col >> (8 * ch)) & 0xFFu;
            }
        }
        const uint32_t out = (sum[0] >> 2) | ((sum[1] >> 2) << 8) | ((sum[2] >> 2) << 16) | ((sum[3] >> 2) << 24);
        *reinterpret_cast<uint32_t *>(fb + (size_t)(out_h - 1 - y) * stride_bytes + (size_t)x * 4) = out;
    }
}

}  // namespace slpr
