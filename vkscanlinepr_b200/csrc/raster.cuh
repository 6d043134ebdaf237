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
#define SLPR_FILL_NARROW 8 /* records up to this many cells are filled by their own thread (measured: 4 -> 0.162, 8 -> 0.147, 16 -> 0.199 ms) */
#endif

__global__ void __launch_bounds__(256) k_fill_cells(const FrameParams *__restrict__ P,
                                                    const FrameCounters *__restrict__ ctr, int capacity,
                                                    const int4 *__restrict__ records, uint32_t *__restrict__ cells,
                                                    int cw) {
    if (frame_void(ctr, capacity)) return;
    const int nrec = ctr->n_records;
    const int nround = (nrec + 31) & ~31;
    const int height = P->height;
    const uint32_t lane = lane_id();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nround; r += gridDim.x * blockDim.x) {
        int cx0 = 0, ncell = 0, cy = 0;
        if (r < nrec) {
            const int4 rec = records[r];
            const int X = rec.x & 0xFFFF, Y = rec.x >> 16;  // VERT:27
            if (Y >= 0 && Y < height) {
                cx0 = X >> 1;
                ncell = min((X + rec.y) >> 1, cw) - cx0;
                cy = Y >> 1;
            }
        }
        const uint32_t prio = (uint32_t)r + 1u;
        if (ncell > 0 && ncell <= SLPR_FILL_NARROW) {
            uint32_t *row = cells + (size_t)cy * cw + cx0;
            for (int c = 0; c < ncell; ++c) atomicMax(row + c, prio);
        }
        // wide spans: the whole warp fills them, one after the other
        uint32_t wide = __ballot_sync(0xFFFFFFFFu, ncell > SLPR_FILL_NARROW);
        while (wide) {
            const int src = __ffs(wide) - 1;
            wide &= wide - 1;
            const int s_cx0 = __shfl_sync(0xFFFFFFFFu, cx0, src);
            const int s_n = __shfl_sync(0xFFFFFFFFu, ncell, src);
            const int s_cy = __shfl_sync(0xFFFFFFFFu, cy, src);
            const uint32_t s_prio = __shfl_sync(0xFFFFFFFFu, prio, src);
            uint32_t *row = cells + (size_t)s_cy * cw + s_cx0;
            for (int c = (int)lane; c < s_n; c += 32) atomicMax(row + c, s_prio);
        }
    }
}

// One thread resolves two horizontally adjacent cells = 4 pixels on 2 image rows. A cell holds 0 (empty) or, + 1,
// the index of the last record that covers it (BY_PATH false: marked by k_fill_cells) or the highest path that
// covers it (BY_PATH true: marked by k_spans<true, .>); its colour is that record's / that path's.
template <bool BY_PATH>
__global__ void __launch_bounds__(256) k_resolve(const FrameParams *__restrict__ P, const int4 *__restrict__ records,
                                                 const uint32_t *__restrict__ fill_info,
                                                 uint32_t *__restrict__ cells, int cw, uint8_t *__restrict__ fb,
                                                 size_t stride_bytes) {
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
                if (v) { col[k] = BY_PATH ? fill_info[v - 1] : (uint32_t)records[v - 1].z; cp[k] = 0; }  // colour bytes R,G,B,A = fill_info (VERT:8-10)
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
template <bool BY_PATH>
__global__ void __launch_bounds__(256) k_resolve_aa4(const FrameParams *__restrict__ P, const int4 *__restrict__ records,
                                                     const uint32_t *__restrict__ fill_info, uint32_t *__restrict__ cells, int cw,
                                                     uint8_t *__restrict__ fb, size_t stride_bytes, int out_w, int out_h) {
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
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t id = k ? v.y : v.x;
                const uint32_t col = id ? (BY_PATH ? fill_info[id - 1] : (uint32_t)records[id - 1].z) : 0xFFFFFFFFu;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) sum[ch] += (col >> (8 * ch)) & 0xFFu;
            }
        }
        const uint32_t out = (sum[0] >> 2) | ((sum[1] >> 2) << 8) | ((sum[2] >> 2) << 16) | ((sum[3] >> 2) << 24);
        *reinterpret_cast<uint32_t *>(fb + (size_t)(out_h - 1 - y) * stride_bytes + (size_t)x * 4) = out;
    }
}

}  // namespace slpr
