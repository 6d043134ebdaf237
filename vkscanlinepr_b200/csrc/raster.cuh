// raster.cuh — stage 5: draw records -> RGBA8 framebuffer. Replaces scanlinepr.vert/.frag and the
// fixed-function LINE_LIST raster (scanlinepr.vert:19-46; SR.cpp:611-656,883-895): every record
// covers pixels x in [X, X+width) on scanline rows {Y, Y+1}; writes are opaque and the later
// record wins; the clear colour is white (SR.cpp:622); image row = H-1-scanline row (VERT:44).
//
// All record coordinates are even, so coverage is resolved on the 2x2-pixel cell grid:
//   k_fill_cells    atomicMax(cell, record index + 1)  — deterministic "later record wins" (big frames: the span
//                   kernel does this itself, spans.cuh)
//   k_resolve       cell -> colour of that record (or white), 2x2 pixels, 128-bit stores; also
//                   re-zeroes the cell grids for the next frame.
// The cell grids of a 4K frame are 8 + 2 MB and stay L2-resident between the two kernels.
#pragma once
#include "common.cuh"

namespace slpr {

#ifndef SLPR_FILL_NARROW
#define SLPR_FILL_NARROW 8 /* records up to this many cells are filled by their own thread (measured: 4 -> 0.162, 8 -> 0.147, 16 -> 0.199 ms) */
#endif

// Coverage is kept on two grids: `cells` (one word per 2x2-pixel cell) and `cells4` (one word per aligned group
// of four cells in a row). A record marks whole groups it covers on the coarse grid and only the cells of its
// partly covered end groups on the fine one; a cell's owner is the larger of its two words. The fill is bound by
// the rate of L2 atomics (~1 per clock and SM), and spans average ten cells: this halves their atomics.
struct CellGrids {
    uint32_t *cells;   // [ch][cw]
    uint32_t *cells4;  // [ch][cw4], cw4 = (cw + 3) / 4
    int cw, cw4;
};

// cells [cx0, cx0 + ncell) of cell row cy, by one thread
__device__ __forceinline__ void mark_cells_thread(const CellGrids &g, int cy, int cx0, int ncell, uint32_t prio) {
    const int end = cx0 + ncell;
    const int g0 = (cx0 + 3) >> 2, g1 = end >> 2;  // whole groups [g0, g1)
    uint32_t *row = g.cells + (size_t)cy * g.cw;
    if (g0 >= g1) {
        for (int c = cx0; c < end; ++c) atomicMax(row + c, prio);
        return;
    }
    for (int c = cx0; c < (g0 << 2); ++c) atomicMax(row + c, prio);
    uint32_t *row4 = g.cells4 + (size_t)cy * g.cw4;
    for (int q = g0; q < g1; ++q) atomicMax(row4 + q, prio);
    for (int c = g1 << 2; c < end; ++c) atomicMax(row + c, prio);
}

// the same by a whole warp (wide spans): lanes over the groups, then over the (at most six) end cells
__device__ __forceinline__ void mark_cells_warp(const CellGrids &g, int cy, int cx0, int ncell, uint32_t prio, int lane) {
    const int end = cx0 + ncell;
    const int g0 = (cx0 + 3) >> 2, g1 = end >> 2;
    uint32_t *row = g.cells + (size_t)cy * g.cw;
    if (g0 >= g1) {
        for (int c = cx0 + lane; c < end; c += 32) atomicMax(row + c, prio);
        return;
    }
    uint32_t *row4 = g.cells4 + (size_t)cy * g.cw4;
    for (int q = g0 + lane; q < g1; q += 32) atomicMax(row4 + q, prio);
    const int head = (g0 << 2) - cx0, tail = end - (g1 << 2);  // 0..3 each
    if (lane < head) atomicMax(row + cx0 + lane, prio);
    else if (lane - head < tail) atomicMax(row + (g1 << 2) + (lane - head), prio);
}

// A warp's 32 records (ncell = 0: nothing to mark): narrow ones by their own lane, wide spans by the whole warp.
__device__ __forceinline__ void mark_records_warp(const CellGrids &g, int cy, int cx0, int ncell, uint32_t prio, int lane) {
    if (ncell > 0 && ncell <= SLPR_FILL_NARROW) mark_cells_thread(g, cy, cx0, ncell, prio);
    uint32_t wide = __ballot_sync(0xFFFFFFFFu, ncell > SLPR_FILL_NARROW);
    while (wide) {
        const int src = __ffs(wide) - 1;
        wide &= wide - 1;
        const int s_cx0 = __shfl_sync(0xFFFFFFFFu, cx0, src);
        const int s_n = __shfl_sync(0xFFFFFFFFu, ncell, src);
        const int s_cy = __shfl_sync(0xFFFFFFFFu, cy, src);
        const uint32_t s_prio = __shfl_sync(0xFFFFFFFFu, prio, src);
        mark_cells_warp(g, s_cy, s_cx0, s_n, s_prio, lane);
    }
}

__global__ void __launch_bounds__(256) k_fill_cells(const FrameParams *__restrict__ P,
                                                    const FrameCounters *__restrict__ ctr, int capacity,
                                                    const int4 *__restrict__ records, CellGrids g) {
    if (ctr->n_fragments > capacity) return;
    const int nrec = ctr->n_records;
    const int nround = (nrec + 31) & ~31;
    const int height = P->height;
    const int lane = (int)lane_id();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nround; r += gridDim.x * blockDim.x) {
        int cx0 = 0, ncell = 0, cy = 0;
        if (r < nrec) {
            const int4 rec = records[r];
            const int X = rec.x & 0xFFFF, Y = rec.x >> 16;  // VERT:27
            if (Y >= 0 && Y < height) {
                cx0 = X >> 1;
                ncell = min((X + rec.y) >> 1, g.cw) - cx0;
                cy = Y >> 1;
            }
        }
        mark_records_warp(g, cy, cx0, ncell, (uint32_t)r + 1u, lane);
    }
}

// One thread resolves one group of four horizontally adjacent cells = 8 pixels on 2 image rows, and clears
// the group's words on both grids for the next frame.
__global__ void __launch_bounds__(256) k_resolve(const FrameParams *__restrict__ P, const int4 *__restrict__ records,
                                                 CellGrids g, uint8_t *__restrict__ fb, size_t stride_bytes) {
    const int width = P->width, height = P->height;
    const int cy0 = P->band_y0 >> 1, cy1 = (P->band_y1 + 1) >> 1;
    const long long total = (long long)g.cw4 * (cy1 - cy0);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int cy = cy0 + (int)(t / g.cw4);
        const int q = (int)(t % g.cw4), cx = q * 4;
        uint32_t *cp = g.cells + (size_t)cy * g.cw + cx;
        uint32_t *cp4 = g.cells4 + (size_t)cy * g.cw4 + q;
        const uint32_t v4 = *cp4;
        if (v4) *cp4 = 0;
        uint32_t col[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            col[k] = 0xFFFFFFFFu;  // clear colour (1,1,1,1), SR.cpp:622
            if (cx + k < g.cw) {
                uint32_t v = cp[k];
                if (v) cp[k] = 0;
                v = max(v, v4);
                if (v) col[k] = (uint32_t)records[v - 1].z;  // colour bytes R,G,B,A = fill_info (VERT:8-10)
            }
        }
        const int px = cx * 2;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int row = cy * 2 + dy;  // scanline row
            if (row >= height) continue;
            uint8_t *dst = fb + (size_t)(height - 1 - row) * stride_bytes + (size_t)px * 4;  // VERT:44 y flip
            if (px + 7 < width && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                reinterpret_cast<uint4 *>(dst)[0] = make_uint4(col[0], col[0], col[1], col[1]);
                reinterpret_cast<uint4 *>(dst)[1] = make_uint4(col[2], col[2], col[3], col[3]);
            } else {
                uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
                for (int k = 0; k < 8; ++k)
                    if (px + k < width) d32[k] = col[k >> 1];
            }
        }
    }
}

}  // namespace slpr
