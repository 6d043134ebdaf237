// bands.cuh — exact row bands across GPUs.
//
// The reference's winding scan is ONE unsegmented prefix sum over all fragments in (path, row, x)
// order (SR.cpp:479-506): a path whose winding deltas do not cancel (the MI0:340 slip
// leaves such residues) shifts the winding of everything after it. A GPU that renders only a band
// of rows therefore needs, for each of its fragments, the deltas of every fragment of the other
// bands that sorts before it. Order inside a path is: rows y = 2, 4, ... (bottom to top), then the
// invalid-key fragments (outside the frame), then row y = 0 — so with bands ordered by rows,
//   winding(f) = local scan(f) + corr[path(f)]
//   corrN[p] = sum_{q<p} (total[q] - mine[q]) + sum_{bands below mine} normal[p]      (rows y >= 2)
//   corrZ[p] = sum_{q<p} (total[q] - mine[q]) + sum_{other bands} (normal + invalid)[p]  (row y = 0)
// where normal / invalid / zero-row are each band's per-path delta sums. The exchange step is an
// all-gather of those 3 * n_paths integers per band (NCCL over NVLink, done by the host between
// slpr_render_band_begin and slpr_render_band_end); everything else stays on the GPU.
//   k_band_sums   one warp per path over its contiguous fragment range -> sums[3][P]
//   k_band_other  d[q] = sum over the OTHER bands of (normal + invalid + zero-row)[q]
//   (scan of d: k_lookback_scan<ScanI32Op>)
//   k_band_corr   corrN, corrZ from the scanned d and the gathered sums
// Every out-of-frame fragment must be owned by exactly one band: make_fragment (geom.cuh) gives rows
// below the frame to the band that starts at 0 and rows above it to the band that ends at H.
#pragma once
#include "geom.cuh"

namespace slpr {

__global__ void __launch_bounds__(256) k_band_sums(const int *__restrict__ seg, uint32_t n_paths,
                                                   const uint64_t *__restrict__ key, const uint32_t *__restrict__ val,
                                                   const FrameCounters *__restrict__ ctr, int capacity, KeyLayout L,
                                                   int *__restrict__ sums) {
    if (ctr->n_fragments > capacity) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t ymask = (1ull << L.bits_y) - 1;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_paths; p += warps) {
        const int b = seg[p], e = seg[p + 1];
        int a = 0, inv = 0, z = 0;
        for (int i = b + lane; i < e; i += 32) {
            const int d = (int)(val[i] >> 30) - 1;
            const uint32_t yk = (uint32_t)((key[i] >> L.bits_x) & ymask);
            if (yk == (uint32_t)L.ny) z += d;
            else if (yk == (uint32_t)(L.ny - 1)) inv += d;
            else a += d;
        }
        a = __reduce_add_sync(0xFFFFFFFFu, a);
        inv = __reduce_add_sync(0xFFFFFFFFu, inv);
        z = __reduce_add_sync(0xFFFFFFFFu, z);
        if (lane == 0) { sums[p] = a; sums[n_paths + p] = inv; sums[2 * (size_t)n_paths + p] = z; }
    }
}

// gathered: [n_ranks][3][P]. One pass over the gathered sums: d[p] (scanned next) and the two per-path terms
// of the corrections, so that k_band_corr only adds the scan.
__global__ void __launch_bounds__(256) k_band_other(const int *__restrict__ gathered, uint32_t n_paths, int n_ranks, int rank,
                                                    int *__restrict__ d, int *__restrict__ corr) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_paths; p += gridDim.x * blockDim.x) {
        int all = 0, below = 0, others = 0;
        for (int r = 0; r < n_ranks; ++r) {
            if (r == rank) continue;
            const int *g = gathered + (size_t)r * 3 * n_paths;
            const int a = g[p], inv = g[n_paths + p], z = g[2 * (size_t)n_paths + p];
            all += a + inv + z;
            others += a + inv;
            if (r < rank) below += a;
        }
        d[p] = all;
        corr[p] = below;
        corr[n_paths + p] = others;
    }
}

// corr: [2][P] = corrN | corrZ; e = exclusive scan of d
__global__ void __launch_bounds__(256) k_band_corr(const int *__restrict__ e, uint32_t n_paths, int *__restrict__ corr) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_paths; p += gridDim.x * blockDim.x) {
        const int v = e[p];
        corr[p] += v;
        corr[n_paths + p] += v;
    }
}

}  // namespace slpr
