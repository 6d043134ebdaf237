// segsort.cuh — segmented sort fast path for scenes whose paths are all small.
//
// The reference sorts each path's fragments separately (naive_seg_sort_pairs.comp:26-97, one
// workgroup per path, segments from gen_fragment.comp:226-244). Fragments are generated in path
// order, so a path's fragments already sit in one contiguous range [seg[p], seg[p+1]); the general
// path (radix.cuh) ignores that and runs a global LSD sort on (path | row | x) — 4-6 full passes
// over the 12-byte pairs. When every path is small (a glyph, a blob: tens to hundreds of
// fragments) it is much cheaper to sort each range on chip, reading and writing every pair ONCE:
//   k_segsort_warp   one warp per path, up to 512 fragments: bitonic network on one word per fragment
//                    (row|x above the fragment's index, so a single unsigned compare gives the
//                    reference's (key, index) order) held in registers, compare-exchange by warp
//                    shuffles. 32-bit words holding row|x relative to the path's minimum are used
//                    whenever they fit (any path covering a small part of the frame), else 64-bit;
//   k_segsort_block  paths of 513..4096 fragments, and shorter ones that span too much of the frame
//                    for the warp kernel's 32-bit words (queued by the warp kernel): bitonic network in
//                    shared memory, one block per path.
// A path with more than 4096 fragments raises FrameCounters::sort_fallback; the host then switches
// the scene to the onesweep radix sort and renders the frame again. Output format is identical to
// the radix path (compact 64-bit key, 32-bit value), so everything downstream is unchanged.
#pragma once
#include "geom.cuh"
#include "bands.cuh"

namespace slpr {

constexpr int SEG_WARP_MAX = 512;    // largest path sorted by one warp (16 words per lane)
constexpr int SEG_BLOCK_MAX = 4096;  // largest path sorted by one block
constexpr int SEG_BLOCK_THREADS = 512;

__device__ __forceinline__ uint64_t seg_pack(uint64_t key, uint32_t val, int yx_bits) {
    const uint64_t yx = key & ((1ull << yx_bits) - 1);
    return (yx << 32) | ((uint64_t)(val & VAL_INDEX_MASK) << 3) | (uint64_t)(val >> 29);
}
__device__ __forceinline__ void seg_unpack(uint64_t w, uint64_t path_bits, uint64_t &key, uint32_t &val) {
    key = path_bits | (w >> 32);
    val = (uint32_t)((w >> 3) & VAL_INDEX_MASK) | ((uint32_t)(w & 7u) << 29);
}

// Bitonic sort of 32*R words held as e[r] = element r*32 + lane, ascending. Up to R = 4 the network
// is fully unrolled (fastest: 0.155 vs 0.32 ms on the 4K bench scene). From R = 8 on the stage loops
// stay rolled and only the R registers of one stage are unrolled: the unrolled R = 8 network is 44 KB
// of code and the kernel then stalls on instruction fetch (measured at 16K: 66 % no_inst, 4.9 -> 2.2 ms).
template <int R, typename W>
__device__ __forceinline__ void warp_bitonic(W (&e)[R], int lane) {
    if constexpr (R <= 4) {
#pragma unroll
        for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= 32) {  // partner lives in another register of the same lane
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int rp = r ^ (j >> 5);
                        if (rp > r) {
                            const bool asc = (((r * 32 + lane) & k) == 0);
                            const W a = e[r], b = e[rp];
                            const W lo = a < b ? a : b, hi = a < b ? b : a;
                            e[r] = asc ? lo : hi;
                            e[rp] = asc ? hi : lo;
                        }
                    }
                } else {  // partner lives in lane ^ j
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const W o = __shfl_xor_sync(0xFFFFFFFFu, e[r], j);
                        const bool asc = (((r * 32 + lane) & k) == 0);
                        const bool lower = (lane & j) == 0;
                        const W mn = e[r] < o ? e[r] : o, mx = e[r] < o ? o : e[r];
                        e[r] = (asc == lower) ? mn : mx;
                    }
                }
            }
        }
    } else {
#pragma unroll 1
        for (int k = 2; k <= 32 * R; k <<= 1) {
            // partner in another register of the same lane (j = 32 * m); register indices are compile-time
#pragma unroll
            for (int m = R / 2; m >= 1; m >>= 1) {
                if (32 * m < k) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if ((r & m) == 0) {
                            const bool asc = (((r * 32) & k) == 0);  // lane bits are below 32 <= j < k
                            const W a = e[r], b = e[r | m];
                            const W lo = a < b ? a : b, hi = a < b ? b : a;
                            e[r] = asc ? lo : hi;
                            e[r | m] = asc ? hi : lo;
                        }
                    }
                }
            }
            // partner in lane ^ j: the five shuffle stages of a phase are unrolled (compile-time lane masks), the
            // phase loop is not
#pragma unroll
            for (int j = 16; j > 0; j >>= 1) {
                if (j >= k) continue;  // phase k starts at j = k / 2
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const W o = __shfl_xor_sync(0xFFFFFFFFu, e[r], j);
                    const bool asc = (((r * 32 + lane) & k) == 0);
                    const W mn = e[r] < o ? e[r] : o, mx = e[r] < o ? o : e[r];
                    e[r] = (asc == lower) ? mn : mx;
                }
            }
        }
    }
}

// 96 words: paths of 65..96 fragments are the largest class of the 4K bench scene (40 % of its paths, half of its
// fragments) and the 128-word network above does a third more work than they need. A bitonic network wants a power
// of two, but only in the number of BLOCKS: with a sorted block of three words per lane, the 15 compare-exchange
// stages of the 32-lane network become merge-splits — a lane takes its partner's block reversed (three shuffles),
// keeps the element-wise minima or maxima (the lower or upper half of the two blocks' union, a bitonic triple) and
// sorts its three words again (min3 / max3 and the middle by exclusive-or). 13 instructions per stage against the
// 25 x 12 + 3 register stages of R = 4; the result is blocked (lane L: ranks 3L..3L+2) and goes through the warp's
// shared-memory slice to the striped order (rank r * 32 + lane) everything around expects.
__device__ __forceinline__ void sort3(uint32_t &a, uint32_t &b, uint32_t &c) {
    const uint32_t mn = min(min(a, b), c), mx = max(max(a, b), c);
    const uint32_t md = a ^ b ^ c ^ mn ^ mx;
    a = mn; b = md; c = mx;
}
__device__ __forceinline__ void warp_mergesplit96(uint32_t (&e)[3], int lane, uint32_t *__restrict__ s_tmp) {
    sort3(e[0], e[1], e[2]);
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
            const uint32_t b0 = __shfl_xor_sync(0xFFFFFFFFu, e[2], j), b1 = __shfl_xor_sync(0xFFFFFFFFu, e[1], j),
                           b2 = __shfl_xor_sync(0xFFFFFFFFu, e[0], j);
            e[0] = keep_min ? min(e[0], b0) : max(e[0], b0);
            e[1] = keep_min ? min(e[1], b1) : max(e[1], b1);
            e[2] = keep_min ? min(e[2], b2) : max(e[2], b2);
            sort3(e[0], e[1], e[2]);
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) s_tmp[3 * lane + r] = e[r];  // (stride 3: no bank conflicts)
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; ++r) e[r] = s_tmp[r * 32 + lane];
}

template <int R>
__device__ __forceinline__ void warp_sort_segment(const uint64_t *__restrict__ key_in, const uint32_t *__restrict__ val_in,
                                                  uint64_t *__restrict__ key_out, uint32_t *__restrict__ val_out, int b, int n,
                                                  int yx_bits, int lane) {
    uint64_t e[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = r * 32 + lane;
        e[r] = (j < n) ? seg_pack(key_in[b + j], val_in[b + j], yx_bits) : ~0ull;
    }
    const uint64_t path_bits = key_in[b] & ~((1ull << yx_bits) - 1);  // same for the whole segment
    warp_bitonic<R, uint64_t>(e, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = r * 32 + lane;
        if (j < n) {
            uint64_t k; uint32_t v;
            seg_unpack(e[r], path_bits, k, v);
            key_out[b + j] = k;
            val_out[b + j] = v;
        }
    }
}

template <int R>
struct SegIdxBits {
    static constexpr int value = (R == 1) ? 5 : (R == 2) ? 6 : (R == 3 || R == 4) ? 7 : (R == 8) ? 8 : 9;
};

// Same, on 32-bit words: (row|x relative to the path's smallest row|x) above the fragment's position
// inside the path — half the shuffle traffic of the 64-bit network. A path covers a small part of the
// frame, so the relative row|x almost always fits in 32 - log2(32 R) bits at any resolution; when it
// does not (warp-uniform test) the 64-bit network sorts the segment instead. A fragment's index is its
// position, so (row|x, position) is the reference's (key, index) order. The value words are loaded together
// with the keys (one exposure to HBM latency per path instead of two: the kernel waits on loads half of its
// time) and parked in the warp's slice of shared memory, from where the sorted positions pick them.
// Two special row values sort after every real row of a path: the invalid key (row rank ny - 1: fragments outside the
// frame, and in band mode the fragments of other bands) and row 0 (rank ny). Left as they are they stretch the path's
// key range over the whole frame height and the path would need 64-bit words (twice the shuffles, and in band mode
// EVERY path that straddles a band edge has them). So, when a path has any (warp-uniform test), they are folded in
// right above its largest real key: rel = (mx - mn + 1) + (special rank << bits_x) + x — same order, small range —
// and unfolded on the way out.
struct SegGeo {
    int bits_x, bits_y, ny;
};

// `sums` (band mode): the path's three winding sums are formed from the values while they pass through (bands.cuh).
template <int R, bool BAND>
__device__ __forceinline__ bool warp_sort_segment32(const uint64_t *__restrict__ key_in, const uint32_t *__restrict__ val_in,
                                                    uint64_t *__restrict__ key_out, uint32_t *__restrict__ val_out, int b, int n,
                                                    int yx_bits, int lane, uint32_t *__restrict__ s_val, SegGeo geo, int (&sums)[3]) {
    constexpr int IB = SegIdxBits<R>::value;
    static_assert(R == 1 || R == 2 || R == 3 || R == 4 || R == 8 || R == 16, "R must be 1, 2, 3, 4, 8 or 16");
    const uint64_t mask = (1ull << yx_bits) - 1;
    const uint32_t special_from = (uint32_t)(geo.ny - 1) << geo.bits_x;  // keys at or above this are invalid / row 0
    uint32_t e[R];
    uint64_t path_bits = 0;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    int a = 0, inv = 0, z = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = r * 32 + lane;
        e[r] = 0u;
        if (j < n) {
            const uint64_t k = key_in[b + j];
            const uint32_t v = val_in[b + j];
            s_val[j] = v;
            path_bits = k & ~mask;
            e[r] = (uint32_t)(k & mask);
            mn = min(mn, e[r]);
            mx = max(mx, e[r]);
            if (BAND) {
                const int d = (int)(v >> 30) - 1;
                if (e[r] < special_from) a += d;
                else if ((e[r] >> geo.bits_x) == (uint32_t)geo.ny) z += d;
                else inv += d;
            }
        }
    }
    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);  // (also orders the s_val stores before the reads below)
    const bool special = BAND && mx >= special_from;  // warp-uniform. (Full frames: only paths at the frame border have such keys; they keep the 64-bit fall-back and the common path stays as lean as it was.)
    uint32_t span = mx - mn + 1u;             // real keys map to [0, span)
    if (special) {  // the real rows' range, and the special keys right above it
        mn = 0xFFFFFFFFu; mx = 0u;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r * 32 + lane < n && e[r] < special_from) { mn = min(mn, e[r]); mx = max(mx, e[r]); }
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if (mn > mx) { mn = 0; mx = 0; }  // no real row at all
        span = mx - mn + 1u;
    }
    const uint32_t range = special ? span + (2u << geo.bits_x) : span;
    if (((range - 1u) >> (32 - IB)) != 0u) return false;  // too much of the frame for a 32-bit word: the caller's fall-back
    if (BAND) {
        sums[0] = __reduce_add_sync(0xFFFFFFFFu, a);
        sums[1] = __reduce_add_sync(0xFFFFFFFFu, inv);
        sums[2] = __reduce_add_sync(0xFFFFFFFFu, z);
    }
    path_bits = __shfl_sync(0xFFFFFFFFu, path_bits, 0);  // lane 0 always holds fragment 0
    if (!special) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = r * 32 + lane;
            e[r] = (j < n) ? (((e[r] - mn) << IB) | (uint32_t)j) : 0xFFFFFFFFu;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = r * 32 + lane;
            const uint32_t rel = (e[r] >= special_from) ? span + (e[r] - special_from) : e[r] - mn;
            e[r] = (j < n) ? ((rel << IB) | (uint32_t)j) : 0xFFFFFFFFu;
        }
    }
    if constexpr (R == 3) warp_mergesplit96(e, lane, s_val + 128);  // (the path's values take the first 96 words of the slice)
    else warp_bitonic<R, uint32_t>(e, lane);
    __syncwarp();
    if (!special) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = r * 32 + lane;
            if (j < n) {
                key_out[b + j] = path_bits | (uint64_t)(mn + (e[r] >> IB));
                val_out[b + j] = s_val[e[r] & ((1u << IB) - 1)];
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = r * 32 + lane;
            if (j < n) {
                const uint32_t rel = e[r] >> IB;
                const uint32_t kk = (rel >= span) ? special_from + (rel - span) : mn + rel;
                key_out[b + j] = path_bits | (uint64_t)kk;
                val_out[b + j] = s_val[e[r] & ((1u << IB) - 1)];
            }
        }
    }
    __syncwarp();  // the slice is rewritten by the warp's next path
    return true;
}

// Sort segment table (gen_fragment.comp:226-244): seg[p] = first record of the first curve whose path is >= p,
// seg[n_paths] = nf — read off the scanned curve offsets through the scene's static table of every path's first
// curve (slpr_load_scene) — and, in the same pass, the path-size statistics for the host's choice between the two
// sorts (slpr.cu: segmented_sort_pays).
// Band mode: also the fragment ranges of the listed live paths, in list order (k_segsort_warp then reads its paths'
// bounds with one coalesced load instead of two dependent scattered ones per path).
__global__ void __launch_bounds__(256) k_path_segments(const uint32_t *__restrict__ path_first_curve, uint32_t n_paths,
                                                       const int *__restrict__ offsets, int *__restrict__ seg,
                                                       FrameCounters *__restrict__ ctr, int capacity,
                                                       const uint32_t *__restrict__ live_paths, int2 *__restrict__ live_range) {
    if (ctr->n_fragments > capacity) return;
    if (live_paths) {
        const uint32_t n_live = (uint32_t)ctr->n_live_paths;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_live; i += gridDim.x * blockDim.x) {
            const uint32_t p = live_paths[i];
            live_range[i] = make_int2(offsets[path_first_curve[p]], offsets[path_first_curve[p + 1]]);
        }
    }
    int mid = 0, big = 0, huge = 0;
    const uint32_t n_round = (n_paths + 1u + 31u) & ~31u;  // whole warps stay in the loop for the reductions
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += gridDim.x * blockDim.x) {
        if (p > n_paths) continue;
        const int s0 = offsets[path_first_curve[p]];
        seg[p] = s0;
        if (p < n_paths) {
            const int n = offsets[path_first_curve[p + 1]] - s0;
            mid += (n > 128 && n <= SEG_WARP_MAX);
            big += (n > SEG_WARP_MAX && n <= SEG_BLOCK_MAX);
            huge += (n > SEG_BLOCK_MAX);
        }
    }
    mid = __reduce_add_sync(0xFFFFFFFFu, mid);
    big = __reduce_add_sync(0xFFFFFFFFu, big);
    huge = __reduce_add_sync(0xFFFFFFFFu, huge);
    if ((threadIdx.x & 31) == 0) {
        if (mid) atomicAdd(&ctr->stat_mid, mid);
        if (big) atomicAdd(&ctr->stat_big, big);
        if (huge) atomicAdd(&ctr->stat_huge, huge);
    }
}

constexpr int SEG_CHUNK = 8;  // consecutive paths per warp trip (one contiguous run of fragments)

// Band mode (exact bands, device-side exchange): the kernel visits only the paths that can reach the band (the list the
// band front end left) and, while it has a path's fragments in hand, forms the three winding sums the exchange needs
// (normal rows | outside the frame | row 0; bands.cuh) — a separate pass over all paths cost as much as a third of
// the sort itself (profiles/README.md, round 2).
struct SegBand {
    int *ticket;                 // band mode: chunks are handed out dynamically (zeroed per frame)
    const uint32_t *live_paths;  // nullptr: every path (full frame)
    const int2 *live_range;      // [n_live_paths] fragment range of each listed path
    BandEntry *sums;             // nullptr: no sums wanted
    SegGeo geo;
};

#ifndef SLPR_SEG_R3
#define SLPR_SEG_R3 1 /* paths of 65..96 fragments on the 96-word merge-split network instead of the 128-word bitonic one */
#endif
#ifndef SLPR_SEG_BAND_BLOCKS
#define SLPR_SEG_BAND_BLOCKS 4
#endif
#ifndef SLPR_SEG_FUSE_SUMS
#define SLPR_SEG_FUSE_SUMS 1 /* 0 (experiments): band sums from k_band_sums_sparse over the live paths instead */
#endif
template <bool BAND>
__global__ void __launch_bounds__(256, BAND ? SLPR_SEG_BAND_BLOCKS : 4) k_segsort_warp(const int *__restrict__ seg, uint32_t n_paths,
                                                      const uint64_t *__restrict__ key_in, const uint32_t *__restrict__ val_in,
                                                      uint64_t *__restrict__ key_out, uint32_t *__restrict__ val_out,
                                                      FrameCounters *__restrict__ ctr, int capacity, int yx_bits,
                                                      int *__restrict__ big_list, SegBand band) {
    __shared__ uint32_t s_vals[256 / 32][SEG_WARP_MAX];  // values of the path a warp is sorting
    if (ctr->n_fragments > capacity) return;
    const int lane = threadIdx.x & 31;
    uint32_t *const s_val = s_vals[threadIdx.x >> 5];
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_items = (BAND && band.live_paths) ? (uint32_t)ctr->n_live_paths : n_paths;
    const uint32_t n_chunks = (n_items + SEG_CHUNK - 1) / SEG_CHUNK;
    // Full frame: chunks strided over the warps (a hundred thousand chunks, a dozen per warp). Band mode: a band has an
    // n-th of them — two per warp at 16K in eight bands — and a static deal leaves the last round half empty, so the
    // warps take chunks from a ticket instead.
    for (uint32_t ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;; ch += warps) {
        if (BAND && band.live_paths) {
            uint32_t t = 0;
            if (lane == 0) t = (uint32_t)atomicAdd(band.ticket, 1);
            ch = __shfl_sync(0xFFFFFFFFu, t, 0);
        }
        if (ch >= n_chunks) break;
        const uint32_t i0 = ch * SEG_CHUNK;
        uint32_t my_path = 0;
        int my_b = 0, my_e = 0;
        if (BAND && band.live_paths) {  // lanes 0..SEG_CHUNK-1 hold the chunk's paths and their fragment ranges
            if (lane < SEG_CHUNK && i0 + lane < n_items) {
                my_path = band.live_paths[i0 + lane];
                const int2 r = band.live_range[i0 + lane];
                my_b = r.x;
                my_e = r.y;
            }
        } else {  // consecutive paths: lanes 0..SEG_CHUNK hold the chunk's bounds
            my_b = seg[min(i0 + (uint32_t)lane, n_paths)];
        }
#pragma unroll 1
        for (int q = 0; q < SEG_CHUNK; ++q) {
            if (i0 + q >= n_items) break;
            uint32_t p;
            int b, n;
            if (BAND && band.live_paths) {
                p = __shfl_sync(0xFFFFFFFFu, my_path, q);
                b = __shfl_sync(0xFFFFFFFFu, my_b, q);
                n = __shfl_sync(0xFFFFFFFFu, my_e, q) - b;
            } else {
                p = i0 + q;
                b = __shfl_sync(0xFFFFFFFFu, my_b, q);
                n = __shfl_sync(0xFFFFFFFFu, my_b, q + 1) - b;
            }
            if (n <= 0) continue;
            // 32-bit network by size; 0 = it did not fit (64-bit network up to 256 fragments, else the block kernel)
            int sums[3] = {0, 0, 0};
            bool fit = false, done = true;
            if (n <= 32) fit = warp_sort_segment32<1, BAND>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane, s_val, band.geo, sums);
            else if (n <= 64) fit = warp_sort_segment32<2, BAND>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane, s_val, band.geo, sums);
#if SLPR_SEG_R3
            else if (n <= 96) fit = warp_sort_segment32<3, BAND>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane, s_val, band.geo, sums);
#endif
            else if (n <= 128) fit = warp_sort_segment32<4, BAND>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane, s_val, band.geo, sums);
            else if (n <= 256) fit = warp_sort_segment32<8, BAND>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane, s_val, band.geo, sums);
            else if (n <= SEG_WARP_MAX) fit = warp_sort_segment32<16, BAND>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane, s_val, band.geo, sums);
            if (!fit) {
                if (BAND && band.sums) {  // the path's winding sums from a pass of its own (rare)
                    int a = 0, inv = 0, z = 0;
                    for (int i = lane; i < n; i += 32) {
                        const int d = (int)(val_in[b + i] >> 30) - 1;
                        if (d != 0) {
                            const uint32_t yk = (uint32_t)(key_in[b + i] >> band.geo.bits_x) & ((1u << band.geo.bits_y) - 1u);
                            if (yk == (uint32_t)band.geo.ny) z += d;
                            else if (yk == (uint32_t)(band.geo.ny - 1)) inv += d;
                            else a += d;
                        }
                    }
                    sums[0] = __reduce_add_sync(0xFFFFFFFFu, a);
                    sums[1] = __reduce_add_sync(0xFFFFFFFFu, inv);
                    sums[2] = __reduce_add_sync(0xFFFFFFFFu, z);
                }
                if (n <= 32) warp_sort_segment<1>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane);
                else if (n <= 64) warp_sort_segment<2>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane);
                else if (n <= 128) warp_sort_segment<4>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane);
                else if (n <= 256) warp_sort_segment<8>(key_in, val_in, key_out, val_out, b, n, yx_bits, lane);
                else done = false;
            }
            if (BAND && band.sums && lane == 0 && (sums[0] | sums[1] | sums[2])) {
                const int slot = atomicAdd(&ctr->n_band_entries, 1);
                if (slot < XB_CAP) band.sums[slot] = BandEntry{p, sums[0], sums[1], sums[2]};
            }
            if (!done && lane == 0) {
                if (n > SEG_BLOCK_MAX) ctr->sort_fallback = 1;  // the host re-renders with the radix sort
                else big_list[atomicAdd(&ctr->n_big_segments, 1)] = (int)p;
            }
        }
    }
}

__global__ void __launch_bounds__(SEG_BLOCK_THREADS) k_segsort_block(const int *__restrict__ seg,
                                                                     const uint64_t *__restrict__ key_in,
                                                                     const uint32_t *__restrict__ val_in,
                                                                     uint64_t *__restrict__ key_out,
                                                                     uint32_t *__restrict__ val_out,
                                                                     const FrameCounters *__restrict__ ctr, int capacity,
                                                                     int yx_bits, const int *__restrict__ big_list) {
    __shared__ uint64_t s[SEG_BLOCK_MAX];
    if (ctr->n_fragments > capacity || ctr->sort_fallback) return;
    const int nbig = ctr->n_big_segments;
    for (int q = blockIdx.x; q < nbig; q += gridDim.x) {
        const int p = big_list[q];
        const int b = seg[p], n = seg[p + 1] - b;
        int m = 512;  // power of two >= n
        while (m < n) m <<= 1;
        for (int j = threadIdx.x; j < m; j += SEG_BLOCK_THREADS)
            s[j] = (j < n) ? seg_pack(key_in[b + j], val_in[b + j], yx_bits) : ~0ull;
        const uint64_t path_bits = key_in[b] & ~((1ull << yx_bits) - 1);
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < m / 2; t += SEG_BLOCK_THREADS) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
                    const int ip = i | j;
                    const bool asc = (i & k) == 0;
                    const uint64_t a = s[i], c = s[ip];
                    if ((a > c) == asc) { s[i] = c; s[ip] = a; }
                }
                __syncthreads();
            }
        }
        for (int j = threadIdx.x; j < n; j += SEG_BLOCK_THREADS) {
            uint64_t k; uint32_t v;
            seg_unpack(s[j], path_bits, k, v);
            key_out[b + j] = k;
            val_out[b + j] = v;
        }
        __syncthreads();
    }
}

}  // namespace slpr
