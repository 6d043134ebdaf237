// spans.cuh — everything after the sort in ONE streaming kernel with two chained decoupled
// look-backs per tile:
//   chain W  shuffle_fragment.comp:17-27 + scan #2 (SR.cpp:479-506): the winding delta travels in the
//            two top bits of the sorted value, so the gather disappears; the prefix sum is global and
//            unsegmented exactly like the reference's;
//   chain F  mark_merged_fragment_and_span.comp:22-94 + scan #3 (SR.cpp:545-573) +
//            gen_merged_fragment_and_span.comp:33-103: the flags need the global winding prefix, so
//            their tile aggregate is published after chain W resolves; both flag counts travel in one
//            2x31-bit word and the draw records are written straight from the tile.
// A tile is 512 threads x 8 consecutive fragments (128-bit loads of 2 keys / 4 values); tiles are
// ticketed for forward progress. Nothing but the draw records (16 B each) is written in the fast
// path; the reference planes (winding scan, flags, flag scan, sorted key/index) only with taps.
#pragma once
#include "geom.cuh"
#include "scan.cuh"
#include "raster.cuh"
#include "bands.cuh"

namespace slpr {

// Tile = SLPR_SP_THREADS x SLPR_SP_ITEMS fragments. Without the look-back chain the tiles are independent and small blocks
// win (their barriers cost less and the ticket balances finer): k_spans on the 1 M-curve 4K frame, threads x blocks/SM:
// 512 x 2: 0.231 ms, 256 x 4: 0.221, 192 x 5: 0.219, 128 x 8: 0.208, 64 x 16: 0.204 but the winding prefix over four times the
// tiles costs what is saved (k_wsum + k_wscan 0.026 -> 0.032 -> 0.072); 128 x 8 is the best frame (1.024 -> 1.011 ms).
#ifndef SLPR_SP_THREADS
#define SLPR_SP_THREADS 128
#endif
#ifndef SLPR_SP_BLOCKS
#define SLPR_SP_BLOCKS 8
#endif
#ifndef SLPR_SP_LOOK
#define SLPR_SP_LOOK 1
#endif
#ifndef SLPR_SP_SLEEP
#define SLPR_SP_SLEEP 0
#endif
#ifndef SLPR_SP_WPRE
#define SLPR_SP_WPRE 1   /* 1: winding prefix per tile from k_wsum + k_wscan (no chain W); 0: chained look-back */
#endif
#ifndef SLPR_SP_FAKE
#define SLPR_SP_FAKE 0   /* timing experiments only: skip the look-back (wrong results) */
#endif
#ifndef SLPR_SP_NOSTORE
#define SLPR_SP_NOSTORE 0 /* timing experiments only: skip the record stores */
#endif
constexpr int SP_THREADS = SLPR_SP_THREADS;
#ifndef SLPR_SP_ITEMS
#define SLPR_SP_ITEMS 8
#endif
constexpr int SP_ITEMS = SLPR_SP_ITEMS;  // consecutive fragments per thread (4 or 8)
constexpr int SP_BLOCKS = SLPR_SP_BLOCKS;
constexpr int SP_LOOK = SLPR_SP_LOOK;  // tile states polled per lane per look-back round trip (window = 32 * SP_LOOK tiles)
constexpr int SP_TILE = SP_THREADS * SP_ITEMS;
#ifndef SLPR_SP_STAGE
#define SLPR_SP_STAGE 1 /* draw records staged in shared memory and written as full rows */
#endif
constexpr int SP_WARP_RECORDS = 32 * SP_ITEMS * 2;  // a fragment emits at most one fragment and one span record
constexpr int SP_STAGE_BYTES = SLPR_SP_STAGE ? (SP_THREADS / 32) * 3 * SP_WARP_RECORDS * 4 : 0;

struct SpanTaps {
    int *wn;      // [nf+1] plane 3 after shuffle + scan #2
    int *sidx;    // plane 1 after sort
    int *skey32;  // plane 0 after sort
    int *flags;   // [2*nf] = [frag | span]                     (MARK:92-93)
    int *scan3;   // [2*nf+1]; second half needs + n_out_frag   (k_scan3_fixup)
};

struct SpanTemp {
    const int *wprefix;            // [tiles] exclusive winding prefix per tile (k_wsum + k_wscan), SLPR_SP_WPRE
    unsigned long long *status_w;  // chain W tile states (zeroed per frame), !SLPR_SP_WPRE
    unsigned long long *status_f;  // chain F tile states
    int *ticket;
};

// One decoupled look-back by a full warp: publishes this tile's aggregate, returns the exclusive
// prefix of all earlier tiles (in every lane) and publishes the inclusive prefix. The chain of
// tiles ripples at (window tiles) per L2 round trip; the window is 32 * SP_LOOK tiles (measured: 32 is best):
// lane l polls tiles look-SP_LOOK*l .. with SP_LOOK independent loads per round trip.
__device__ __forceinline__ unsigned long long warp_lookback(unsigned long long *status, long long tile,
                                                           unsigned long long tile_total, int lane) {
    if (lane == 0) st_status64(status + tile, ((tile == 0) ? ST_PREFIX : ST_AGG) | (tile_total & ST_MASK));
    unsigned long long excl = 0;
    if (tile > 0 && !SLPR_SP_FAKE) {
        long long look = tile - 1;
        while (true) {
            const long long first_idx = look - (long long)lane * SP_LOOK;
            unsigned long long w[SP_LOOK];
            bool empty;
            do {
                empty = false;
#pragma unroll
                for (int q = 0; q < SP_LOOK; ++q) {
                    const long long idx = first_idx - q;
                    w[q] = (idx >= 0) ? ld_status64(status + idx) : ST_PREFIX;
                    empty |= (w[q] >> 62) == 0;
                }
                if (SLPR_SP_SLEEP && empty) __nanosleep(SLPR_SP_SLEEP);
            } while (__any_sync(0xFFFFFFFFu, empty));
            // this lane's partial: nearest -> farthest, up to and including its first inclusive prefix
            unsigned long long part = 0;
            bool has_prefix = false;
#pragma unroll
            for (int q = 0; q < SP_LOOK; ++q) {
                if (!has_prefix) {
                    part += w[q] & ST_MASK;
                    has_prefix = (w[q] >> 62) == 2;
                }
            }
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, has_prefix);
            const int first = pm ? (__ffs(pm) - 1) : 32;
            excl += warp_sum_u64((lane <= first) ? part : 0ull);
            if (pm) break;
            look -= 32 * SP_LOOK;
        }
        if (lane == 0) st_status64(status + tile, ST_PREFIX | ((excl + tile_total) & ST_MASK));
    }
    return excl;
}

struct KeyFields {
    uint32_t path;
    int x, y;
};

__device__ __forceinline__ KeyFields decode_key(const KeyLayout &L, uint64_t k) {
    KeyFields f;
    const uint32_t xk = (uint32_t)(k & ((1ull << L.bits_x) - 1));
    const uint32_t yk = (uint32_t)((k >> L.bits_x) & ((1ull << L.bits_y) - 1));
    f.path = (uint32_t)(k >> (L.bits_x + L.bits_y));
    if (yk == (uint32_t)(L.ny - 1)) { f.x = 0x7FFF; f.y = 0x7FFF; }  // the invalid key decodes to (32767, 32767), MARK:41-42
    else { f.x = (int)xk * 2 - FRAG_SIZE; f.y = (yk == (uint32_t)L.ny) ? 0 : (int)(yk + 1) * 2; }
    return f;
}

// FILL: the kernel marks the cells its records cover itself (stage 5 coverage fused in) — with the record's PATH
//       (+1), not its index: records are emitted in path order and a path has one colour, so "the later record
//       wins" (opaque overwrite in primitive order, SR.cpp:893-895) is "the highest path wins", and k_resolve
//       finds the colour in fill_info[path]. Nobody then reads the 16-byte draw records, so they are
// REC:  only written when asked for (SLPR_FLAG_RECORDS / taps), or for the separate coverage pass (!FILL).
// BLEND (with FILL): translucent paths append list nodes instead (SLPR_FLAG_BLEND, raster.cuh mark_cells32).
template <bool FILL, bool REC, bool TAPS, bool BLEND = false>
__global__ void __launch_bounds__(SP_THREADS, SP_BLOCKS) k_spans(const uint64_t *__restrict__ skey,
                                                         const uint32_t *__restrict__ sval,
                                                         const uint32_t *__restrict__ fill_info,
                                                         int4 *__restrict__ records, FrameCounters *__restrict__ ctr,
                                                         KeyLayout L, int width, int height, int capacity, SpanTaps taps,
                                                         SpanTemp tmp, const int *__restrict__ band_corr, uint32_t n_paths,
                                                         uint32_t *__restrict__ cells, int cw, BandTable btab, BlendList bl) {
    __shared__ uint32_t s_warp[SP_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ long long s_tile;
#if SLPR_SP_STAGE
    extern __shared__ uint32_t s_stage[];  // [warps][3][SP_WARP_RECORDS]
#endif
    const int nf = ctr->n_fragments;
    if (frame_void(ctr, capacity)) return;
    const long long n = nf;
    const long long ntiles = (n == 0) ? 1 : (n + SP_TILE - 1) / SP_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbp = btab.count ? *btab.count : 0;  // exact bands, sparse exchange: break points of this frame (usually 0)
    constexpr bool CHAIN = REC || TAPS;  // record indices wanted: the flag-count look-back chain across tiles
    int acc_frag = 0, acc_span = 0;      // !CHAIN: this warp's counts over all the block's tiles (lane 31)

    while (true) {
        // tiles in ticket order: with the look-back chain a tile's predecessors are then always running; without it
        // (independent tiles) the ticket still balances the load better than a static deal (measured 0.256 vs 0.270 ms)
        if (tid == 0) s_tile = (long long)atomicAdd(tmp.ticket, 1);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= ntiles) break;
        const long long i0 = tile * SP_TILE + (long long)tid * SP_ITEMS;

        // ---- load 8 consecutive sorted keys and values
        uint64_t k[SP_ITEMS + 1];  // k[0] = key of element i0-1
        uint32_t dpack = 0;        // 8 x 2-bit (delta + 1)
        uint32_t rpack = 0;        // 8 x 1-bit fill rule of the fragment's path (carried in bit 29 of the value)
        if (i0 + SP_ITEMS <= n) {
#pragma unroll
            for (int j = 0; j < SP_ITEMS; j += 2) {
                const int4 q = *reinterpret_cast<const int4 *>(skey + i0 + j);
                k[j + 1] = ((uint64_t)(uint32_t)q.y << 32) | (uint32_t)q.x;
                k[j + 2] = ((uint64_t)(uint32_t)q.w << 32) | (uint32_t)q.z;
            }
#pragma unroll
            for (int j = 0; j < SP_ITEMS; j += 4) {
                const int4 q = *reinterpret_cast<const int4 *>(sval + i0 + j);
                const uint32_t v[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    dpack |= (v[e] >> 30) << (2 * (j + e));
                    rpack |= ((v[e] >> 29) & 1u) << (j + e);
                    if (TAPS && taps.sidx) taps.sidx[i0 + j + e] = (int)(v[e] & VAL_INDEX_MASK);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < SP_ITEMS; ++j) {
                const bool in = i0 + j < n;
                k[j + 1] = in ? skey[i0 + j] : 0ull;
                const uint32_t v = in ? sval[i0 + j] : (1u << 30);
                dpack |= (v >> 30) << (2 * j);
                rpack |= ((v >> 29) & 1u) << j;
                if (TAPS && in && taps.sidx) taps.sidx[i0 + j] = (int)(v & VAL_INDEX_MASK);
            }
        }
        {   // key of the element before this thread's run: neighbour lane, or global for lane 0
            const uint32_t lo = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)k[SP_ITEMS], 1);
            const uint32_t hi = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)(k[SP_ITEMS] >> 32), 1);
            k[0] = ((uint64_t)hi << 32) | lo;
            if (lane == 0) k[0] = (i0 > 0 && i0 - 1 < n) ? skey[i0 - 1] : 0ull;
        }

        // ---- chain W: exclusive prefix of the winding deltas (int32, wraps like the shader's adds)
        int dsum = 0;
#pragma unroll
        for (int j = 0; j < SP_ITEMS; ++j) dsum += (int)((dpack >> (2 * j)) & 3u) - 1;
        int wincl = dsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xFFFFFFFFu, wincl, d);
            if (lane >= d) wincl += o;
        }
        if (lane == 31) s_warp[warp] = (uint32_t)wincl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t p = (lane < SP_THREADS / 32) ? s_warp[lane] : 0u;
            uint32_t pi = p;
#pragma unroll
            for (int d = 1; d < SP_THREADS / 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pi, d);
                if (lane >= d) pi += o;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, pi, SP_THREADS / 32 - 1);
#if SLPR_SP_WPRE
            const unsigned long long excl = (unsigned long long)(uint32_t)tmp.wprefix[tile];
#else
            const unsigned long long excl = warp_lookback(tmp.status_w, tile, (unsigned long long)total, lane);
#endif
            if (lane < SP_THREADS / 32) s_warp[lane] = pi - p;  // exclusive offset of each warp
            if (lane == 0) {
                s_prefix = excl;
                if (!SLPR_SP_WPRE && tile == ntiles - 1) {
                    const int wtotal = (int)(uint32_t)(excl + total);
                    ctr->wn_total = wtotal;
                    if (TAPS && taps.wn) taps.wn[n] = wtotal;
                }
            }
        }
        __syncthreads();
        int wn = (int)(uint32_t)s_prefix + (int)s_warp[warp] + (wincl - dsum);  // winding left of element i0
        if (CHAIN) __syncthreads();  // s_warp is reused by chain F (without it, only after the next ticket barrier)

        // ---- flags (MARK:38-93) from the keys and the winding prefix
        uint32_t fmask = 0, smask = 0;
        {
            uint32_t bt_key = 0xFFFFFFFFu;  // (path << 1 | row 0) of the cached break-point look-up
            int bt_corr = 0;
            KeyFields a = decode_key(L, k[0]);
#pragma unroll
            for (int j = 0; j < SP_ITEMS; ++j) {
                const KeyFields b = decode_key(L, k[j + 1]);
                if (i0 + j < n) {
                    if (TAPS && taps.wn) taps.wn[i0 + j] = wn;
                    const bool oob = (b.x < 0 || b.y < 0 || b.x >= width || b.y >= height);  // MARK:45,69
                    uint32_t frag, span = 0;
                    if (i0 + j == 0) {
                        frag = oob ? 0u : 1u;  // MARK:44-51
                    } else {
                        frag = (!oob && k[j] != k[j + 1]) ? 1u : 0u;  // MARK:69-77 (the compact key holds the path)
                        const bool even_odd = (rpack >> j) & 1u;  // fill_rule[path] == 1
                        // MARK:82 (rule values other than 0/1 never set the flag there; the loader only produces 0/1)
                        // exact row bands (bands.cuh): the deltas of the other bands' fragments that sort before this one
                        int wf = wn;
                        if (band_corr) wf += band_corr[(b.y == 0 ? n_paths : 0u) + b.path];
                        else if (nbp > 0) {
                            const uint32_t bk = (b.path << 1) | (b.y == 0 ? 1u : 0u);
                            if (bk != bt_key) { bt_key = bk; bt_corr = band_table_lookup(btab, nbp, b.path, b.y == 0); }
                            wf += bt_corr;
                        }
                        const bool wn_flag = even_odd ? ((wf & 1) != 0) : (wf != 0);
                        span = (a.y == b.y && (a.x + FRAG_SIZE) < b.x && a.path == b.path && wn_flag) ? 1u : 0u;  // MARK:84
                    }
                    fmask |= frag << j;
                    smask |= span << j;
                    if (TAPS && taps.flags) {
                        taps.flags[i0 + j] = (int)frag;
                        taps.flags[n + i0 + j] = (int)span;
                        uint32_t pp;
                        taps.skey32[i0 + j] = unpack_key32(L, k[j + 1], pp);
                    }
                }
                wn += (int)((dpack >> (2 * j)) & 3u) - 1;
                a = b;
            }
        }

        // ---- chain F: exclusive prefix of (frag count | span count << 16) inside the tile, 2x31 bits across tiles.
        //      Only the record indices need it. When no records are written (the default big-frame path: the cells carry
        //      the path, REC and TAPS off) a tile depends on NO other tile: the chain — one look-back window per ~1.4 us
        //      poll, 135 hops for 4321 tiles, the longest serial path of the kernel — is dropped and the two counts are
        //      summed per warp over the block's tiles and added to the frame counters once, at the end.
        const uint32_t cnt = (uint32_t)__popc(fmask) | ((uint32_t)__popc(smask) << 16);
        uint32_t cincl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, cincl, d);
            if (lane >= d) cincl += o;
        }
        int frag_before = 0, span_before = 0;
        if (CHAIN) {
        if (lane == 31) s_warp[warp] = cincl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t p = (lane < SP_THREADS / 32) ? s_warp[lane] : 0u;
            uint32_t pi = p;
#pragma unroll
            for (int d = 1; d < SP_THREADS / 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pi, d);
                if (lane >= d) pi += o;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, pi, SP_THREADS / 32 - 1);
            const unsigned long long total64 = (unsigned long long)(total & 0xFFFFu) | ((unsigned long long)(total >> 16) << 31);
            const unsigned long long excl = warp_lookback(tmp.status_f, tile, total64, lane);
            if (lane < SP_THREADS / 32) s_warp[lane] = pi - p;
            if (lane == 0) {
                s_prefix = excl;
                if (tile == ntiles - 1) {  // SR.cpp:578-580
                    const unsigned long long tot = excl + total64;
                    const int nfrag = (int)(tot & 0x7FFFFFFFu), nspan = (int)((tot >> 31) & 0x7FFFFFFFu);
                    ctr->n_out_frag = nfrag;
                    ctr->n_span = nspan;
                    ctr->n_records = nfrag + nspan;
                    if (TAPS && taps.scan3) { taps.scan3[n] = 0; taps.scan3[2 * n] = nspan; }  // + n_out_frag in k_scan3_fixup
                }
            }
        }
        __syncthreads();
        const unsigned long long tp = s_prefix;
        const uint32_t local = s_warp[warp] + (cincl - cnt);
        frag_before = (int)(tp & 0x7FFFFFFFu) + (int)(local & 0xFFFFu);
        span_before = (int)((tp >> 31) & 0x7FFFFFFFu) + (int)(local >> 16);
        } else if (lane == 31) {
            acc_frag += (int)(cincl & 0xFFFFu);
            acc_span += (int)(cincl >> 16);
        }

        // ---- emit the draw records (GEN:42-102). A thread's records are consecutive in the output
        //      and so are the threads of a warp, so the warp first lays its records out in shared memory
        //      (three 32-bit planes: position, width | fragment ordinal, colour) and then writes them as
        //      full 512-byte rows: 16-byte stores scattered at 128-byte strides cost the kernel a third
        //      of its time (measured with SLPR_SP_NOSTORE).

#if SLPR_SP_STAGE
        uint32_t *const wp = s_stage + warp * (3 * SP_WARP_RECORDS);
        const uint32_t wexcl = cincl - cnt;                           // packed counts of the lanes before this one
        const int wfrag = (int)(wexcl & 0xFFFFu);
        uint32_t slot = (uint32_t)wfrag + (wexcl >> 16);              // first staging slot of this thread
        const uint32_t wtot_p = __shfl_sync(0xFFFFFFFFu, cincl, 31);
        const uint32_t wtot = (wtot_p & 0xFFFFu) + (wtot_p >> 16);    // records of the whole warp
        const int warp_frag_base = frag_before - wfrag;               // fragments emitted before this warp
        const long long gbase = (long long)frag_before + span_before - slot;
#endif
        if (((fmask | smask) || (TAPS && taps.scan3)) && !(SLPR_SP_NOSTORE && frag_before >= 0)) {
            KeyFields a = decode_key(L, k[0]);
#pragma unroll
            for (int j = 0; j < SP_ITEMS; ++j) {
                const KeyFields b = decode_key(L, k[j + 1]);
                const uint32_t frag = (fmask >> j) & 1u, span = (smask >> j) & 1u;
                if (TAPS && taps.scan3 && i0 + j < n) {
                    taps.scan3[i0 + j] = frag_before;
                    taps.scan3[n + i0 + j] = span_before;
                }
                if (frag | span) {
#if SLPR_SP_STAGE
                    if (frag) {  // GEN:77: (y<<16 | x, 2, rgba, inclusive fragment index)
                        wp[slot] = ((uint32_t)b.y << 16) | (uint32_t)b.x;
                        wp[SP_WARP_RECORDS + slot] = 2u | ((uint32_t)(frag_before - warp_frag_base + 1) << 16);
                        wp[2 * SP_WARP_RECORDS + slot] = b.path;
                        ++slot;
                    }
                    if (span) {  // GEN:85-102: from the previous fragment's right edge to this fragment
                        const int xs = max(0, a.x + FRAG_SIZE);
                        wp[slot] = ((uint32_t)a.y << 16) | (uint32_t)xs;
                        wp[SP_WARP_RECORDS + slot] = (uint32_t)(b.x - xs) & 0xFFFFu;
                        wp[2 * SP_WARP_RECORDS + slot] = b.path;
                        ++slot;
                    }
#else
                    const int fill = (int)fill_info[b.path];
                    const int oi = frag_before + span_before;  // GEN:64-66
                    if (frag)  // GEN:77: (y<<16 | x, 2, rgba, inclusive fragment index)
                        records[oi] = make_int4((int)(((uint32_t)b.y << 16) | (uint32_t)b.x), 2, fill, frag_before + 1);
                    if (span) {  // GEN:85-102: from the previous fragment's right edge to this fragment
                        const int xs = max(0, a.x + FRAG_SIZE);
                        records[oi + (int)frag] = make_int4((int)(((uint32_t)a.y << 16) | (uint32_t)xs), b.x - xs, fill, 0);
                    }
#endif
                }
                frag_before += (int)frag;
                span_before += (int)span;
                a = b;
            }
        }
#if SLPR_SP_STAGE
        __syncwarp();
        if (!SLPR_SP_NOSTORE) {
            // One pass over the warp's staged records, 32 at a time. REC: the record goes out (full 512-byte rows; its
            // colour is fetched here, one gather per record). FILL: its cells are marked — atomicMax(cell, path + 1),
            // fire and forget; narrow records by their own lane, wide spans by the whole warp (as k_fill_cells).
            for (uint32_t q0 = 0; q0 < wtot; q0 += 32) {
                const uint32_t q = q0 + (uint32_t)lane;
                int cx0 = 0, ncell = 0, cy = 0;
                uint32_t prio = 0, alpha = 255u;
                if (q < wtot) {
                    const uint32_t pos = wp[q], w1 = wp[SP_WARP_RECORDS + q], ord = w1 >> 16, path = wp[2 * SP_WARP_RECORDS + q];
                    if (REC)
                        records[gbase + q] = make_int4((int)pos, (int)(w1 & 0xFFFFu), (int)fill_info[path],
                                                       ord ? warp_frag_base + (int)ord : 0);
                    if (FILL) {
                        const int X = (int)(pos & 0xFFFFu), Y = (int)pos >> 16;  // VERT:27
                        if (Y >= 0 && Y < height) {
                            cx0 = X >> 1;
                            ncell = min((X + (int)(w1 & 0xFFFFu)) >> 1, cw) - cx0;
                            cy = Y >> 1;
                        }
                        prio = path + 1u;
                        if (BLEND && ncell > 0) alpha = fill_info[path] >> 24;
                    }
                }
                if (FILL) mark_cells32<BLEND>(cells, cw, cx0, ncell, cy, prio, alpha, bl, ctr, (uint32_t)lane);
            }
        }
        __syncwarp();
#endif
        // the __syncthreads after the next ticket fetch orders the reuse of s_warp / s_prefix
    }
    if (!CHAIN && lane == 31 && (acc_frag | acc_span)) {  // SR.cpp:578-580, as sums
        atomicAdd(&ctr->n_out_frag, acc_frag);
        atomicAdd(&ctr->n_span, acc_span);
        atomicAdd(&ctr->n_records, acc_frag + acc_span);
    }
}

// ------------------------------------------------------------------------------------------------
// Winding prefix per span tile without a chain: tile sums of the deltas (k_wsum, reads the 4-byte
// values once), then one block scans the few thousand tile sums (k_wscan). k_spans then starts
// every tile with its global winding prefix known and needs a single look-back chain (the flag
// counts), whose aggregate no longer waits for another chain to resolve.
// ------------------------------------------------------------------------------------------------
constexpr int WSUM_THREADS = 256;
__global__ void __launch_bounds__(WSUM_THREADS) k_wsum(const uint32_t *__restrict__ sval, const FrameCounters *__restrict__ ctr,
                                                       int capacity, int *__restrict__ wsum) {
    const int nf = ctr->n_fragments;
    if (frame_void(ctr, capacity)) return;
    const long long n = nf;
    const long long ntiles = (n == 0) ? 1 : (n + SP_TILE - 1) / SP_TILE;
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (WSUM_THREADS / 32);
    // a warp per span tile (the tile size is k_spans' business; this kernel's launch shape is its own)
    for (long long tile = (long long)blockIdx.x * (WSUM_THREADS / 32) + (threadIdx.x >> 5); tile < ntiles; tile += warps) {
        const long long t0 = tile * SP_TILE;
        int dsum = 0;
        if (t0 + SP_TILE <= n) {
#pragma unroll 4
            for (int k = lane; k < SP_TILE / 4; k += 32) {
                const int4 q = ld_stream(reinterpret_cast<const int4 *>(sval + t0) + k);
                dsum += (int)((uint32_t)q.x >> 30) + (int)((uint32_t)q.y >> 30) + (int)((uint32_t)q.z >> 30) + (int)((uint32_t)q.w >> 30) - 4;
            }
        } else {
            for (long long i = t0 + lane; i < n; i += 32) dsum += (int)(sval[i] >> 30) - 1;
        }
        dsum = __reduce_add_sync(0xFFFFFFFFu, dsum);
        if (lane == 0) wsum[tile] = dsum;
    }
}

// One block scans the tile sums in chunks of 8192 that pass through shared memory: global accesses coalesced, every
// thread scans eight neighbours, one block-wide scan of the 1024 partial sums per chunk.
__global__ void __launch_bounds__(1024) k_wscan(FrameCounters *__restrict__ ctr, int capacity, int *__restrict__ wsum,
                                                int *__restrict__ wn_tap) {
    constexpr int PER = 8, CHUNK = 1024 * PER;
    __shared__ int s_v[CHUNK + CHUNK / 32];  // (one pad word per 32: a thread's eight neighbours start on different banks)
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int nf = ctr->n_fragments;
    if (frame_void(ctr, capacity)) return;
    const long long n = nf;
    const int ntiles = (int)((n == 0) ? 1 : (n + SP_TILE - 1) / SP_TILE);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    for (int base = 0; base < ntiles; base += CHUNK) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int j = tid + 1024 * k;
            s_v[j + (j >> 5)] = (base + j < ntiles) ? wsum[base + j] : 0;
        }
        __syncthreads();
        int t[PER], v = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int j = tid * PER + k;
            t[k] = s_v[j + (j >> 5)];
            v += t[k];
        }
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += o;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int carry = s_carry;
        int run = carry + s_w[warp] + incl - v;  // exclusive winding prefix of this thread's first tile
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int j = tid * PER + k;
            s_v[j + (j >> 5)] = run;
            run += t[k];
        }
        __syncthreads();
        if (tid == 1023) s_carry = run;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int j = tid + 1024 * k;
            if (base + j < ntiles) wsum[base + j] = s_v[j + (j >> 5)];
        }
        __syncthreads();
    }
    if (tid == 0) {
        ctr->wn_total = s_carry;
        if (wn_tap) wn_tap[n] = s_carry;
    }
}

// scan3[nf + i] += n_out_frag for i in [0, nf] (tap only): turns the two separate counts into the
// reference's single scan over the concatenated [frag | span] flag array (SR.cpp:545-573).
__global__ void k_scan3_fixup(const FrameCounters *__restrict__ ctr, int capacity, int *__restrict__ scan3) {
    const int nf = ctr->n_fragments;
    if (frame_void(ctr, capacity)) return;
    const int add = ctr->n_out_frag;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nf; i += gridDim.x * blockDim.x) scan3[nf + i] += add;
}

}  // namespace slpr
