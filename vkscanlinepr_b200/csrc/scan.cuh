// scan.cuh — single-pass exclusive prefix sum with decoupled look-back (replaces the reference's
// one-workgroup serial naive_scan.comp:27-73, dispatched 3x per frame at SR.cpp:347,496,563).
//
// One kernel, templated on an "Op" that says how an element is loaded (possibly computed on the
// fly from other arrays) and what is done with its exclusive prefix. This lets mark + scan #3 +
// emit (MARK/GEN shaders) run as a single streaming pass.
//
// Tile = 512 threads x 4 int4 vectors = 8192 elements; every global access of the plain int32 op
// is a 128-bit coalesced LDG/STG. Tiles are handed out by an atomic ticket so that a tile's
// predecessors are always scheduled (forward progress for the look-back spin). The tile status is
// one 64-bit word: bits 63..62 state (0 empty, 1 tile aggregate, 2 inclusive prefix), bits 61..0
// value; one aligned 64-bit store publishes both, so no fence is required.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace slpr {

#ifndef SLPR_SCAN_THREADS
#define SLPR_SCAN_THREADS 512
#endif
#ifndef SLPR_SCAN_VECS
#define SLPR_SCAN_VECS 4
#endif
#ifndef SLPR_SCAN_BLOCKS
#define SLPR_SCAN_BLOCKS 2
#endif
#ifndef SLPR_SCAN_LOOK
#define SLPR_SCAN_LOOK 4
#endif
constexpr int SCAN_THREADS = SLPR_SCAN_THREADS;
constexpr int SCAN_VECS_MAX = 8;                              // int4 vectors per thread (Op::VECS <= this)
constexpr int SCAN_TILE_MIN = SCAN_THREADS * 4;               // smallest tile any op may use (VECS = 1)
template <class Op> constexpr int scan_tile() { return SCAN_THREADS * Op::VECS * 4; }

#define ST_MASK ((1ull << 62) - 1)
#define ST_AGG (1ull << 62)
#define ST_PREFIX (2ull << 62)

struct ScanTemp {
    unsigned long long *status;  // [max tiles], zeroed before the launch
    int *ticket;                 // zeroed before the launch
};

__device__ __forceinline__ unsigned long long warp_incl_scan_u64(unsigned long long v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if ((int)lane_id() >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

// Round 2. What bounds a decoupled look-back scan on this part is the speed at which the "prefix known" frontier
// moves along the chain of tiles: one look-back window per L2 round trip (~0.7 us). With the first version's
// 4096-element tiles and 32-tile window that is 32 x 32 KB per round trip = 1.5 TB/s — exactly what it measured
// (27 % of the HBM roofline at 35 M elements). Now: tiles of 8192 elements (512 threads x 16) and a window of
// 128 tiles (every lane polls four predecessors per round trip): 128 x 64 KB per round trip, above HBM speed, so
// the kernel is bound by its loads and stores. Layout: a warp owns 512 consecutive elements, four 128-bit vectors
// per lane (each instruction of the warp reads 512 contiguous bytes). Outputs are 32-bit (they wrap like the
// shader's int adds); the tile status carries the exact 62-bit total (capacity check of scan #1).
constexpr int SCAN_LOOK = SLPR_SCAN_LOOK;  // predecessors polled per lane and round trip

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS, Op::MIN_BLOCKS) k_lookback_scan(Op op, ScanTemp tmp) {
    constexpr int SCAN_VECS = Op::VECS;
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_VECS * 4;
    constexpr int WARPS = SCAN_THREADS / 32;
    static_assert(WARPS <= 32, "warp partials must fit one warp");
    __shared__ unsigned long long s_wtot[WARPS];  // exact total of each warp's 512 elements
    __shared__ uint32_t s_woff[WARPS];            // exclusive offset of each warp inside the tile (low 32 bits)
    __shared__ uint32_t s_prefix;
    __shared__ long long s_tile;

    const long long n = op.count();
    if (n < 0) return;  // the op signalled "skip" (capacity overflow)
    const long long ntiles = (n == 0) ? 1 : (n + SCAN_TILE - 1) / SCAN_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    while (true) {
        if (tid == 0) s_tile = (long long)atomicAdd(tmp.ticket, 1);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= ntiles) break;
        const long long wbase = tile * SCAN_TILE + (long long)warp * (SCAN_VECS * 128);

        uint32_t x[SCAN_VECS][4];
#pragma unroll
        for (int v = 0; v < SCAN_VECS; ++v) op.load(wbase + (long long)(v * 32 + lane) * 4, n, x[v]);

        uint32_t ex[SCAN_VECS];  // exclusive offset of vector v's first element inside the warp's chunk
        unsigned long long exact = 0;
        uint32_t run = 0;
#pragma unroll
        for (int v = 0; v < SCAN_VECS; ++v) {
            exact += (unsigned long long)x[v][0] + x[v][1] + x[v][2] + x[v][3];
            const uint32_t sv = x[v][0] + x[v][1] + x[v][2] + x[v][3];
            uint32_t incl = sv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            ex[v] = run + incl - sv;
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        exact = warp_sum_u64(exact);
        if (lane == 0) s_wtot[warp] = exact;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long p = (lane < WARPS) ? s_wtot[lane] : 0ull;
            const unsigned long long pi = warp_incl_scan_u64(p);
            if (lane < WARPS) s_woff[lane] = (uint32_t)(pi - p);
            const unsigned long long tile_total = __shfl_sync(0xFFFFFFFFu, pi, 31);
            unsigned long long *status = tmp.status;
            if (lane == 0) st_status64(status + tile, ((tile == 0) ? ST_PREFIX : ST_AGG) | (tile_total & ST_MASK));
            unsigned long long excl = 0;
            if (tile > 0) {
                long long look = tile - 1;
                while (true) {
                    const long long first_idx = look - (long long)lane * SCAN_LOOK;  // lane l: tiles look-4l .. look-4l-3
                    unsigned long long w[SCAN_LOOK];
                    bool empty;
                    do {
                        empty = false;
#pragma unroll
                        for (int q = 0; q < SCAN_LOOK; ++q) {
                            const long long idx = first_idx - q;
                            w[q] = (idx >= 0) ? ld_status64(status + idx) : ST_PREFIX;
                            empty |= (w[q] >> 62) == 0;
                        }
                    } while (__any_sync(0xFFFFFFFFu, empty));
                    unsigned long long part = 0;  // nearest -> farthest, up to and including this lane's first inclusive prefix
                    bool has_prefix = false;
#pragma unroll
                    for (int q = 0; q < SCAN_LOOK; ++q) {
                        if (!has_prefix) {
                            part += w[q] & ST_MASK;
                            has_prefix = (w[q] >> 62) == 2;
                        }
                    }
                    const uint32_t pm = __ballot_sync(0xFFFFFFFFu, has_prefix);
                    const int first = pm ? (__ffs(pm) - 1) : 32;
                    excl += warp_sum_u64((lane <= first) ? part : 0ull);
                    if (pm) break;
                    look -= 32 * SCAN_LOOK;
                }
                if (lane == 0) st_status64(status + tile, ST_PREFIX | ((excl + tile_total) & ST_MASK));
            }
            if (lane == 0) {
                s_prefix = (uint32_t)excl;
                if (tile == ntiles - 1) op.finish(n, (excl + tile_total) & ST_MASK);
            }
        }
        __syncthreads();
        const uint32_t before = s_prefix + s_woff[warp];
#pragma unroll
        for (int v = 0; v < SCAN_VECS; ++v) {
            uint32_t e[4];
            e[0] = before + ex[v];
            e[1] = e[0] + x[v][0];
            e[2] = e[1] + x[v][1];
            e[3] = e[2] + x[v][2];
            op.store(wbase + (long long)(v * 32 + lane) * 4, n, e);
        }
        // s_tile / s_wtot / s_woff / s_prefix are rewritten only after the next trip's first barrier (s_tile: by
        // thread 0 before it, but every thread read it before the look-back barrier above)
    }
}

// ------------------------------------------------------------------------------------------------
// Op 1: plain int32 exclusive scan, out[i] = sum_{j<i} in[j], i in [0,n] (naive_scan.comp semantics;
// sums wrap modulo 2^32 like the shader's int adds). In-place is allowed.
// ------------------------------------------------------------------------------------------------
struct ScanI32Op {
    static constexpr int VECS = SLPR_SCAN_VECS, MIN_BLOCKS = SLPR_SCAN_BLOCKS;
    const int *in;
    int *out;
    long long n_static;
    int *total_out;  // optional: also receives the total (e.g. FrameCounters::n_fragments)
    int capacity;    // optional (with overflow_out): raise the flag when total > capacity
    int *overflow_out;
    __device__ long long count() const { return n_static; }
    __device__ void load(long long i, long long n, uint32_t x[4]) const {
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(in + i) & 15) == 0)) {
            const int4 v = ld_stream(reinterpret_cast<const int4 *>(in + i));
            x[0] = (uint32_t)v.x; x[1] = (uint32_t)v.y; x[2] = (uint32_t)v.z; x[3] = (uint32_t)v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = (i + k < n) ? (uint32_t)in[i + k] : 0u;
        }
    }
    __device__ void store(long long i, long long n, const uint32_t e[4]) const {
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(out + i) & 15) == 0)) {
            st_stream(reinterpret_cast<int4 *>(out + i), make_int4((int)e[0], (int)e[1], (int)e[2], (int)e[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k < n) out[i + k] = (int)e[k];
        }
    }
    __device__ void finish(long long n, unsigned long long total) const {
        out[n] = (int)total;
        // the exact (62-bit) total decides; a total beyond int32 is reported as INT_MAX so that every
        // `n_fragments > capacity` guard downstream still fires
        if (total_out) *total_out = (total > 0x7FFFFFFFull) ? 0x7FFFFFFF : (int)total;
        if (overflow_out && total > (unsigned long long)(capacity < 0 ? 0 : capacity)) *overflow_out = 1;
    }
};


// ------------------------------------------------------------------------------------------------
// k_scan_tma — the same exclusive scan for LARGE arrays, restructured so that HBM never idles (round 2).
//
// Measured on the B200 (profiles/README.md): in k_lookback_scan a tile's life is a series of latencies — ticket,
// loads, block scan, look-back polls, stores — and whatever the tile shape it stays at 27-43 % of the HBM roofline at
// 35 M elements. Instrumented, the cost is the look-back itself: one poll of a window of predecessors takes ~1.4 us
// under the traffic of a few hundred polling CTAs, a tile needs 2-5 of them, and with 32 KB tiles that is more
// than the tile's transfer time. So here
//   * a CTA's unit of work is a CHUNK of two 8192-element sub-tiles (64 KB in, 64 KB out): half as many look-backs,
//     each over half as many predecessors, and a window of 160 chunks per poll (consecutive entries per lane
//     group, i.e. whole 32-byte sectors);
//   * input arrives in a seven-stage shared-memory ring (224 KB) by 1-D bulk copies (cp.async.bulk, the TMA unit; one
//     thread issues, an mbarrier per stage counts the bytes); a stage is re-armed the moment the scan warps have
//     taken its sub-tile into registers, so 5-6 sub-tiles (160-192 KB per SM) are always in flight;
//   * eight scan warps scan chunk j locally in registers, publish its aggregate — and then finish chunk j - 1
//     (add its global prefix, stream it out with 128-bit stores), whose prefix
//   * a ninth warp has resolved meanwhile with the decoupled look-back. Hand-overs are mbarriers (waiting in
//     hardware, not spinning).
// Chunks are dealt round-robin (chunk = blockIdx.x + j * gridDim.x, one CTA per SM), so consecutive chunks live in
// different CTAs and their aggregates appear in parallel; this needs every CTA resident, which is why the frame's
// own small scans (curve counts, band tables: at most a few MB) keep the ticketed k_lookback_scan.
// ------------------------------------------------------------------------------------------------
constexpr int ST_SCAN_WARPS = 8;
constexpr int ST_SCAN_THREADS = ST_SCAN_WARPS * 32;
constexpr int ST_THREADS = ST_SCAN_THREADS + 32;  // + the look-back warp
constexpr int ST_VECS = 8;                         // int4 vectors per scan thread and sub-tile
constexpr int ST_SUB = ST_SCAN_THREADS * ST_VECS * 4;  // 8192 elements = 32 KB
constexpr int ST_SUBS = 2;                         // sub-tiles per chunk
constexpr int ST_TILE = ST_SUB * ST_SUBS;          // elements per chunk (the unit of the look-back chain)
constexpr int ST_STAGES = 7;                       // 32 KB sub-tile stages in shared memory
constexpr size_t ST_SMEM_BYTES = (size_t)ST_STAGES * ST_SUB * 4;
#ifndef SLPR_ST_LOOK
#define SLPR_ST_LOOK 5  /* 32-entry groups polled per L2 round trip by the look-back warp: window = 32 x this chunks */
#endif
constexpr int ST_LOOK = SLPR_ST_LOOK;
#ifndef SLPR_ST_CTAS
#define SLPR_ST_CTAS 1
#endif

__global__ void __launch_bounds__(ST_THREADS, 1) k_scan_tma(ScanI32Op op, ScanTemp tmp) {
    extern __shared__ __align__(128) unsigned char st_smem[];
    __shared__ __align__(8) uint64_t s_full[ST_STAGES];
    __shared__ __align__(8) uint64_t s_agg_bar[2], s_pref_bar[2];  // hand-over scan warps <-> look-back warp
    __shared__ unsigned long long s_wtot[2][ST_SCAN_WARPS];
    __shared__ unsigned long long s_agg[2];
    __shared__ uint32_t s_pref[2];

    const long long n = op.count();
    if (n < 0) return;
    const long long nsub = (n == 0) ? 1 : (n + ST_SUB - 1) / ST_SUB;  // sub-tiles; [0, nsub_full) are whole (bulk copies)
    const long long nsub_full = n / ST_SUB;
    const long long nchunks = (nsub + ST_SUBS - 1) / ST_SUBS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long G = gridDim.x;
    const long long first = blockIdx.x;
    const int n_mine = (first < nchunks) ? (int)((nchunks - first + G - 1) / G) : 0;
    unsigned long long *const status = tmp.status;

    if (tid == 0) {
        for (int s = 0; s < ST_STAGES; ++s) mbar_init(&s_full[s], 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&s_agg_bar[s], 1); mbar_init(&s_pref_bar[s], 1); }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == ST_SCAN_WARPS) {  // ---- the look-back warp
        for (int j = 0; j < n_mine; ++j) {
            const long long chunk = first + (long long)j * G;
            mbar_wait(&s_agg_bar[j & 1], (uint32_t)(j >> 1) & 1u);
            const unsigned long long chunk_total = *reinterpret_cast<volatile unsigned long long *>(&s_agg[j & 1]);
            unsigned long long excl = 0;
            if (chunk > 0) {
                long long look = chunk - 1;
                while (true) {
                    // group q of the window: chunks look - 32 q - lane (a lane group reads 32 consecutive status words)
                    unsigned long long w[ST_LOOK];
                    bool empty;
                    do {
                        empty = false;
#pragma unroll
                        for (int q = 0; q < ST_LOOK; ++q) {
                            const long long idx = look - 32 * q - lane;
                            w[q] = (idx >= 0) ? ld_status64(status + idx) : ST_PREFIX;
                            empty |= (w[q] >> 62) == 0;
                        }
                    } while (__any_sync(0xFFFFFFFFu, empty));
                    // nearest -> farthest: whole groups until the first one that holds an inclusive prefix, then its lanes up to it
                    unsigned long long part = 0;
                    bool found = false;
#pragma unroll
                    for (int q = 0; q < ST_LOOK; ++q) {
                        if (!found) {
                            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, (w[q] >> 62) == 2);
                            const int firstp = pm ? (__ffs(pm) - 1) : 32;
                            if (lane <= firstp) part += w[q] & ST_MASK;
                            found = pm != 0;
                        }
                    }
                    excl += warp_sum_u64(part);
                    if (found) break;
                    look -= 32 * ST_LOOK;
                }
                if (lane == 0) st_status64(status + chunk, ST_PREFIX | ((excl + chunk_total) & ST_MASK));
            }
            if (lane == 0) {
                s_pref[j & 1] = (uint32_t)excl;
                mbar_arrive(&s_pref_bar[j & 1]);  // (release: the prefix is visible to whoever sees the phase flip)
                if (chunk == nchunks - 1) op.finish(n, (excl + chunk_total) & ST_MASK);
            }
        }
        return;
    }

    // ---- the scan warps. My sub-tiles are numbered m = 0, 1, ... (chunk j = sub-tiles 2j, 2j + 1); sub-tile m uses
    //      stage m % ST_STAGES for the (m / ST_STAGES)-th time.
    const long long m_total = (long long)n_mine * ST_SUBS;
    auto global_sub = [&](long long m) { return (first + (m / ST_SUBS) * G) * ST_SUBS + (m % ST_SUBS); };
    auto issue_sub = [&](long long m) {  // thread 0: bulk copy of my m-th sub-tile (if it is a whole one) into its stage
        if (m >= m_total) return;
        const long long sub = global_sub(m);
        if (sub < nsub_full) {
            const int st = (int)(m % ST_STAGES);
            mbar_arrive_expect_tx(&s_full[st], ST_SUB * 4);
            tma_load_1d(st_smem + (size_t)st * ST_SUB * 4, op.in + sub * ST_SUB, ST_SUB * 4, &s_full[st]);
        }
    };
    if (tid == 0)
        for (int m = 0; m < ST_STAGES; ++m) issue_sub(m);
    const int4 *const ring = reinterpret_cast<const int4 *>(st_smem);
    uint32_t x_old[ST_SUBS][ST_VECS][4], ex_old[ST_SUBS][ST_VECS];  // chunk j - 1, parked until its prefix is known
#pragma unroll
    for (int u = 0; u < ST_SUBS; ++u)
#pragma unroll
        for (int v = 0; v < ST_VECS; ++v) { ex_old[u][v] = 0; x_old[u][v][0] = x_old[u][v][1] = x_old[u][v][2] = x_old[u][v][3] = 0; }
    long long m = 0;
    for (int j = 0; j <= n_mine; ++j) {
        const long long chunk = first + (long long)j * G;
        uint32_t x[ST_SUBS][ST_VECS][4] = {}, ex[ST_SUBS][ST_VECS] = {};
        if (j < n_mine) {  // ---- local scan of chunk j
            uint32_t carry = 0;
            unsigned long long carry_exact = 0;
#pragma unroll
            for (int u = 0; u < ST_SUBS; ++u, ++m) {
                const long long sub = chunk * ST_SUBS + u;
                if (sub < nsub) {  // (uniform) the last chunk may have one sub-tile only
                    const int st = (int)(m % ST_STAGES);
                    const long long wbase = sub * ST_SUB + (long long)warp * (ST_VECS * 128);
                    if (sub < nsub_full) {
                        mbar_wait(&s_full[st], (uint32_t)(m / ST_STAGES) & 1u);
                        const int4 *const buf = ring + (size_t)st * (ST_SUB / 4) + warp * (ST_VECS * 32);
#pragma unroll
                        for (int v = 0; v < ST_VECS; ++v) {
                            const int4 q = buf[v * 32 + lane];
                            x[u][v][0] = (uint32_t)q.x; x[u][v][1] = (uint32_t)q.y; x[u][v][2] = (uint32_t)q.z; x[u][v][3] = (uint32_t)q.w;
                        }
                    } else {  // the ragged last sub-tile: guarded loads
#pragma unroll
                        for (int v = 0; v < ST_VECS; ++v) op.load(wbase + (long long)(v * 32 + lane) * 4, n, x[u][v]);
                    }
                    unsigned long long exact = 0;
                    uint32_t run = 0;
#pragma unroll
                    for (int v = 0; v < ST_VECS; ++v) {
                        exact += (unsigned long long)x[u][v][0] + x[u][v][1] + x[u][v][2] + x[u][v][3];
                        const uint32_t sv = x[u][v][0] + x[u][v][1] + x[u][v][2] + x[u][v][3];
                        uint32_t incl = sv;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                            if (lane >= d) incl += o;
                        }
                        ex[u][v] = run + incl - sv;
                        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
                    }
                    exact = warp_sum_u64(exact);
                    if (lane == 0) s_wtot[m & 1][warp] = exact;
                    named_barrier_sync(1, ST_SCAN_THREADS);  // warp totals visible; everybody has taken the sub-tile out of its stage
                    if (tid == 0) issue_sub(m + ST_STAGES);   // so the stage is re-armed at once
                    unsigned long long before = 0, total = 0;
#pragma unroll
                    for (int w = 0; w < ST_SCAN_WARPS; ++w) {
                        const unsigned long long t = s_wtot[m & 1][w];
                        if (w < warp) before += t;
                        total += t;
                    }
                    const uint32_t base = carry + (uint32_t)before;
#pragma unroll
                    for (int v = 0; v < ST_VECS; ++v) ex[u][v] += base;  // chunk-local exclusive prefix of the vector's first element
                    carry += (uint32_t)total;
                    carry_exact += total;
                }
            }
            if (tid == 0) {  // the aggregate goes out at once; the look-back warp resolves the prefix while chunk j + 1 is scanned
                st_status64(status + chunk, ((chunk == 0) ? ST_PREFIX : ST_AGG) | (carry_exact & ST_MASK));
                s_agg[j & 1] = carry_exact;
                mbar_arrive(&s_agg_bar[j & 1]);
            }
        }
        if (j > 0) {  // ---- finish chunk j - 1: its prefix has had a whole chunk's time to arrive
            const long long pchunk = chunk - G;
            mbar_wait(&s_pref_bar[(j - 1) & 1], (uint32_t)((j - 1) >> 1) & 1u);
            const uint32_t pref = *reinterpret_cast<volatile uint32_t *>(&s_pref[(j - 1) & 1]);
#pragma unroll
            for (int u = 0; u < ST_SUBS; ++u) {
                const long long sub = pchunk * ST_SUBS + u;
                if (sub < nsub) {
                    const long long wbase = sub * ST_SUB + (long long)warp * (ST_VECS * 128);
#pragma unroll
                    for (int v = 0; v < ST_VECS; ++v) {
                        uint32_t e[4];
                        e[0] = pref + ex_old[u][v];
                        e[1] = e[0] + x_old[u][v][0];
                        e[2] = e[1] + x_old[u][v][1];
                        e[3] = e[2] + x_old[u][v][2];
                        op.store(wbase + (long long)(v * 32 + lane) * 4, n, e);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < ST_SUBS; ++u)
#pragma unroll
            for (int v = 0; v < ST_VECS; ++v) {
                ex_old[u][v] = ex[u][v];
#pragma unroll
                for (int k = 0; k < 4; ++k) x_old[u][v][k] = x[u][v][k];
            }
    }
}

}  // namespace slpr
