// radix.cuh — onesweep-style least-significant-digit radix sort of (u64 key, u32 value) pairs.
// Replaces the reference's per-path O(n^2) odd-even transposition sort
// (workdir/shaders/common/naive_seg_sort_pairs.comp:26-97): a stable sort on the compact key
// (path | row rank | cell x) yields exactly the per-segment signed (key, index) order.
//
// Structure (one read of the keys for all histograms, then one read + one write per 8-bit digit):
//   k_radix_hist      all passes' 256-bin digit histograms in one sweep over the keys
//   k_radix_hist_scan exclusive scan of each pass's bins -> global digit base offsets
//   k_onesweep        per pass: tile-local ranking with warp-shuffle (match.any) histograms, chained
//                     decoupled look-back per digit for the cross-tile prefix, scatter staged through
//                     shared memory so that global stores are runs of consecutive addresses.
// Tiles are taken from an atomic ticket (forward progress of the look-back). No tensor cores: this
// is HBM-bound byte shuffling (12 B read + 12 B written per element per pass).
#pragma once
#include "common.cuh"

namespace slpr {

#ifndef SLPR_RS_THREADS
#define SLPR_RS_THREADS 384
#endif
#ifndef SLPR_RS_ITEMS
#define SLPR_RS_ITEMS 16
#endif
#ifndef SLPR_RS_BLOCKS
#define SLPR_RS_BLOCKS 2
#endif
#ifndef SLPR_RS_MATCH_INSTR
#define SLPR_RS_MATCH_INSTR 1   /* 1: match.any instruction, 0: eight ballots (measured: about equal) */
#endif
#ifndef SLPR_RS_VAL_POS
#define SLPR_RS_VAL_POS 2       /* where the values are loaded: 0 after ranking, 1 before look-back, 2 at the scatter */
#endif
#ifndef SLPR_RS_CHAINS
#define SLPR_RS_CHAINS 1        /* independent ranking chains per warp (each with its own shared histogram row) */
#endif
#ifndef SLPR_RS_LB_WINDOW
#define SLPR_RS_LB_WINDOW 4     /* predecessors polled per look-back round trip */
#endif
constexpr int RS_THREADS = SLPR_RS_THREADS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = SLPR_RS_ITEMS;
constexpr int RS_BLOCKS_PER_SM = SLPR_RS_BLOCKS;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 6144 pairs per tile
constexpr int RS_CHAINS = SLPR_RS_CHAINS;
constexpr int RS_ROWS = RS_WARPS * RS_CHAINS;   // histogram rows: (warp, chain) in tile order
static_assert(RS_ITEMS % RS_CHAINS == 0, "items must split evenly into ranking chains");
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 8;
#define LB_MASK ((1u << 30) - 1)
#define LB_AGG (1u << 30)
#define LB_PREFIX (2u << 30)
constexpr size_t RS_SMEM_BYTES = (size_t)RS_TILE * 12 + (size_t)RS_ROWS * RS_BINS * 4 + 3 * RS_BINS * 4 + 64;

struct SortCount {  // where the element count comes from (device counter for the frame, static for the API)
    const int *n_dev;
    long long n_static;
    int capacity;
    __device__ long long get() const {
        if (!n_dev) return n_static;
        const int n = *n_dev;
        return n > capacity ? -1 : n;
    }
};

struct SortTemp {
    uint32_t *hist;      // [RS_MAX_PASSES][256] digit counts, then exclusive offsets
    uint32_t *lookback;  // [passes][tiles_cap][256], zeroed before the sort
    int *tickets;        // [RS_MAX_PASSES], zeroed before the sort
    int tiles_cap;
};

// ------------------------------------------------------------------------------------------------
// Histogram of every pass's digit in one sweep. Each thread takes 4 consecutive keys (2 x LDG.128)
// and merges equal digits among them before touching shared memory; bins are replicated x4 (by
// lane) so that a warp whose keys share a digit (the path field) does not serialise 32-way.
// ------------------------------------------------------------------------------------------------
constexpr int RH_THREADS = 256;
constexpr int RH_REPL = 4;

__global__ void __launch_bounds__(RH_THREADS) k_radix_hist(const uint64_t *__restrict__ keys, SortCount cnt, int passes,
                                                           uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[RS_MAX_PASSES * RS_BINS * RH_REPL];
    const long long n = cnt.get();
    if (n <= 0) return;
    for (int i = threadIdx.x; i < passes * RS_BINS * RH_REPL; i += RH_THREADS) h[i] = 0;
    __syncthreads();
    const int repl = threadIdx.x & (RH_REPL - 1);
    const long long nvec = (n + 3) / 4;
    for (long long v = (long long)blockIdx.x * RH_THREADS + threadIdx.x; v < nvec; v += (long long)gridDim.x * RH_THREADS) {
        const long long i = v * 4;
        uint64_t k[4];
        int m = 4;
        if (i + 3 < n) {
            const int4 a = *reinterpret_cast<const int4 *>(keys + i);
            const int4 b = *reinterpret_cast<const int4 *>(keys + i + 2);
            k[0] = ((uint64_t)(uint32_t)a.y << 32) | (uint32_t)a.x;
            k[1] = ((uint64_t)(uint32_t)a.w << 32) | (uint32_t)a.z;
            k[2] = ((uint64_t)(uint32_t)b.y << 32) | (uint32_t)b.x;
            k[3] = ((uint64_t)(uint32_t)b.w << 32) | (uint32_t)b.z;
        } else {
            m = (int)(n - i);
            for (int j = 0; j < 4; ++j) k[j] = (j < m) ? keys[i + j] : 0ull;
        }
        for (int p = 0; p < passes; ++p) {
            uint32_t cur = (uint32_t)(k[0] >> (8 * p)) & 0xFFu, run = 1;
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                if (j < m) {
                    const uint32_t d = (uint32_t)(k[j] >> (8 * p)) & 0xFFu;
                    if (d == cur) ++run;
                    else { atomicAdd(&h[(p * RS_BINS + cur) * RH_REPL + repl], run); cur = d; run = 1; }
                }
            }
            atomicAdd(&h[(p * RS_BINS + cur) * RH_REPL + repl], run);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_BINS; i += RH_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int r = 0; r < RH_REPL; ++r) s += h[i * RH_REPL + r];
        if (s) atomicAdd(&hist[i], s);
    }
}

__global__ void __launch_bounds__(RS_BINS) k_radix_hist_scan(uint32_t *__restrict__ hist) {
    __shared__ uint32_t ws[RS_BINS / 32];
    uint32_t *h = hist + blockIdx.x * RS_BINS;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t v = h[t];
    uint32_t s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, s, d);
        if (lane >= d) s += o;
    }
    if (lane == 31) ws[w] = s;
    __syncthreads();
    uint32_t base = 0;
    for (int i = 0; i < w; ++i) base += ws[i];
    h[t] = base + s - v;
}

// Lanes of the warp whose 8-bit digit equals mine: 8 ballots + 8 LOP3 (the match.any instruction is
// microcoded and slower for 8-bit labels).
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
#if SLPR_RS_MATCH_INSTR
    return __match_any_sync(0xFFFFFFFFu, d);
#else
    uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
#endif
}

// ------------------------------------------------------------------------------------------------
// One onesweep pass over digit bits [shift, shift+8).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS, RS_BLOCKS_PER_SM) k_onesweep(const uint64_t *__restrict__ keys_in,
                                                            const uint32_t *__restrict__ vals_in,
                                                            uint64_t *__restrict__ keys_out,
                                                            uint32_t *__restrict__ vals_out, SortCount cnt, int pass,
                                                            int shift, SortTemp tmp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem_raw);                              // [RS_TILE]
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(smem_raw + (size_t)RS_TILE * 8);        // [RS_TILE]
    uint32_t *s_whist = reinterpret_cast<uint32_t *>(smem_raw + (size_t)RS_TILE * 12);      // [RS_ROWS][256]
    uint32_t *s_start = s_whist + RS_ROWS * RS_BINS;                                       // [256] tile digit start
    uint32_t *s_gbase = s_start + RS_BINS;                                                  // [256] global base - start
    uint32_t *s_misc = s_gbase + RS_BINS;                                                   // [256]: warp sums, ticket
    const long long n = cnt.get();
    if (n <= 0) return;
    const long long ntiles = (n + RS_TILE - 1) / RS_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t *ghist = tmp.hist + pass * RS_BINS;
    uint32_t *lookback = tmp.lookback + (size_t)pass * tmp.tiles_cap * RS_BINS;
    const uint32_t lt = lanemask_lt();

    while (true) {
        if (tid == 0) s_misc[32] = (uint32_t)atomicAdd(&tmp.tickets[pass], 1);
        for (int i = tid; i < RS_ROWS * RS_BINS; i += RS_THREADS) s_whist[i] = 0;
        __syncthreads();
        const long long tile = (long long)s_misc[32];
        if (tile >= ntiles) break;
        const long long base = tile * RS_TILE;
        const int valid = (int)min((long long)RS_TILE, n - base);

        // ---- load keys, warp-striped: item i of this lane sits at warp_base + i*32 + lane
        const int wbase = warp * 32 * RS_ITEMS;
        uint64_t key[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int o = wbase + i * 32 + lane;
            key[i] = (o < valid) ? keys_in[base + o] : ~0ull;
        }
        // ---- rank inside the warp's chunk: match.any groups lanes with the same digit; the lowest lane
        //      of each group bumps the warp-private bin, everyone takes bin_before + rank in group.
        //      The chunk is cut into RS_CHAINS consecutive sub-chunks with their own histogram rows, ranked
        //      in an interleaved loop: the serial shared-memory read-modify-write chains overlap.
        uint16_t rank[RS_ITEMS];
        constexpr int PER = RS_ITEMS / RS_CHAINS;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
#pragma unroll
            for (int g = 0; g < RS_CHAINS; ++g) {
                const int it = g * PER + i;
                uint32_t *wh = s_whist + (warp * RS_CHAINS + g) * RS_BINS;
                const uint32_t d = (uint32_t)(key[it] >> shift) & 0xFFu;
                const uint32_t peers = match_digit(d);
                const int leader = __ffs(peers) - 1;
                uint32_t before = 0;
                if (lane == leader) { before = wh[d]; wh[d] = before + __popc(peers); }
                before = __shfl_sync(0xFFFFFFFFu, before, leader);
                rank[it] = (uint16_t)(before + __popc(peers & lt));
            }
            __syncwarp();
        }
#if SLPR_RS_VAL_POS == 0
        // values are loaded now so that their latency overlaps the histogram scan and the look-back
        uint32_t val[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int o = wbase + i * 32 + lane;
            val[i] = (o < valid) ? vals_in[base + o] : 0u;
        }
#endif
        __syncthreads();

        // ---- per digit: exclusive scan over warps (in place), tile count, look-back
        uint32_t count = 0;
        if (tid < RS_BINS) {
#pragma unroll
            for (int w = 0; w < RS_ROWS; ++w) {
                const uint32_t c = s_whist[w * RS_BINS + tid];
                s_whist[w * RS_BINS + tid] = count;
                count += c;
            }
            if (tid == RS_BINS - 1) count -= (uint32_t)(RS_TILE - valid);  // padding keys all land in bin 255
            // publish the aggregate early so that successors can make progress
            st_status32(lookback + (size_t)tile * RS_BINS + tid, ((tile == 0) ? LB_PREFIX : LB_AGG) | count);
            // exclusive scan of the 256 tile counts -> start of each digit inside the staged tile
            uint32_t s = (tid == RS_BINS - 1) ? count + (uint32_t)(RS_TILE - valid) : count;
            const uint32_t mine = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, s, d);
                if (lane >= d) s += o;
            }
            if (lane == 31) s_misc[warp] = s;
            s_start[tid] = s - mine;  // warp-local exclusive; fixed below
        }
        __syncthreads();
#if SLPR_RS_VAL_POS == 1
        // values are fetched here: after the ranking (whose registers are dead now) and before the
        // look-back, whose L2 round trips hide the load latency
        uint32_t val[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int o = wbase + i * 32 + lane;
            val[i] = (o < valid) ? vals_in[base + o] : 0u;
        }
#endif
        if (tid < RS_BINS) {
            uint32_t wb = 0;
            for (int w = 0; w < warp; ++w) wb += s_misc[w];
            const uint32_t start = s_start[tid] + wb;
            // fold the digit's start into every row offset: the scatter then needs one table read per key
#pragma unroll
            for (int w = 0; w < RS_ROWS; ++w) s_whist[w * RS_BINS + tid] += start;
            // chained look-back for this digit; SLPR_RS_LB_WINDOW predecessors are polled per round trip
            uint32_t excl = 0;
            if (tile > 0) {
                long long t = tile - 1;
                bool done = false;
                while (!done) {
                    uint32_t w[SLPR_RS_LB_WINDOW];
#pragma unroll
                    for (int k = 0; k < SLPR_RS_LB_WINDOW; ++k)
                        w[k] = (t - k >= 0) ? ld_status32(lookback + (size_t)(t - k) * RS_BINS + tid) : LB_PREFIX;
#pragma unroll
                    for (int k = 0; k < SLPR_RS_LB_WINDOW; ++k) {
                        if (done) break;
                        if ((w[k] >> 30) == 0) break;  // not published yet: poll again from here
                        excl += w[k] & LB_MASK;
                        --t;
                        if ((w[k] >> 30) == 2) done = true;
                    }
                }
                st_status32(lookback + (size_t)tile * RS_BINS + tid, LB_PREFIX | ((excl + count) & LB_MASK));
            }
            s_gbase[tid] = ghist[tid] + excl - start;
            s_start[tid] = start;
        }
        __syncthreads();

        // ---- scatter into shared memory in digit order
#if SLPR_RS_VAL_POS == 2
        uint32_t val[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int o = wbase + i * 32 + lane;
            val[i] = (o < valid) ? vals_in[base + o] : 0u;
        }
#endif
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & 0xFFu;
            const uint32_t pos = s_whist[(warp * RS_CHAINS + i / (RS_ITEMS / RS_CHAINS)) * RS_BINS + d] + rank[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
        __syncthreads();
        // ---- write out: consecutive threads write consecutive addresses inside each digit run
        for (int j = tid; j < valid; j += RS_THREADS) {
            const uint64_t k = s_keys[j];
            const uint32_t d = (uint32_t)(k >> shift) & 0xFFu;
            const size_t g = (size_t)(s_gbase[d] + (uint32_t)j);
            keys_out[g] = k;
            vals_out[g] = s_vals[j];
        }
        __syncthreads();
    }
}

}  // namespace slpr
