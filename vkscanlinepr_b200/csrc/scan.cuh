// scan.cuh — single-pass exclusive prefix sum with decoupled look-back (replaces the reference's
// one-workgroup serial naive_scan.comp:27-73, dispatched 3x per frame at SR.cpp:347,496,563).
//
// One kernel, templated on an "Op" that says how an element is loaded (possibly computed on the
// fly from other arrays) and what is done with its exclusive prefix. This lets mark + scan #3 +
// emit (MARK/GEN shaders) run as a single streaming pass.
//
// Tile = 256 threads x 4 int4 vectors = 4096 elements; every global access of the plain int32 op
// is a 128-bit coalesced LDG/STG. Tiles are handed out by an atomic ticket so that a tile's
// predecessors are always scheduled (forward progress for the look-back spin). The tile status is
// one 64-bit word: bits 63..62 state (0 empty, 1 tile aggregate, 2 inclusive prefix), bits 61..0
// value; one aligned 64-bit store publishes both, so no fence is required.
#pragma once
#include "common.cuh"

namespace slpr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_VECS_MAX = 4;                              // int4 vectors per thread (Op::VECS <= this)
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

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS, Op::MIN_BLOCKS) k_lookback_scan(Op op, ScanTemp tmp) {
    constexpr int SCAN_VECS = Op::VECS;
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_VECS * 4;
    static_assert(SCAN_VECS >= 1 && SCAN_VECS <= SCAN_VECS_MAX, "partials must fit one warp");
    __shared__ unsigned long long s_part[32];  // (vector, warp) partials, scanned by warp 0
    __shared__ unsigned long long s_prefix;
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
        const long long base = tile * SCAN_TILE;

        unsigned long long x[SCAN_VECS][4];
        typename Op::Aux aux[SCAN_VECS];
#pragma unroll
        for (int v = 0; v < SCAN_VECS; ++v) op.load(base + (long long)(v * SCAN_THREADS + tid) * 4, n, x[v], aux[v]);

        unsigned long long s[SCAN_VECS], incl[SCAN_VECS];
#pragma unroll
        for (int v = 0; v < SCAN_VECS; ++v) s[v] = x[v][0] + x[v][1] + x[v][2] + x[v][3];
#pragma unroll
        for (int v = 0; v < SCAN_VECS; ++v) incl[v] = warp_incl_scan_u64(s[v]);
        if (lane == 31) {
#pragma unroll
            for (int v = 0; v < SCAN_VECS; ++v) s_part[v * (SCAN_THREADS / 32) + warp] = incl[v];
        }
        __syncthreads();
        if (warp == 0) {
            const unsigned long long p = (lane < SCAN_VECS * (SCAN_THREADS / 32)) ? s_part[lane] : 0ull;
            const unsigned long long pi = warp_incl_scan_u64(p);
            s_part[lane] = pi - p;  // exclusive offset of (vector, warp) inside the tile
            const unsigned long long tile_total = __shfl_sync(0xFFFFFFFFu, pi, 31);
            unsigned long long *status = tmp.status;
            if (lane == 0) st_status64(status + tile, ((tile == 0) ? ST_PREFIX : ST_AGG) | (tile_total & ST_MASK));
            unsigned long long excl = 0;
            if (tile > 0) {
                long long look = tile - 1;
                while (true) {
                    const long long idx = look - lane;
                    unsigned long long w = (idx >= 0) ? ld_status64(status + idx) : ST_PREFIX;
                    while (__any_sync(0xFFFFFFFFu, (w >> 62) == 0)) {
                        if ((w >> 62) == 0) w = ld_status64(status + idx);
                    }
                    const uint32_t pm = __ballot_sync(0xFFFFFFFFu, (w >> 62) == 2);
                    const int first = pm ? (__ffs(pm) - 1) : 32;
                    excl += warp_sum_u64((lane <= first) ? (w & ST_MASK) : 0ull);
                    if (pm) break;
                    look -= 32;
                }
                if (lane == 0) st_status64(status + tile, ST_PREFIX | ((excl + tile_total) & ST_MASK));
            }
            if (lane == 0) {
                s_prefix = excl;
                if (tile == ntiles - 1) op.finish(n, (excl + tile_total) & ST_MASK);
            }
        }
        __syncthreads();
        const unsigned long long tile_prefix = s_prefix;
#pragma unroll
        for (int v = 0; v < SCAN_VECS; ++v) {
            unsigned long long e[4];
            e[0] = (tile_prefix + s_part[v * (SCAN_THREADS / 32) + warp] + (incl[v] - s[v])) & ST_MASK;
            e[1] = (e[0] + x[v][0]) & ST_MASK;
            e[2] = (e[1] + x[v][1]) & ST_MASK;
            e[3] = (e[2] + x[v][2]) & ST_MASK;
            op.store(base + (long long)(v * SCAN_THREADS + tid) * 4, n, e, x[v], aux[v]);
        }
        __syncthreads();  // s_tile / s_part are reused by the next tile
    }
}

struct NoAux {};

// ------------------------------------------------------------------------------------------------
// Op 1: plain int32 exclusive scan, out[i] = sum_{j<i} in[j], i in [0,n] (naive_scan.comp semantics;
// sums wrap modulo 2^32 like the shader's int adds). In-place is allowed.
// ------------------------------------------------------------------------------------------------
struct ScanI32Op {
    using Aux = NoAux;
    static constexpr int VECS = 4, MIN_BLOCKS = 4;
    const int *in;
    int *out;
    long long n_static;
    int *total_out;  // optional: also receives the total (e.g. FrameCounters::n_fragments)
    int capacity;    // optional (with overflow_out): raise the flag when total > capacity
    int *overflow_out;
    __device__ long long count() const { return n_static; }
    __device__ void load(long long i, long long n, unsigned long long x[4], Aux &) const {
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(in + i) & 15) == 0)) {
            const int4 v = ld_stream(reinterpret_cast<const int4 *>(in + i));
            x[0] = (uint32_t)v.x; x[1] = (uint32_t)v.y; x[2] = (uint32_t)v.z; x[3] = (uint32_t)v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = (i + k < n) ? (uint32_t)in[i + k] : 0u;
        }
    }
    __device__ void store(long long i, long long n, const unsigned long long e[4], const unsigned long long *,
                          const Aux &) const {
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

}  // namespace slpr
