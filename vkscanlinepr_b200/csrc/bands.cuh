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


// ================================================================================================
// Device-side exchange (round 2): no host hand-over, no collective call, no pass over all paths.
//
// Every (path, row) of a closed path has winding deltas that cancel, so a band's three sums per path are zero for
// all but the few paths with a residue (the MI0:340 slip: a few dozen per million curves). The exchange is
// therefore SPARSE: a band publishes only its non-zero (path, normal, invalid, row-0) entries — typically a few
// hundred bytes instead of 12 bytes for every path of the scene — by storing them straight into a mailbox in every
// other band's HBM (peer-mapped over NVLink; CUDA IPC between the one-process-per-GPU ranks) followed by a
// system-scope flag. Each band then merges what it received into a sorted table of break points
//   bp_path[i], bp_cum[i] = sum of D over listed paths <= bp_path[i], bp_n[i], bp_z[i]
// and k_spans looks its corrections up by binary search (corrN[p] = cum(before p) + n[p], corrZ[p] = cum(before p)
// + z[p], as in the dense scheme above), only when the table is not empty.
//   k_band_sums_sparse  per-path sums of the band's fragments -> local list of non-zero entries
//   k_band_publish      local list -> mailbox[slot][me] of every band (remote stores), then flag = frame_seq + 1
//   k_band_merge        waits for every band's flag (bounded spin), merges, sorts, writes the break-point table
//   k_band_done / k_band_wait_done   "my pixels of frame seq are in the root's frame buffer" flags (gather)
// Two mailbox slots (frame_seq & 1): a band cannot publish frame i+2 before every band has consumed frame i
// (it must first merge frame i+1, which needs everybody's frame i+1 list, published after their frame i merge).
// A frame that is void on one band (fragment-buffer overflow, a path too long for the segmented sort, more than
// XB_CAP entries) is published as such; every band then marks its own frame void (FrameCounters::band_void) and
// the caller renders that frame again on all bands (SLPR_ERR_RETRY).
// ================================================================================================
constexpr int XB_MAX_BANDS = 16;
constexpr int XB_CAP = 2048;               // entries one band may publish per frame
constexpr int XB_TOTAL = 4096;             // entries one band can merge per frame (sum over the other bands)
constexpr unsigned long long XB_VOID = 1ull << 63;
#ifndef SLPR_XB_TIMEOUT_NS
#define SLPR_XB_TIMEOUT_NS 2000000000ull   /* a band that never publishes must not hang the others: give up after 2 s */
#endif

struct __align__(16) BandEntry { uint32_t path; int a, inv, z; };

struct BandInbox {                 // one per (slot, sender), in the RECEIVER's memory
    unsigned long long flag;       // frame_seq + 1 once count and entries are complete; | XB_VOID: sender's frame is void
    uint32_t count, pad;
    BandEntry e[XB_CAP];
};
struct BandMailbox {
    BandInbox in[2][XB_MAX_BANDS];
    unsigned long long done[2][XB_MAX_BANDS];  // gather: band r's pixels of frame seq are in this band's frame buffer
};

struct BandPeers {
    BandMailbox *box[XB_MAX_BANDS];  // every band's mailbox as addressable from THIS device (own or peer-mapped)
    int n_bands, me, root;
};

// break-point table of one frame (device memory of the band)
struct BandTable {
    uint32_t *path;  // [XB_TOTAL] ascending, unique
    int *cum;        // [XB_TOTAL] inclusive prefix of D = sum over the other bands of (a + inv + z)
    int *n;          // [XB_TOTAL] sum over the bands below of a
    int *z;          // [XB_TOTAL] sum over the other bands of (a + inv)
    const int *count;  // -> FrameCounters::n_band_bp
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// correction of a fragment of path p (row 0 fragments sort after everything else of their path)
__device__ __forceinline__ int band_table_lookup(const BandTable &t, int nbp, uint32_t p, bool row0) {
    int lo = 0, hi = nbp;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t.path[mid] < p) lo = mid + 1; else hi = mid;
    }
    int c = lo > 0 ? t.cum[lo - 1] : 0;
    if (lo < nbp && t.path[lo] == p) c += row0 ? t.z[lo] : t.n[lo];
    return c;
}

// (radix-sort mode only: with the segmented sort k_segsort_warp forms these sums while it has each path in hand)
// A warp takes 32 listed paths per trip — one lane looks up one path's fragment range — and then sums the non-empty
// ones together, one after the other.
__global__ void __launch_bounds__(256) k_band_sums_sparse(const int *__restrict__ seg, uint32_t n_paths,
                                                          const uint64_t *__restrict__ key, const uint32_t *__restrict__ val,
                                                          FrameCounters *__restrict__ ctr, int capacity, KeyLayout L,
                                                          BandEntry *__restrict__ list, const uint32_t *__restrict__ live_paths) {
    if (ctr->n_fragments > capacity) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t ymask = (1ull << L.bits_y) - 1;
    const uint32_t n_items = live_paths ? (uint32_t)ctr->n_live_paths : n_paths;
    const uint32_t n_groups = (n_items + 31u) >> 5;
    for (uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
        const uint32_t item = g * 32u + (uint32_t)lane;
        uint32_t my_p = 0;
        int my_b = 0, my_e = 0;
        if (item < n_items) {
            my_p = live_paths ? live_paths[item] : item;
            my_b = seg[my_p];
            my_e = seg[my_p + 1];
        }
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, my_e > my_b);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t p = __shfl_sync(0xFFFFFFFFu, my_p, src);
            const int b = __shfl_sync(0xFFFFFFFFu, my_b, src), e = __shfl_sync(0xFFFFFFFFu, my_e, src);
            int a = 0, inv = 0, z = 0;
            for (int i = b + lane; i < e; i += 32) {
                const int d = (int)(val[i] >> 30) - 1;
                if (d == 0) continue;
                const uint32_t yk = (uint32_t)((key[i] >> L.bits_x) & ymask);
                if (yk == (uint32_t)L.ny) z += d;
                else if (yk == (uint32_t)(L.ny - 1)) inv += d;
                else a += d;
            }
            a = __reduce_add_sync(0xFFFFFFFFu, a);
            inv = __reduce_add_sync(0xFFFFFFFFu, inv);
            z = __reduce_add_sync(0xFFFFFFFFu, z);
            if (lane == 0 && (a | inv | z)) {
                const int slot = atomicAdd(&ctr->n_band_entries, 1);
                if (slot < XB_CAP) list[slot] = BandEntry{p, a, inv, z};
            }
        }
    }
}

// grid = n_bands blocks: block b stores this band's list into band b's mailbox (its own included)
__global__ void __launch_bounds__(256) k_band_publish(const FrameParams *__restrict__ P, const FrameCounters *__restrict__ ctr,
                                                      int capacity, int radix_mode, const BandEntry *__restrict__ list, BandPeers peers) {
    const uint32_t seq = (uint32_t)P->frame_seq;
    BandInbox *dst = &peers.box[blockIdx.x]->in[seq & 1u][peers.me];
    const int n = ctr->n_band_entries;
    const bool is_void = ctr->overflow != 0 || ctr->n_fragments > capacity || n > XB_CAP || (ctr->stat_huge != 0 && !radix_mode);
    const int cnt = is_void ? 0 : n;
    const uint4 *src = reinterpret_cast<const uint4 *>(list);
    uint4 *out = reinterpret_cast<uint4 *>(dst->e);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) out[i] = src[i];
    if (threadIdx.x == 0) dst->count = (uint32_t)cnt;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys_u64(&dst->flag, ((unsigned long long)seq + 1ull) | (is_void ? XB_VOID : 0ull));
}

// One block. Dynamic shared memory: XB_TOTAL x (8 + 4 + 4 + 4 + 4) bytes.
constexpr int XB_MERGE_THREADS = 1024;
constexpr size_t XB_MERGE_SMEM = (size_t)XB_TOTAL * 24;
__global__ void __launch_bounds__(XB_MERGE_THREADS) k_band_merge(const FrameParams *__restrict__ P, FrameCounters *__restrict__ ctr,
                                                                 BandPeers peers, BandTable tab) {
    extern __shared__ __align__(16) unsigned char xb_smem[];
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(xb_smem);           // path << 32 | position
    int *s_d = reinterpret_cast<int *>(xb_smem + (size_t)XB_TOTAL * 8);                     // a + inv + z
    int *s_n = s_d + XB_TOTAL;                                                              // a if the sender is below me
    int *s_z = s_n + XB_TOTAL;                                                              // a + inv
    int *s_aux = s_z + XB_TOTAL;                                                            // scan scratch
    __shared__ int s_off[XB_MAX_BANDS + 1];
    __shared__ int s_bad;
    __shared__ int s_wsum[32];
    const uint32_t seq = (uint32_t)P->frame_seq;
    const BandMailbox *mine = peers.box[peers.me];
    const int tid = threadIdx.x, G = peers.n_bands;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // ---- every band's list for this frame has arrived (bounded spin: a band that died must not hang this one)
    if (tid < G) {
        const unsigned long long want = (unsigned long long)seq + 1ull;
        const unsigned long long *f = &mine->in[seq & 1u][tid].flag;
        const unsigned long long t0 = global_timer_ns();
        unsigned long long v;
        int bad = 0;
        while (((v = ld_acquire_sys_u64(f)) & ~XB_VOID) != want) {
            if (global_timer_ns() - t0 > SLPR_XB_TIMEOUT_NS) { bad = 2; break; }
            __nanosleep(200);
        }
        if (!bad && (v & XB_VOID)) bad = 1;
        if (bad) atomicMax(&s_bad, bad);
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) { ctr->band_void = s_bad; ctr->n_band_bp = 0; }
        return;
    }
    if (tid == 0) {
        int off = 0;
        for (int r = 0; r < G; ++r) {
            s_off[r] = off;
            if (r != peers.me) off += (int)mine->in[seq & 1u][r].count;
        }
        s_off[G] = off;
        if (off > XB_TOTAL) s_bad = 3;
    }
    __syncthreads();
    const int S = s_off[G];
    if (s_bad) {
        if (tid == 0) { ctr->band_void = s_bad; ctr->n_band_bp = 0; }
        return;
    }
    if (S == 0) {  // the common case: no band saw a residue
        if (tid == 0) ctr->n_band_bp = 0;
        return;
    }
    int m = 1;
    while (m < S) m <<= 1;
    for (int r = 0; r < G; ++r) {
        if (r == peers.me) continue;
        const BandInbox *in = &mine->in[seq & 1u][r];
        const int n_r = s_off[r + 1] - s_off[r];
        for (int i = tid; i < n_r; i += XB_MERGE_THREADS) {
            const BandEntry e = in->e[i];
            const int pos = s_off[r] + i;
            s_key[pos] = ((unsigned long long)e.path << 32) | (unsigned)pos;
            s_d[pos] = e.a + e.inv + e.z;
            s_n[pos] = (r < peers.me) ? e.a : 0;
            s_z[pos] = e.a + e.inv;
        }
    }
    for (int i = S + tid; i < m; i += XB_MERGE_THREADS) s_key[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < m / 2; t += XB_MERGE_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ip = i | j;
                const bool asc = (i & k) == 0;
                const unsigned long long a = s_key[i], c = s_key[ip];
                if ((a > c) == asc) { s_key[i] = c; s_key[ip] = a; }
            }
            __syncthreads();
        }
    }
    // ---- runs of equal path -> one break point each (a run has at most n_bands - 1 entries). Every thread owns
    //      XB_TOTAL / threads consecutive sorted positions; heads sum their run.
    constexpr int PER = XB_TOTAL / XB_MERGE_THREADS;
    int head_cnt = 0, dsum = 0;
    int hD[PER], hN[PER], hZ[PER];
    uint32_t hP[PER];
    bool hh[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = tid * PER + q;
        hh[q] = false; hD[q] = hN[q] = hZ[q] = 0; hP[q] = 0;
        if (i < S) {
            const uint32_t p = (uint32_t)(s_key[i] >> 32);
            if (i == 0 || (uint32_t)(s_key[i - 1] >> 32) != p) {
                int D = 0, N = 0, Z = 0;
                for (int j = i; j < S && (uint32_t)(s_key[j] >> 32) == p; ++j) {
                    const int src = (int)(uint32_t)s_key[j];
                    D += s_d[src]; N += s_n[src]; Z += s_z[src];
                }
                hh[q] = true; hD[q] = D; hN[q] = N; hZ[q] = Z; hP[q] = p;
                ++head_cnt; dsum += D;
            }
        }
    }
    // block exclusive scan of (head_cnt, dsum)
    const int lane = tid & 31, warp = tid >> 5;
    int ic = head_cnt, id = dsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int oc = __shfl_up_sync(0xFFFFFFFFu, ic, d), od = __shfl_up_sync(0xFFFFFFFFu, id, d);
        if (lane >= d) { ic += oc; id += od; }
    }
    if (lane == 31) { s_wsum[warp] = ic; s_aux[warp] = id; }
    __syncthreads();
    int bc = 0, bd = 0;
    for (int w = 0; w < warp; ++w) { bc += s_wsum[w]; bd += s_aux[w]; }
    int pos = bc + ic - head_cnt, cum = bd + id - dsum;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        if (hh[q]) {
            cum += hD[q];
            tab.path[pos] = hP[q]; tab.cum[pos] = cum; tab.n[pos] = hN[q]; tab.z[pos] = hZ[q];
            ++pos;
        }
    }
    if (tid == XB_MERGE_THREADS - 1) ctr->n_band_bp = pos;
}

// After k_resolve wrote this band's pixels into the root's (peer-mapped) frame buffer.
__global__ void k_band_done(const FrameParams *__restrict__ P, BandPeers peers) {
    __threadfence_system();
    st_release_sys_u64(&peers.box[peers.root]->done[(uint32_t)P->frame_seq & 1u][peers.me], (unsigned long long)(uint32_t)P->frame_seq + 1ull);
}

// The same after a copy-engine push of the band (slpr_band_push): the sequence number travels as an argument.
__global__ void k_band_done_seq(uint32_t seq, int root, BandPeers peers) {
    __threadfence_system();
    st_release_sys_u64(&peers.box[root]->done[seq & 1u][peers.me], (unsigned long long)seq + 1ull);
}

// On the root: every band's pixels of frame seq have landed (bounded spin).
__global__ void k_band_wait_done(uint32_t seq, BandPeers peers, FrameCounters *__restrict__ ctr) {
    const int r = threadIdx.x;
    if (r >= peers.n_bands || r == peers.me) return;
    const unsigned long long *f = &peers.box[peers.me]->done[seq & 1u][r];
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys_u64(f) < (unsigned long long)seq + 1ull) {  // flags only grow
        if (global_timer_ns() - t0 > SLPR_XB_TIMEOUT_NS) { ctr->band_void = 2; break; }
        __nanosleep(500);
    }
}

}  // namespace slpr
