// slpr.cu — context, device arena, frame orchestration and the C ABI of libslpr.so (include/slpr.h).
//
// Host side of the hot path, i.e. what ScanlineVGRasterizer::drawFrame does
// (VkScanlinePR/src/core/scanline/scanline_rasterizer.cpp:282-696), re-designed for CUDA:
//   * one stream, no host round trip inside a frame: the data-dependent sizes (n_fragments,
//     n_records) stay in device memory (FrameCounters) and every downstream kernel is a
//     grid-stride / ticketed persistent kernel that reads them there;
//   * buffers are sized once from a counting pre-pass and grown only if a later frame overflows
//     (detected on the device, handled at the next synchronisation point by re-rendering);
//   * the whole frame is captured once into a CUDA graph and replayed; only the 96-byte
//     FrameParams block (matrix rows, band) is refreshed per frame.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <utility>
#include <string>
#include <vector>

#include "../../include/slpr.h"
#include "geom.cuh"
#include "radix.cuh"
#include "raster.cuh"
#include "scan.cuh"
#include "segsort.cuh"
#include "bands.cuh"
#include "spans.cuh"
#include "walk.cuh"

using namespace slpr;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(SLPR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                          \
    } while (0)

extern "C" const char *slpr_last_error(void) { return g_err.c_str(); }
extern "C" void slpr_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }
extern "C" const char *slpr_version(void) { return "slpr 0.1 (sm_100a)"; }

static int ceil_log2(uint64_t v) {  // bits needed to represent values in [0, v-1]
    int b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__global__ void k_set_params(FrameParams *dst, FrameParams src) { *dst = src; }

// ------------------------------------------------------------------------------------------------
struct slpr_ctx {
    int device = 0;
    uint32_t W = 0, H = 0, flags = 0;  // W x H: the frame the caller sees
    uint32_t ss = 1, iW = 0, iH = 0;   // the pipeline's own resolution: ss x (W, H); ss = 4 with SLPR_FLAG_AA4, else 1
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t side_stream = nullptr;  // the long-piece kernels run beside k_walk (a fork and a join inside the frame, also when captured)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool fork_long = true;             // (measured on B200: 0.025-0.03 ms per frame on the shipped scenes at 4K)
    int num_sms = NUM_SMS_B200;
    int walk_blocks_per_sm = 7, span_blocks_per_sm = 4;  // resident blocks of the persistent kernels (occupancy API)
    int scan_tma_blocks_per_sm = 0;                      // k_scan_tma: CTAs that fit an SM (its tiles are dealt round-robin)

    // scene (slpr_load_scene)
    bool scene_loaded = false;
    uint32_t np = 0, nc = 0, P = 0;
    float2 *d_pos = nullptr;
    uint32_t *d_pos_path = nullptr, *d_cpm = nullptr, *d_ctype = nullptr, *d_cpath = nullptr, *d_frule = nullptr,
             *d_finfo = nullptr;
    float2 *d_tpos = nullptr;
    int *d_pvis = nullptr;
    uint32_t *d_block_cnt = nullptr;  // [WALK_BUCKETS][mono_blocks] pieces per (length bucket, k_monotonize_count block)
    int mono_blocks = 0;
    uint32_t *d_vhist = nullptr;      // [(windows + 1) * WALK_BUCKETS] pieces per virtual bucket (geom.cuh PieceLayout)
    PieceLayout lay{};
    uint32_t *d_live = nullptr;  // [nc] band mode: curves whose path comes near the band (k_band_live)
    float4 *d_pobj = nullptr;    // [P] object-space box of each path's control points (static)
    std::vector<float4> h_pobj;  // its host copy, and the curves' types and paths: slpr_set_curve_weights may have to open a box
    std::vector<uint32_t> h_ctype, h_cpath;
    uint32_t *d_pfc = nullptr;   // [P+1] first curve whose path is >= p (static; [P] = n_curves)
    uint32_t *d_pfp = nullptr;   // [P+1] first point whose path is >= p, when the points are grouped by path (else null)
    uint8_t *d_plive = nullptr;  // [P] band mode: the path can reach the band (k_path_cull)
    uint32_t *d_live_paths = nullptr;  // [P] band mode: the live paths, listed (k_band_paths / k_path_cull) for the sort
    int2 *d_live_range = nullptr;      // [P] band mode: fragment range of each listed path (k_path_segments)
    float *d_cweight = nullptr;  // [nc] SLPR_FLAG_FULL_RVG: weight of every ARC's middle control point (1 until slpr_set_curve_weights)
    float *d_cut = nullptr;
    int *d_count = nullptr, *d_offset = nullptr, *d_seg_tap = nullptr;  // d_seg_tap: [P+1] sort segment table (always built)
    int *d_big = nullptr;              // [P] paths queued for the block-level segmented sort
    bool radix_mode = false;           // false: one-pass segmented sort; true: onesweep radix sort
    bool fill_fused = false;           // true: k_spans marks the cells itself (no k_fill_cells), see choose_fill_mode
    bool long_mode = false;            // true: pieces of 62+ crossings are walked chain by chain (k_long_chains / k_long_emit)
    uint32_t *d_slots = nullptr;       // [5*nc] (length bucket << 26 | rank) of every monotone piece
    PieceRec *d_pieces = nullptr;      // [5*nc] piece records in length-sorted order
    float2 *d_boundary = nullptr;      // [5*nc] first / last emitted parameter of every piece
    uint4 *d_fixlist = nullptr;        // [4*nc] pieces whose predecessor's boundary fragment must be redone (k_walk -> k_piece_fix)

    // frame state
    FrameParams hp{};
    FrameParams *d_params = nullptr;
    FrameCounters *h_ctr = nullptr;  // pinned
    bool have_mvp = false;

    // capacity-dependent buffers
    int cap = 0;
    int2 *d_inter = nullptr;
    uint64_t *d_key[2] = {nullptr, nullptr};
    uint32_t *d_val[2] = {nullptr, nullptr};
    int *d_wn = nullptr;
    int4 *d_rec = nullptr;
    int *d_wsum = nullptr;  // per span tile: delta sum, then exclusive winding prefix
    // taps
    int *t_key32 = nullptr, *t_path = nullptr, *t_wind = nullptr, *t_skey32 = nullptr, *t_sidx = nullptr,
        *t_flags = nullptr, *t_scan3 = nullptr;

    // zero-per-frame temp block: [FrameCounters][tickets][hist][scan status x3][sort lookback]
    unsigned char *d_temp = nullptr;
    size_t temp_bytes = 0, temp_small_bytes = 0;
    FrameCounters *d_ctr = nullptr;
    int *d_tickets = nullptr;  // 3 scan tickets + RS_MAX_PASSES sort tickets + 1 curve-walk work counter
    uint32_t *d_hist = nullptr;
    unsigned long long *d_status[3] = {nullptr, nullptr, nullptr};
    uint32_t *d_lookback = nullptr;
    int sort_tiles_cap = 0;

    KeyLayout L{};
    int key_bits = 0, passes = 0, sorted_buf = 0;

    uint32_t *d_cells = nullptr;
    int cw = 0, ch = 0;
    uint32_t *d_heads = nullptr;       // SLPR_FLAG_BLEND: per-cell list heads, the node pool and its size (grown on demand)
    uint2 *d_nodes = nullptr;
    uint32_t node_cap = 0;
    uint8_t *d_fb = nullptr;
    uint8_t *d_fb2 = nullptr;          // second framebuffer of the pipelined host path (lazy)
    uint8_t *fb_cur = nullptr;         // framebuffer the next frame renders into (d_fb unless pipelining)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rendered[2] = {}, ev_copied[2] = {};
    bool copy_pending[2] = {false, false};
    unsigned pipe_frame = 0;
    // what each of the two pipeline slots holds, so that a frame found invalid after the fact (it outgrew the
    // fragment buffers, or a path outgrew the segmented sort) can be rendered again into the caller's buffer
    struct PipeSlot {
        float rows[16];
        uint8_t *rgba = nullptr;
        size_t stride = 0;
        bool in_flight = false, was_radix = false;
        uint32_t node_cap = 0;       // SLPR_FLAG_BLEND: the node pool this frame was rendered with
        FrameCounters *h = nullptr;  // pinned: this frame's counters, copied right behind its kernels
    } pslot[2];
    uint64_t pipe_redone = 0;          // frames the pipelined path had to render twice
    size_t fb_stride = 0;
    uint8_t *target = nullptr;
    size_t target_stride = 0;

    cudaGraphExec_t gexec = nullptr;   // graph of the frame for `graph_fb`
    cudaGraphExec_t gexec2 = nullptr;  // and for the second framebuffer
    bool graph_valid = false, graph2_valid = false;
    // frames rendered into caller-owned targets (slpr_set_target): one captured graph per target, two kept,
    // so that a caller can alternate between two frame buffers (bench.py --mode bands) without re-capturing
    struct TargetGraph { uint8_t *target = nullptr; size_t stride = 0; cudaGraphExec_t ge = nullptr; bool valid = false; uint64_t used = 0; int launches = 0; } tgraph[2];
    uint64_t tgraph_clock = 0;
    // exact row bands (bands.cuh): caller-owned exchange buffers, scratch, and the graph of the first half
    int *x_sums = nullptr;            // [3][P] this band's per-path winding sums (written by render_band_begin)
    const int *x_gathered = nullptr;  // [n_ranks][3][P] all bands' sums (read by render_band_end)
    int x_ranks = 0, x_rank = 0;
    int *x_d = nullptr, *x_e = nullptr, *x_corr = nullptr;  // [P], [P+1], [2][P]
    unsigned char *x_scan_temp = nullptr;                  // ticket + tile states of the path scan
    size_t x_scan_bytes = 0;
    cudaGraphExec_t gexec_a = nullptr, gexec_s = nullptr;  // fragments | sort (the back half uses the per-target graphs)
    bool graph_a_valid = false, graph_s_valid = false, band_begun = false;
    int launches_a = 0, launches_s = 0;
    cudaEvent_t x_event = nullptr;  // sums and counters of the band in flight are complete
    // exact row bands, device-side sparse exchange (bands.cuh, second half): mailboxes in every band's HBM
    BandMailbox *p_box = nullptr;     // this band's mailbox (cudaMalloc: exportable through CUDA IPC)
    BandPeers peers{};                // n_bands == 0: not configured
    BandEntry *p_list = nullptr;      // [XB_CAP] this band's non-zero per-path sums
    uint32_t *p_tab_path = nullptr;   // [XB_TOTAL] merged break-point table of the frame
    int *p_tab_cum = nullptr, *p_tab_n = nullptr, *p_tab_z = nullptr;
    cudaEvent_t ev_band_rendered[2] = {}, ev_band_pushed[2] = {};  // slpr_band_push: per frame-buffer slot (frame_seq & 1)
    bool band_push_pending[2] = {false, false};
    std::vector<void *> ipc_mapped;   // peer allocations opened with slpr_ipc_import
    std::vector<void *> dev_allocs;   // slpr_alloc_device
    std::vector<std::pair<void *, size_t>> host_allocs;  // slpr_host_alloc
    cudaEvent_t ev[SLPR_STAGE_COUNT + 1] = {};
    bool stage_times_valid = false;
    uint64_t launches = 0;
    int launches_per_frame = 0;
    bool frame_pending = false, frame_done = false;

    // stand-alone primitive temps
    unsigned char *d_prim_temp = nullptr;
    size_t prim_temp_bytes = 0;
};

// The 16-byte draw records (output_buf of gen_merged_fragment_and_span.comp) are written when the caller asked for
// them (SLPR_FLAG_RECORDS, or taps) and whenever coverage is a separate pass over them (small frames).
static bool want_records(const slpr_ctx *c) { return (c->flags & (SLPR_FLAG_RECORDS | SLPR_FLAG_TAPS)) != 0; }

static void invalidate_graphs(slpr_ctx *c) {
    c->graph_valid = c->graph2_valid = c->graph_a_valid = c->graph_s_valid = false;
    for (auto &t : c->tgraph) t.valid = false;
}

static void free_capacity(slpr_ctx *c) {
    cudaFree(c->d_inter); c->d_inter = nullptr;
    for (int i = 0; i < 2; ++i) { cudaFree(c->d_key[i]); cudaFree(c->d_val[i]); c->d_key[i] = nullptr; c->d_val[i] = nullptr; }
    cudaFree(c->d_wn); c->d_wn = nullptr;
    cudaFree(c->d_rec); c->d_rec = nullptr;
    cudaFree(c->d_wsum); c->d_wsum = nullptr;
    cudaFree(c->t_key32); cudaFree(c->t_path); cudaFree(c->t_wind); cudaFree(c->t_skey32); cudaFree(c->t_sidx);
    cudaFree(c->t_flags); cudaFree(c->t_scan3);
    c->t_key32 = c->t_path = c->t_wind = c->t_skey32 = c->t_sidx = c->t_flags = c->t_scan3 = nullptr;
    cudaFree(c->d_temp); c->d_temp = nullptr;
    if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
    if (c->gexec2) { cudaGraphExecDestroy(c->gexec2); c->gexec2 = nullptr; }
    for (auto &t : c->tgraph)
        if (t.ge) { cudaGraphExecDestroy(t.ge); t.ge = nullptr; }
    invalidate_graphs(c);
    c->cap = 0;
}

static void free_exchange(slpr_ctx *c) {
    cudaFree(c->x_d); cudaFree(c->x_e); cudaFree(c->x_corr); cudaFree(c->x_scan_temp);
    c->x_d = c->x_e = c->x_corr = nullptr;
    c->x_scan_temp = nullptr;
    c->x_sums = nullptr; c->x_gathered = nullptr;
    c->x_ranks = c->x_rank = 0;
    c->band_begun = false;
    if (c->gexec_a) { cudaGraphExecDestroy(c->gexec_a); c->gexec_a = nullptr; }
    if (c->gexec_s) { cudaGraphExecDestroy(c->gexec_s); c->gexec_s = nullptr; }
    if (c->x_event) { cudaEventDestroy(c->x_event); c->x_event = nullptr; }
}

static void free_scene(slpr_ctx *c) {
    free_exchange(c);
    cudaFree(c->d_pos); cudaFree(c->d_pos_path); cudaFree(c->d_cpm); cudaFree(c->d_ctype); cudaFree(c->d_cpath);
    cudaFree(c->d_frule); cudaFree(c->d_finfo); cudaFree(c->d_tpos); cudaFree(c->d_pvis); cudaFree(c->d_pobj); cudaFree(c->d_pfc); cudaFree(c->d_pfp); cudaFree(c->d_plive); cudaFree(c->d_live_paths); c->d_live_paths = nullptr; cudaFree(c->d_live_range); c->d_live_range = nullptr; cudaFree(c->d_live); cudaFree(c->d_block_cnt); cudaFree(c->d_vhist); cudaFree(c->d_cut); cudaFree(c->d_cweight); c->d_cweight = nullptr;
    cudaFree(c->d_count); cudaFree(c->d_offset); cudaFree(c->d_seg_tap);
    cudaFree(c->d_big); c->d_big = nullptr;
    cudaFree(c->d_slots); cudaFree(c->d_pieces); cudaFree(c->d_boundary); cudaFree(c->d_fixlist);
    c->d_slots = nullptr; c->d_pieces = nullptr; c->d_boundary = nullptr; c->d_fixlist = nullptr;
    c->d_pos = nullptr; c->d_pos_path = c->d_cpm = c->d_ctype = c->d_cpath = c->d_frule = c->d_finfo = nullptr;
    c->d_tpos = nullptr; c->d_pvis = nullptr; c->d_pobj = nullptr; c->d_pfc = nullptr; c->d_pfp = nullptr; c->d_plive = nullptr; c->d_live = nullptr; c->d_block_cnt = nullptr; c->d_vhist = nullptr; c->d_cut = nullptr; c->d_count = c->d_offset = c->d_seg_tap = nullptr;
    c->scene_loaded = false;
}

static int alloc_capacity(slpr_ctx *c, int cap) {
    free_capacity(c);
    cap = (int)align_up((size_t)std::max(cap, 1 << 14), 4096);
    const size_t n = (size_t)cap;
    for (int i = 0; i < 2; ++i) {
        CU(cudaMalloc(&c->d_key[i], n * 8));
        CU(cudaMalloc(&c->d_val[i], n * 4));
        // never hand uninitialised pairs to a kernel: a frame that is abandoned half way (overflow, sort fall-back)
        // leaves these partly unwritten, and the keys index per-path tables
        CU(cudaMemsetAsync(c->d_key[i], 0, n * 8, c->stream));
        CU(cudaMemsetAsync(c->d_val[i], 0, n * 4, c->stream));
    }
    // draw records (32 B per fragment of capacity): allocated by ensure_records() when a frame will write them
    CU(cudaMalloc(&c->d_wsum, (n / SP_TILE + 4) * sizeof(int)));
    if (c->flags & SLPR_FLAG_TAPS) {
        CU(cudaMalloc(&c->d_wn, (n + 4) * 4));
        CU(cudaMalloc(&c->d_inter, (n + 1) * sizeof(int2)));
        CU(cudaMalloc(&c->t_key32, (n + 1) * 4));
        CU(cudaMalloc(&c->t_path, n * 4));
        CU(cudaMalloc(&c->t_wind, n * 4));
        CU(cudaMalloc(&c->t_skey32, n * 4));
        CU(cudaMalloc(&c->t_sidx, n * 4));
        CU(cudaMalloc(&c->t_flags, 2 * n * 4));
        CU(cudaMalloc(&c->t_scan3, (2 * n + 1) * 4));
    }
    // per-frame zeroed temp block
    constexpr size_t frag_tile = SCAN_TILE_MIN < SP_TILE ? SCAN_TILE_MIN : SP_TILE;  // status arrays 1 and 2 also serve k_spans' tiles
    const size_t scan_tiles[3] = {(size_t)c->nc / SCAN_TILE_MIN + 2, n / frag_tile + 2, n / frag_tile + 2};
    c->sort_tiles_cap = (int)(n / RS_TILE + 2);
    size_t off = 0;
    const size_t o_ctr = off; off += align_up(sizeof(FrameCounters), 256);
    const size_t o_tick = off; off += align_up((3 + RS_MAX_PASSES + 1) * sizeof(int), 256);
    const size_t o_hist = off; off += align_up((size_t)RS_MAX_PASSES * RS_BINS * 4, 256);
    size_t o_status[3];
    for (int i = 0; i < 3; ++i) { o_status[i] = off; off += align_up(scan_tiles[i] * 8, 256); }
    c->temp_small_bytes = off;
    const size_t o_lb = off; off += align_up((size_t)c->passes * c->sort_tiles_cap * RS_BINS * 4, 256);
    c->temp_bytes = off;
    CU(cudaMalloc(&c->d_temp, c->temp_bytes));
    c->d_ctr = reinterpret_cast<FrameCounters *>(c->d_temp + o_ctr);
    c->d_tickets = reinterpret_cast<int *>(c->d_temp + o_tick);
    c->d_hist = reinterpret_cast<uint32_t *>(c->d_temp + o_hist);
    for (int i = 0; i < 3; ++i) c->d_status[i] = reinterpret_cast<unsigned long long *>(c->d_temp + o_status[i]);
    c->d_lookback = reinterpret_cast<uint32_t *>(c->d_temp + o_lb);
    c->cap = cap;
    return SLPR_OK;
}

// ------------------------------------------------------------------------------------------------
// Device -> host copy of a frame: one linear copy when both sides are tightly packed (the usual case), else 2-D.
static cudaError_t copy_frame_to_host(const slpr_ctx *c, uint8_t *dst, size_t dst_stride, const uint8_t *src, size_t src_stride, cudaStream_t s) {
    const size_t row = (size_t)c->W * 4;
    if (dst_stride == row && src_stride == row) return cudaMemcpyAsync(dst, src, row * c->H, cudaMemcpyDeviceToHost, s);
    return cudaMemcpy2DAsync(dst, dst_stride, src, src_stride, row, c->H, cudaMemcpyDeviceToHost, s);
}

extern "C" slpr_ctx *slpr_create(int device, uint32_t width, uint32_t height, uint32_t flags) {
    if (width == 0 || height == 0 || width > 32766 || height > 32766) {
        fail(SLPR_ERR_INVALID, "slpr_create: width/height must be in [1, 32766] (16-bit key fields, SURVEY D.6)");
        return nullptr;
    }
    if ((flags & SLPR_FLAG_AA4) && (width > 8191 || height > 8191)) {
        fail(SLPR_ERR_INVALID, "slpr_create: with SLPR_FLAG_AA4 width/height must be at most 8191 (the pipeline runs at four times the size)");
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fail(SLPR_ERR_CUDA, "slpr_create: no usable CUDA device (%s); this library has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) { fail(SLPR_ERR_INVALID, "slpr_create: device %d out of range [0,%d)", device, ndev); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail(SLPR_ERR_CUDA, "cudaSetDevice(%d) failed", device); return nullptr; }
    slpr_ctx *c = new slpr_ctx();
    c->device = device; c->W = width; c->H = height; c->flags = flags;
    c->ss = (flags & SLPR_FLAG_AA4) ? 4u : 1u;
    c->iW = width * c->ss; c->iH = height * c->ss;
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    c->stream = c->own_stream;
    ok = ok && cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_params, sizeof(FrameParams)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_ctr, sizeof(FrameCounters)) == cudaSuccess;
    c->cw = (int)(c->iW + 1) / 2; c->ch = (int)(c->iH + 1) / 2;
    ok = ok && cudaMalloc(&c->d_cells, (size_t)c->cw * c->ch * 4) == cudaSuccess;
    ok = ok && cudaMemset(c->d_cells, 0, (size_t)c->cw * c->ch * 4) == cudaSuccess;
    if (flags & SLPR_FLAG_BLEND) {
        c->node_cap = (uint32_t)std::max<size_t>((size_t)1 << 20, (size_t)c->cw * c->ch);
        ok = ok && cudaMalloc(&c->d_heads, (size_t)c->cw * c->ch * 4) == cudaSuccess;
        ok = ok && cudaMemset(c->d_heads, 0, (size_t)c->cw * c->ch * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_nodes, (size_t)c->node_cap * sizeof(uint2)) == cudaSuccess;
    }
    c->fb_stride = (size_t)width * 4;
    ok = ok && cudaMalloc(&c->d_fb, c->fb_stride * height) == cudaSuccess;
    c->fb_cur = c->d_fb;
    for (auto &ev : c->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM_BYTES) == cudaSuccess;
    {
        const bool full = (flags & SLPR_FLAG_FULL_RVG) != 0, fma = (flags & SLPR_FLAG_CONTRACT_FMA) != 0;
        auto walk = full ? (fma ? k_walk<true, true> : k_walk<true, false>) : (fma ? k_walk<false, true> : k_walk<false, false>);
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->walk_blocks_per_sm, walk, WALK_THREADS, 0) == cudaSuccess;
    }
    ok = ok && cudaFuncSetAttribute(k_scan_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->scan_tma_blocks_per_sm, k_scan_tma, ST_THREADS, ST_SMEM_BYTES) == cudaSuccess;
    for (auto fn : {k_spans<false, true, false>, k_spans<true, true, false>, k_spans<true, false, false>, k_spans<false, true, true>, k_spans<true, true, true>,
                    k_spans<true, true, false, true>, k_spans<true, false, false, true>, k_spans<true, true, true, true>})
        ok = ok && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_STAGE_BYTES) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->span_blocks_per_sm, k_spans<true, false, false>, SP_THREADS, SP_STAGE_BYTES) == cudaSuccess;
    if (!ok) {
        fail(SLPR_ERR_CUDA, "slpr_create: device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        slpr_destroy(c);
        return nullptr;
    }
    memset(&c->hp, 0, sizeof c->hp);
    const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(c->hp.rows, ident, sizeof ident);
    c->hp.width = (int)c->iW; c->hp.height = (int)c->iH;
    c->hp.band_y0 = 0; c->hp.band_y1 = (int)c->iH; c->hp.cull = 0;
    return c;
}

extern "C" void slpr_destroy(slpr_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void *p : c->ipc_mapped) cudaIpcCloseMemHandle(p);
    for (void *p : c->dev_allocs) cudaFree(p);
    for (auto &h : c->host_allocs) { cudaHostUnregister(h.first); munmap(h.first, h.second); }
    cudaFree(c->p_box); cudaFree(c->p_list); cudaFree(c->p_tab_path); cudaFree(c->p_tab_cum); cudaFree(c->p_tab_n); cudaFree(c->p_tab_z);
    free_capacity(c);
    free_scene(c);
    cudaFree(c->d_params); cudaFreeHost(c->h_ctr); cudaFreeHost(c->pslot[0].h); cudaFreeHost(c->pslot[1].h); cudaFree(c->d_cells); cudaFree(c->d_heads); cudaFree(c->d_nodes); cudaFree(c->d_fb); cudaFree(c->d_fb2); cudaFree(c->d_prim_temp);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (int i = 0; i < 2; ++i) { if (c->ev_rendered[i]) cudaEventDestroy(c->ev_rendered[i]); if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); }
    for (int i = 0; i < 2; ++i) { if (c->ev_band_rendered[i]) cudaEventDestroy(c->ev_band_rendered[i]); if (c->ev_band_pushed[i]) cudaEventDestroy(c->ev_band_pushed[i]); }
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
}

extern "C" int slpr_set_stream(slpr_ctx *c, void *s) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = s ? reinterpret_cast<cudaStream_t>(s) : c->own_stream;
    return SLPR_OK;
}

template <class T>
static int upload(T **dst, const void *src, size_t count) {
    CU(cudaMalloc(dst, std::max<size_t>(count, 1) * sizeof(T)));
    if (count) CU(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return SLPR_OK;
}

extern "C" int slpr_load_scene(slpr_ctx *c, const float *pos_xy, const uint32_t *pos_path, uint32_t n_points,
                               const uint32_t *curve_pos_map, const uint32_t *curve_type, const uint32_t *curve_path,
                               uint32_t n_curves, const uint32_t *fill_rule, const uint32_t *fill_rgba8, uint32_t n_paths) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if ((n_points && (!pos_xy || !pos_path)) || (n_curves && (!curve_pos_map || !curve_type || !curve_path)) ||
        (n_paths && (!fill_rule || !fill_rgba8)))
        return fail(SLPR_ERR_INVALID, "slpr_load_scene: null array with non-zero count");
    if (n_points >= (1u << 30) || n_curves >= (1u << 30) || n_paths >= (1u << 30))
        return fail(SLPR_ERR_INVALID, "slpr_load_scene: counts must be < 2^30");
    // validate indices on the host once: the kernels trust them
    for (uint32_t i = 0; i < n_curves; ++i) {
        const uint32_t npts = curve_type[i] & 7u;
        if (curve_path[i] >= n_paths) return fail(SLPR_ERR_INVALID, "slpr_load_scene: curve %u has path %u >= n_paths", i, curve_path[i]);
        if ((uint64_t)curve_pos_map[i] + std::max(npts, 1u) > (uint64_t)n_points && npts)
            return fail(SLPR_ERR_INVALID, "slpr_load_scene: curve %u points [%u,+%u) exceed n_points", i, curve_pos_map[i], npts);
        if (i && curve_path[i] < curve_path[i - 1]) return fail(SLPR_ERR_INVALID, "slpr_load_scene: curve_path must be non-decreasing");
    }
    for (uint32_t i = 0; i < n_paths; ++i)
        if (fill_rule[i] > 1u) return fail(SLPR_ERR_INVALID, "slpr_load_scene: path %u has fill rule %u (0 = nonzero, 1 = even-odd)", i, fill_rule[i]);
    for (uint32_t i = 0; i < n_points; ++i)
        if (pos_path[i] >= n_paths) return fail(SLPR_ERR_INVALID, "slpr_load_scene: point %u has path %u >= n_paths", i, pos_path[i]);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CU(cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; ++i) { c->pslot[i].in_flight = false; c->copy_pending[i] = false; }  // frames of the old scene are not looked at again
    free_capacity(c);
    free_scene(c);
    c->np = n_points; c->nc = n_curves; c->P = n_paths;
    int rc;
    if ((rc = upload(&c->d_pos, pos_xy, n_points))) return rc;
    if ((rc = upload(&c->d_pos_path, pos_path, n_points))) return rc;
    if ((rc = upload(&c->d_cpm, curve_pos_map, n_curves))) return rc;
    if ((rc = upload(&c->d_ctype, curve_type, n_curves))) return rc;
    if ((rc = upload(&c->d_cpath, curve_path, n_curves))) return rc;
    if ((rc = upload(&c->d_frule, fill_rule, n_paths))) return rc;
    if ((rc = upload(&c->d_finfo, fill_rgba8, n_paths))) return rc;
    CU(cudaMalloc(&c->d_tpos, std::max<size_t>(n_points, 1) * sizeof(float2)));
    CU(cudaMalloc(&c->d_pvis, std::max<size_t>(n_paths, 1) * 4));
    {   // object-space box of every path (band mode culls whole paths with it)
        std::vector<float4> box(std::max<size_t>(n_paths, 1), make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f));
        for (uint32_t i = 0; i < n_points; ++i) {
            float4 &b = box[pos_path[i]];
            const float x = pos_xy[2 * i], y = pos_xy[2 * i + 1];
            if (!(x == x) || !(y == y)) { b = make_float4(-3.0e38f, -3.0e38f, 3.0e38f, 3.0e38f); continue; }  // NaN: never culled
            b.x = std::min(b.x, x); b.y = std::min(b.y, y); b.z = std::max(b.z, x); b.w = std::max(b.w, y);
        }
        // A curve type the shaders have no arm for is evaluated as the point (0, 0) whatever its control points are
        // (gen_fragment.comp:132-156 loads nothing; ARC too without SLPR_FLAG_FULL_RVG), so its fragments are not where its
        // box is: such a path is never culled (found by tools/fuzz_parity.py: the fragment at the origin went missing).
        const bool full_rvg = (c->flags & SLPR_FLAG_FULL_RVG) != 0;
        for (uint32_t i = 0; i < n_curves; ++i) {
            const uint32_t t = curve_type[i];
            if (!(t == T_LINE || t == T_CUBIC || t == T_QUADRIC || (full_rvg && t == T_ARC)))
                box[curve_path[i]] = make_float4(-3.0e38f, -3.0e38f, 3.0e38f, 3.0e38f);
        }
        c->h_pobj = box;
        c->h_ctype.assign(curve_type, curve_type + n_curves);
        c->h_cpath.assign(curve_path, curve_path + n_curves);
        // first curve of every path (curve_path is non-decreasing): the per-frame segment table is offsets[] read through it
        std::vector<uint32_t> pfc((size_t)n_paths + 1);
        uint32_t cur = 0;
        for (uint32_t p = 0; p <= n_paths; ++p) {
            while (cur < n_curves && curve_path[cur] < p) ++cur;
            pfc[p] = cur;
        }
        if ((rc = upload(&c->d_pfc, pfc.data(), pfc.size()))) return rc;
        // and, when the points are grouped by path (loadVG's flattening always does), its first point: band mode then
        // only touches the paths that reach the band (k_band_paths)
        bool grouped = true;
        for (uint32_t i = 1; i < n_points && grouped; ++i) grouped = pos_path[i] >= pos_path[i - 1];
        if (grouped) {
            std::vector<uint32_t> pfp((size_t)n_paths + 1);
            uint32_t pt = 0;
            for (uint32_t p = 0; p <= n_paths; ++p) {
                while (pt < n_points && pos_path[pt] < p) ++pt;
                pfp[p] = pt;
            }
            if ((rc = upload(&c->d_pfp, pfp.data(), pfp.size()))) return rc;
        }
        CU(cudaMalloc(&c->d_pobj, box.size() * sizeof(float4)));
        CU(cudaMemcpy(c->d_pobj, box.data(), box.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_plive, std::max<size_t>(n_paths, 1)));
        CU(cudaMalloc(&c->d_live_paths, std::max<size_t>(n_paths, 1) * 4));
        CU(cudaMalloc(&c->d_live_range, std::max<size_t>(n_paths, 1) * sizeof(int2)));
    }
    CU(cudaMalloc(&c->d_live, std::max<size_t>(n_curves, 1) * 4));
    c->mono_blocks = (int)std::max<long long>(1, std::min<long long>(((long long)n_curves + 255) / 256, (long long)c->num_sms * 8));
    CU(cudaMalloc(&c->d_block_cnt, (size_t)WALK_BUCKETS * c->mono_blocks * 4));
    c->lay.n_blocks = (uint32_t)c->mono_blocks;
    c->lay.n_windows = (uint32_t)std::max<long long>(1, std::min<long long>(((long long)n_curves + WALK_WINDOW_CURVES - 1) / WALK_WINDOW_CURVES,
                                                                               std::min<long long>(WALK_MAX_WINDOWS, c->mono_blocks)));
    c->lay.long_min = WALK_LONG;
    if (c->flags & SLPR_FLAG_WINDOWED_WALK) c->lay.n_windows = (uint32_t)std::min(8, c->mono_blocks);
    else if ((long long)n_curves <= WALK_WINDOWED_CURVES) { c->lay.n_windows = 1; c->lay.long_min = 0; }  // one longest-first order
    c->lay.blocks_per_window = (c->lay.n_blocks + c->lay.n_windows - 1) / c->lay.n_windows;
    CU(cudaMalloc(&c->d_vhist, (size_t)WALK_VBUCKETS_MAX * 4));
    CU(cudaMalloc(&c->d_cut, std::max<size_t>(n_curves, 1) * 5 * 4));
    if (c->flags & SLPR_FLAG_FULL_RVG) {
        std::vector<float> ones(std::max<size_t>(n_curves, 1), 1.0f);
        if ((rc = upload(&c->d_cweight, ones.data(), ones.size()))) return rc;
    }
    CU(cudaMalloc(&c->d_count, ((size_t)n_curves + 4) * 4));
    CU(cudaMalloc(&c->d_offset, ((size_t)n_curves + 4) * 4));
    CU(cudaMalloc(&c->d_slots, std::max<size_t>(n_curves, 1) * 5 * 4));
    CU(cudaMalloc(&c->d_pieces, std::max<size_t>(n_curves, 1) * 5 * sizeof(PieceRec)));
    CU(cudaMalloc(&c->d_boundary, std::max<size_t>(n_curves, 1) * 5 * sizeof(float2)));
    CU(cudaMalloc(&c->d_fixlist, std::max<size_t>(n_curves, 1) * 4 * sizeof(uint4)));  // a curve has at most 4 pieces after its first
    CU(cudaMalloc(&c->d_seg_tap, ((size_t)n_paths + 1) * 4));
    CU(cudaMalloc(&c->d_big, std::max<size_t>(n_paths, 1) * 4));
    c->radix_mode = (c->flags & SLPR_FLAG_RADIX_SORT) != 0;
    // compact key geometry: x cell in [0,(W'+4)/2], row rank in [0,ny], path in [0,P)
    const int Wp = (int)(c->iW & ~1u);
    c->L.ny = (int)(c->iH + 1) / 2;
    c->L.bits_x = std::max(1, ceil_log2((uint64_t)(Wp + 4) / 2 + 1));
    c->L.bits_y = std::max(1, ceil_log2((uint64_t)c->L.ny + 1));
    c->L.bits_path = ceil_log2(std::max<uint64_t>(n_paths, 1));
    c->key_bits = c->L.bits_x + c->L.bits_y + c->L.bits_path;
    c->passes = std::max(1, (c->key_bits + 7) / 8);
    c->scene_loaded = true;
    c->frame_done = c->frame_pending = false;
    return SLPR_OK;
}

extern "C" int slpr_set_curve_weights(slpr_ctx *c, const float *curve_weight, uint32_t n_curves) {
    if (!c || !curve_weight) return fail(SLPR_ERR_INVALID, "slpr_set_curve_weights: null argument");
    if (!c->scene_loaded || n_curves != c->nc) return fail(SLPR_ERR_STATE, "slpr_set_curve_weights: load the scene first; n_curves must match it");
    if (!(c->flags & SLPR_FLAG_FULL_RVG)) return fail(SLPR_ERR_STATE, "slpr_set_curve_weights: the context was not created with SLPR_FLAG_FULL_RVG");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (n_curves) CU(cudaMemcpy(c->d_cweight, curve_weight, (size_t)n_curves * 4, cudaMemcpyHostToDevice));
    // a rational arc stays inside its control points' hull only with a non-negative weight: otherwise the band front
    // end must not cull its path by the box (geom.cuh does the same per curve)
    bool opened = false;
    for (uint32_t i = 0; i < n_curves; ++i)
        if (c->h_ctype[i] == T_ARC && !(curve_weight[i] > 0.0f)) {
            c->h_pobj[c->h_cpath[i]] = make_float4(-3.0e38f, -3.0e38f, 3.0e38f, 3.0e38f);
            opened = true;
        }
    if (opened) CU(cudaMemcpy(c->d_pobj, c->h_pobj.data(), c->h_pobj.size() * sizeof(float4), cudaMemcpyHostToDevice));
    return SLPR_OK;
}

extern "C" int slpr_set_mvp(slpr_ctx *c, const float rows[16]) {
    if (!c || !rows) return fail(SLPR_ERR_INVALID, "slpr_set_mvp: null argument");
    memcpy(c->hp.rows, rows, 16 * sizeof(float));
    if (c->ss != 1)  // supersampled: x and y rows scaled by a power of two — exact, so sample positions are ss x the plain frame's
        for (int i = 0; i < 8; ++i) c->hp.rows[i] *= (float)c->ss;
    c->have_mvp = true;
    return SLPR_OK;
}

extern "C" int slpr_set_band(slpr_ctx *c, uint32_t y0, uint32_t y1) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (y0 >= y1 || y1 > c->H || (y0 & 1) || ((y1 & 1) && y1 != c->H))
        return fail(SLPR_ERR_INVALID, "slpr_set_band: need even 0 <= y_begin < y_end <= height (got %u,%u)", y0, y1);
    const int cull = (y0 != 0 || y1 != c->H) ? 1 : 0;
    if (cull != c->hp.cull) invalidate_graphs(c);  // the band-mode kernels take one more table (path row boxes)
    c->hp.band_y0 = (int)(y0 * c->ss); c->hp.band_y1 = (int)(y1 * c->ss);
    c->hp.cull = cull;
    return SLPR_OK;
}

extern "C" int slpr_set_target(slpr_ctx *c, void *dev_rgba, size_t stride_bytes) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (dev_rgba && stride_bytes < (size_t)c->W * 4) return fail(SLPR_ERR_INVALID, "slpr_set_target: stride smaller than a row");
    c->target = reinterpret_cast<uint8_t *>(dev_rgba);
    c->target_stride = stride_bytes;
    return SLPR_OK;
}

// ------------------------------------------------------------------------------------------------
// Before a frame is enqueued (never inside a stream capture): the record buffer, if this frame writes records.
static int ensure_records(slpr_ctx *c) {
    if (c->d_rec || (c->fill_fused && !want_records(c))) return SLPR_OK;
    CU(cudaMalloc(&c->d_rec, (2 * (size_t)c->cap + 1) * sizeof(int4)));
    return SLPR_OK;
}

static int grid_for(const slpr_ctx *c, long long work_items, int threads, int per_sm) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)c->num_sms * per_sm;
    return (int)std::max(1ll, std::min(blocks, cap));
}

// Counting prefix of the frame: transform, monotonize+count, scan #1. `timed` records stage events.
static int enqueue_count_phase(slpr_ctx *c, cudaStream_t s, bool timed, int &launches) {
    CU(cudaMemsetAsync(c->d_temp, 0, c->temp_bytes, s));
    CU(cudaMemsetAsync(c->d_pvis, 0, std::max<size_t>(c->P, 1) * 4, s));
    if (timed) CU(cudaEventRecord(c->ev[0], s));
    const bool band = c->hp.cull != 0;
    const bool fma = (c->flags & SLPR_FLAG_CONTRACT_FMA) != 0;
    LiveCurves live{nullptr, c->d_ctr};
    if (band && c->d_pfp) {  // one pass over the paths: cull, transform the live ones, list their curves
        CU(cudaMemsetAsync(c->d_count, 0, (size_t)c->nc * 4, s));  // k_monotonize_count only writes the live curves
        (fma ? k_band_paths<true> : k_band_paths<false>)<<<grid_for(c, c->P, 256, 8), 256, 0, s>>>(c->d_params, c->P, c->d_pobj, c->d_pfp, c->d_pfc, c->d_pos, c->d_tpos, c->d_pvis,
                                                                c->d_live, c->d_ctr, c->d_live_paths);
        ++launches;
        if (timed) CU(cudaEventRecord(c->ev[1], s));
        live.list = c->d_live;
    } else {
        const uint8_t *plive = band ? c->d_plive : nullptr;
        if (plive) {
            k_path_cull<<<grid_for(c, c->P, 256, 8), 256, 0, s>>>(c->d_params, c->P, c->d_pobj, c->d_plive, c->d_live_paths, c->d_ctr);
            ++launches;
        }
        (fma ? k_transform<true> : k_transform<false>)<<<grid_for(c, c->np, 256, 8), 256, 0, s>>>(c->d_params, c->np, c->d_pos, c->d_pos_path, c->d_tpos, c->d_pvis, plive);
        ++launches;
        if (timed) CU(cudaEventRecord(c->ev[1], s));
        if (plive) {
            k_band_live<<<grid_for(c, ((long long)c->nc + LIVE_CHUNK - 1) / LIVE_CHUNK, 256, 8), 256, 0, s>>>(c->nc, c->d_cpath, plive, c->d_count,
                                                                                                         c->d_live, c->d_ctr);
            ++launches;
            live.list = c->d_live;
        }
    }
    (fma ? k_monotonize_count<true> : k_monotonize_count<false>)<<<c->mono_blocks, 256, 0, s>>>(c->d_params, c->nc, c->d_ctype, c->d_cpm, c->d_cpath,
                                                                   c->d_tpos, c->d_pvis, c->d_cut, c->d_count, c->d_slots,
                                                                   c->d_block_cnt, live, c->lay, FullRvg{c->d_cweight});
    k_bucket_scan<<<dim3(WALK_BUCKETS, c->lay.n_windows), c->lay.blocks_per_window > 256 ? 1024 : 128, 0, s>>>(c->d_block_cnt, c->lay, c->d_vhist);
    launches += 2;
    if (timed) CU(cudaEventRecord(c->ev[2], s));
    ScanI32Op op1{c->d_count, c->d_offset, (long long)c->nc, &c->d_ctr->n_fragments, c->cap, &c->d_ctr->overflow};
    k_lookback_scan<ScanI32Op><<<grid_for(c, (long long)c->nc / 16 + 1, SCAN_THREADS, ScanI32Op::MIN_BLOCKS), SCAN_THREADS, 0, s>>>(
        op1, ScanTemp{c->d_status[0], c->d_tickets + 0});
    ++launches;
    if (timed) CU(cudaEventRecord(c->ev[3], s));
    return SLPR_OK;
}

// The frame in three pieces: enqueue_fragments (transform .. fragment generation; with a band exchange
// configured it also leaves the per-path winding sums in x_sums), enqueue_sort, enqueue_back.
static int enqueue_fragments(slpr_ctx *c, cudaStream_t s, bool timed, int &launches) {
    int rc = enqueue_count_phase(c, s, timed, launches);
    if (rc) return rc;
    FragTaps ft{c->t_key32, c->t_path, c->t_wind};
    const bool fma = (c->flags & SLPR_FLAG_CONTRACT_FMA) != 0, full = c->d_cweight != nullptr;
    (fma ? k_piece_emit<true> : k_piece_emit<false>)<<<c->mono_blocks, 256, 0, s>>>(c->d_params, c->nc, c->d_ctype, c->d_cpm, c->d_cpath, c->d_frule, c->d_tpos, c->d_cut,
                                                             c->d_offset, c->d_slots, c->d_ctr, c->cap,
                                                             PieceRanks{c->d_block_cnt, c->d_vhist, c->lay},
                                                             LiveCurves{c->hp.cull ? c->d_live : nullptr, c->d_ctr}, c->d_pieces, FullRvg{c->d_cweight});
    if (timed) CU(cudaEventRecord(c->ev[4], s));
    if (c->long_mode) {
        // few, long pieces (small scenes at large frames): two independent chains per piece, then a parallel emit — on a
        // side stream, beside k_walk, which takes the other pieces: the longest chain is what such a frame waits for
        LongScratch ls{c->d_val[1], reinterpret_cast<uint32_t *>(c->d_key[1])};
        const int lgrid = c->num_sms * 8;
        auto chains = full ? (fma ? k_long_chains<true, true> : k_long_chains<true, false>) : (fma ? k_long_chains<false, true> : k_long_chains<false, false>);
        auto emit = full ? (fma ? k_long_emit<true, true> : k_long_emit<true, false>) : (fma ? k_long_emit<false, true> : k_long_emit<false, false>);
        cudaStream_t ls_ = c->fork_long ? c->side_stream : s;
        if (c->fork_long) { CU(cudaEventRecord(c->ev_fork, s)); CU(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0)); }
        chains<<<lgrid, 128, 0, ls_>>>(c->d_pieces, c->d_ctr, c->cap, ls);
        emit<<<c->num_sms * 4, LONG_EMIT_THREADS, 0, ls_>>>(c->d_params, c->d_pieces, c->d_ctr, c->cap, ls, c->L, c->d_key[0], c->d_val[0], ft, c->d_inter, c->d_boundary, c->d_fixlist);
        if (c->fork_long) CU(cudaEventRecord(c->ev_join, c->side_stream));
        launches += 2;
    }
    auto walk = full ? (fma ? k_walk<true, true> : k_walk<true, false>) : (fma ? k_walk<false, true> : k_walk<false, false>);
    walk<<<c->num_sms * std::max(1, c->walk_blocks_per_sm), WALK_THREADS, 0, s>>>(
        c->d_params, c->d_pieces, c->d_ctr, c->cap, WalkTemp{c->d_tickets + 3 + RS_MAX_PASSES},
        c->L, c->d_key[0], c->d_val[0], ft, c->d_inter, c->d_boundary, c->d_fixlist, c->long_mode ? 1 : 0);
    launches += 1;
    if (c->long_mode && c->fork_long) CU(cudaStreamWaitEvent(s, c->ev_join, 0));
    (fma ? k_piece_fix<true> : k_piece_fix<false>)<<<8, 256, 0, s>>>(c->d_params, c->d_ctype, c->d_cpm, c->d_cpath, c->d_frule, c->d_tpos, c->d_ctr, c->cap, c->d_boundary,
                                  c->d_fixlist, c->L, c->d_key[0], c->d_val[0], ft, FullRvg{c->d_cweight});
    launches += 2;
    if (timed) CU(cudaEventRecord(c->ev[5], s));
    k_path_segments<<<grid_for(c, (long long)c->P + 1, 256, 4), 256, 0, s>>>(c->d_pfc, c->P, c->d_offset, c->d_seg_tap, c->d_ctr, c->cap,
                                                                             c->hp.cull ? c->d_live_paths : nullptr, c->d_live_range);
    ++launches;
    if (c->x_sums) {  // before the sort: the radix sort reuses buffer 0
        k_band_sums<<<grid_for(c, (long long)c->P * 32, 256, 8), 256, 0, s>>>(c->d_seg_tap, c->P, c->d_key[0], c->d_val[0], c->d_ctr,
                                                                             c->cap, c->L, c->x_sums);
        ++launches;
    } else if (c->peers.n_bands > 0 && (c->radix_mode || !SLPR_SEG_FUSE_SUMS)) {
        // sparse exchange, radix sort: the paths with a residue from a pass of its own (before the sort, which reuses
        // buffer 0); with the segmented sort k_segsort_warp forms the sums while it has the path in hand
        k_band_sums_sparse<<<grid_for(c, (long long)c->P * 32, 256, 8), 256, 0, s>>>(c->d_seg_tap, c->P, c->d_key[0], c->d_val[0], c->d_ctr,
                                                                                    c->cap, c->L, c->p_list, c->hp.cull ? c->d_live_paths : nullptr);
        ++launches;
    }
    CU(cudaGetLastError());
    return SLPR_OK;
}

static int enqueue_sort(slpr_ctx *c, cudaStream_t s, bool timed, int &launches) {
    int cur = 0;
    if (!c->radix_mode) {  // every path sorted on chip, one read + one write of the pairs (segsort.cuh)
        if (timed) CU(cudaEventRecord(c->ev[6], s));
        const int yx_bits = c->L.bits_x + c->L.bits_y;
        const bool peer_band = c->peers.n_bands > 0 && !c->x_sums;
        SegBand sb{c->d_tickets + 3, c->hp.cull ? c->d_live_paths : nullptr, c->d_live_range, (peer_band && SLPR_SEG_FUSE_SUMS) ? c->p_list : nullptr, SegGeo{c->L.bits_x, c->L.bits_y, c->L.ny}};
        auto segsort = (sb.live_paths || sb.sums) ? k_segsort_warp<true> : k_segsort_warp<false>;
        segsort<<<grid_for(c, ((long long)c->P + SEG_CHUNK - 1) / SEG_CHUNK * 32, 256, 8), 256, 0, s>>>(
            c->d_seg_tap, c->P, c->d_key[0], c->d_val[0], c->d_key[1], c->d_val[1], c->d_ctr, c->cap, yx_bits, c->d_big, sb);
        k_segsort_block<<<c->num_sms * 2, SEG_BLOCK_THREADS, 0, s>>>(c->d_seg_tap, c->d_key[0], c->d_val[0], c->d_key[1], c->d_val[1],
                                                                     c->d_ctr, c->cap, yx_bits, c->d_big);
        launches += 2;
        cur = 1;
    } else {
    SortCount cnt{&c->d_ctr->n_fragments, 0, c->cap};
    SortTemp st{c->d_hist, c->d_lookback, c->d_tickets + 3, c->sort_tiles_cap};
    k_radix_hist<<<c->num_sms * 4, RH_THREADS, 0, s>>>(c->d_key[0], cnt, c->passes, c->d_hist);
    k_radix_hist_scan<<<c->passes, RS_BINS, 0, s>>>(c->d_hist);
    launches += 2;
    if (timed) CU(cudaEventRecord(c->ev[6], s));
    for (int p = 0; p < c->passes; ++p) {
        k_onesweep<<<c->num_sms * RS_BLOCKS_PER_SM, RS_THREADS, RS_SMEM_BYTES, s>>>(c->d_key[cur], c->d_val[cur], c->d_key[cur ^ 1],
                                                                     c->d_val[cur ^ 1], cnt, p, 8 * p, st);
        ++launches;
        cur ^= 1;
    }
    }
    c->sorted_buf = cur;
    if (c->peers.n_bands > 0 && !c->x_sums) {  // the band's non-zero sums, stored straight into every band's mailbox
        k_band_publish<<<c->peers.n_bands, 256, 0, s>>>(c->d_params, c->d_ctr, c->cap, c->radix_mode ? 1 : 0, c->p_list, c->peers);
        ++launches;
    }
    if (timed) CU(cudaEventRecord(c->ev[7], s));
    CU(cudaGetLastError());
    return SLPR_OK;
}

// Second half: winding prefix, spans, draw records, pixels. With a band exchange configured it starts by
// turning the gathered sums of all bands into this band's winding corrections.
static int enqueue_back(slpr_ctx *c, cudaStream_t s, bool timed, int &launches) {
    const bool taps = (c->flags & SLPR_FLAG_TAPS) != 0;
    const int wide = c->num_sms * 8;
    const int cur = c->sorted_buf;
    const int *corr = nullptr;
    if (c->x_sums) {
        CU(cudaMemsetAsync(c->x_scan_temp, 0, c->x_scan_bytes, s));
        k_band_other<<<grid_for(c, c->P, 256, 8), 256, 0, s>>>(c->x_gathered, c->P, c->x_ranks, c->x_rank, c->x_d, c->x_corr);
        ScanI32Op op{c->x_d, c->x_e, (long long)c->P, nullptr, 0, nullptr};
        ScanTemp t{reinterpret_cast<unsigned long long *>(c->x_scan_temp + 256), reinterpret_cast<int *>(c->x_scan_temp)};
        k_lookback_scan<ScanI32Op><<<grid_for(c, (long long)(c->P / scan_tile<ScanI32Op>() + 1), 1, ScanI32Op::MIN_BLOCKS), SCAN_THREADS, 0, s>>>(op, t);
        k_band_corr<<<grid_for(c, c->P, 256, 8), 256, 0, s>>>(c->x_e, c->P, c->x_corr);
        launches += 3;
        corr = c->x_corr;
    }
    BandTable btab{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!c->x_sums && c->peers.n_bands > 0) {
        btab = BandTable{c->p_tab_path, c->p_tab_cum, c->p_tab_n, c->p_tab_z, &c->d_ctr->n_band_bp};
        k_band_merge<<<1, XB_MERGE_THREADS, XB_MERGE_SMEM, s>>>(c->d_params, c->d_ctr, c->peers, btab);
        ++launches;
    }
    // ---- winding scan + mark + flag scan + emit: one kernel, two chained look-backs
    SpanTaps stp{c->d_wn, c->t_sidx, c->t_skey32, c->t_flags, c->t_scan3};
    SpanTemp stmp{c->d_wsum, c->d_status[1], c->d_status[2], c->d_tickets + 1};
#if SLPR_SP_WPRE
    k_wsum<<<c->num_sms * 8, WSUM_THREADS, 0, s>>>(c->d_val[cur], c->d_ctr, c->cap, c->d_wsum);
    k_wscan<<<1, 1024, 0, s>>>(c->d_ctr, c->cap, c->d_wsum, c->d_wn);
    launches += 2;
#endif
    if (timed) CU(cudaEventRecord(c->ev[8], s));
    const int span_grid = c->num_sms * std::max(1, c->span_blocks_per_sm);
    const bool blend = (c->flags & SLPR_FLAG_BLEND) != 0;
    const BlendList bl{c->d_heads, c->d_nodes, c->node_cap};
    auto spans = taps ? (c->fill_fused ? k_spans<true, true, true> : k_spans<false, true, true>)
                      : !c->fill_fused ? k_spans<false, true, false> : (want_records(c) ? k_spans<true, true, false> : k_spans<true, false, false>);
    if (blend && c->fill_fused)
        spans = taps ? k_spans<true, true, true, true> : (want_records(c) ? k_spans<true, true, false, true> : k_spans<true, false, false, true>);
    spans<<<span_grid, SP_THREADS, SP_STAGE_BYTES, s>>>(c->d_key[cur], c->d_val[cur], c->d_finfo, c->d_rec, c->d_ctr, c->L, (int)c->iW, (int)c->iH,
                                                       c->cap, stp, stmp, corr, c->P, c->d_cells, c->cw, btab, bl);
    ++launches;
    if (taps) {
        k_scan3_fixup<<<wide, 256, 0, s>>>(c->d_ctr, c->cap, c->t_scan3);
        ++launches;
    }
    if (timed) CU(cudaEventRecord(c->ev[9], s));
    // ---- pixels
    uint8_t *fb = c->target ? c->target : c->fb_cur;
    const size_t stride = c->target ? c->target_stride : c->fb_stride;
    if (!c->fill_fused) {  // small frames: a grid-wide pass over the records spreads the few wide spans better
        if (blend) k_fill_cells<true><<<wide, 256, 0, s>>>(c->d_params, c->d_ctr, c->cap, c->d_rec, c->d_cells, c->cw, bl);
        else k_fill_cells<false><<<wide, 256, 0, s>>>(c->d_params, c->d_ctr, c->cap, c->d_rec, c->d_cells, c->cw, bl);
        ++launches;
    }
    if (timed) CU(cudaEventRecord(c->ev[10], s));
    if (c->ss == 4) {  // SLPR_FLAG_AA4: 2 x 2 cells of the 4x frame -> one pixel (box filter)
        auto resolve = c->fill_fused ? (blend ? k_resolve_aa4<true, true> : k_resolve_aa4<true, false>)
                                     : (blend ? k_resolve_aa4<false, true> : k_resolve_aa4<false, false>);
        resolve<<<wide, 256, 0, s>>>(c->d_params, c->d_rec, c->d_finfo, c->d_cells, c->cw, fb, stride, (int)c->W, (int)c->H, bl);
    } else {
        auto resolve = c->fill_fused ? (blend ? k_resolve<true, true> : k_resolve<true, false>)
                                     : (blend ? k_resolve<false, true> : k_resolve<false, false>);
        resolve<<<wide, 256, 0, s>>>(c->d_params, c->d_rec, c->d_finfo, c->d_cells, c->cw, fb, stride, bl);
    }
    ++launches;
    if (timed) CU(cudaEventRecord(c->ev[11], s));
    if (!c->x_sums && c->peers.n_bands > 0 && c->peers.root >= 0 && c->peers.root != c->peers.me) {  // gather by direct stores: tell the root this band's pixels are in place
        k_band_done<<<1, 1, 0, s>>>(c->d_params, c->peers);
        ++launches;
    }
    CU(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, s));
    CU(cudaGetLastError());
    return SLPR_OK;
}

static int enqueue_frame(slpr_ctx *c, cudaStream_t s, bool timed, int &launches) {
    int rc = enqueue_fragments(c, s, timed, launches);
    if (rc) return rc;
    rc = enqueue_sort(c, s, timed, launches);
    if (rc) return rc;
    return enqueue_back(c, s, timed, launches);
}

// Where are the cells of the draw records marked? Fused into k_spans the 16-byte records are not read back
// (-0.03 ms at 18 M records, -0.8 ms at 146 M), but a frame of a few hundred thousand records occupies only a few
// dozen span tiles, whose warps then fill its wide spans one after the other (+0.1 ms on tiger at 4K): measured on
// the B200, profiles/README.md. Hysteresis keeps an animation from flipping (a flip re-captures the graph).
#ifndef SLPR_FILL_FUSED
#define SLPR_FILL_FUSED -1 /* -1: by frame size; 0 / 1: forced (experiments) */
#endif
static bool choose_fill_mode(const slpr_ctx *c, long long n_fragments) {
    if (SLPR_FILL_FUSED >= 0) return SLPR_FILL_FUSED != 0;
    if (c->flags & SLPR_FLAG_FUSED_FILL) return true;
    if (c->flags & SLPR_FLAG_SEPARATE_FILL) return false;
    return c->fill_fused ? (n_fragments > (3ll << 19)) : (n_fragments > (1ll << 21));
}

// Long pieces chain by chain (walk.cuh)? Worth it when a frame has some but few of them: a warp per chain with one
// lane walking a curve's bisections is the opposite of what a million-piece frame wants, where k_walk packs 32 pieces
// of equal length into a warp. From the last frame's count, with hysteresis.
// Measured (profiles/README.md): shipped scenes at 4K 1.4-3.4x faster with it; the 16K frame (9 M pieces, thousands of
// them long) 4 % slower — there the walk is bound by throughput, not by its longest piece. Hence both limits.
static bool choose_long_mode(const slpr_ctx *c, const FrameCounters &k) {
    if (c->flags & SLPR_FLAG_NO_LONG_WALK) return false;
    const int hi_long = c->long_mode ? 6144 : 4096, hi_pieces = c->long_mode ? 600000 : 400000;
    return k.n_long > 0 && k.n_long <= hi_long && k.n_pieces <= hi_pieces;
}

static int size_buffers_from_count(slpr_ctx *c) {
    // first frame of a scene (or after an overflow): run the counting prefix synchronously to size
    // the fragment buffers, with 25 % head-room for later frames.
    if (!c->d_temp) {
        int rc = alloc_capacity(c, 1 << 14);  // provisional, gives the count phase its temp block
        if (rc) return rc;
    }
    k_set_params<<<1, 1, 0, c->stream>>>(c->d_params, c->hp);
    int l = 1;
    int rc = enqueue_count_phase(c, c->stream, false, l);
    if (rc) return rc;
    c->launches += l;
    CU(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const long long nf = c->h_ctr->n_fragments;
    if (nf < 0 || nf >= (1ll << 29) - (1ll << 26)) return fail(SLPR_ERR_INVALID, "frame has %lld fragments; limit is 2^29", nf);
    if (nf + 16 > c->cap) {
        const long long want = nf + nf / 4 + 65536;
        rc = alloc_capacity(c, (int)std::min<long long>(want, (1ll << 29) - 1));
        if (rc) return rc;
    }
    c->fill_fused = choose_fill_mode(c, nf);
    return SLPR_OK;
}

// The captured graph of the frame for the current target / framebuffer (captured and instantiated on first use and
// whenever a mode, a capacity or the band set-up changed). `launches` receives its kernel count.
static int frame_graph(slpr_ctx *c, cudaGraphExec_t *ge_out, int *launches) {
    const bool second = !c->target && c->fb_cur == c->d_fb2 && c->d_fb2;
    slpr_ctx::TargetGraph *tg = nullptr;
    if (c->target) {  // the entry of this target, else the least recently used one
        for (auto &t : c->tgraph)
            if (t.valid && t.target == c->target && t.stride == c->target_stride) tg = &t;
        if (!tg) {
            tg = (c->tgraph[0].used <= c->tgraph[1].used) ? &c->tgraph[0] : &c->tgraph[1];
            tg->valid = false;
            tg->target = c->target;
            tg->stride = c->target_stride;
        }
        tg->used = ++c->tgraph_clock;
    }
    cudaGraphExec_t &ge = tg ? tg->ge : (second ? c->gexec2 : c->gexec);
    bool &valid = tg ? tg->valid : (second ? c->graph2_valid : c->graph_valid);
    int &n = tg ? tg->launches : c->launches_per_frame;
    if (!valid) {
        if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int l = 0;
        int rc = enqueue_frame(c, c->stream, false, l);
        cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(SLPR_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(SLPR_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        n = l;
        valid = true;
    }
    *ge_out = ge;
    *launches = n;
    return SLPR_OK;
}

static int render_checks(slpr_ctx *c, const char *who) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (!c->scene_loaded) return fail(SLPR_ERR_STATE, "%s: no scene loaded (call slpr_load_scene first)", who);
    if (c->x_sums) return fail(SLPR_ERR_STATE, "%s: a band exchange is configured; use slpr_render_band_begin / _end", who);
    CU(cudaSetDevice(c->device));
    if (c->cap == 0) {
        int rc = size_buffers_from_count(c);
        if (rc) return rc;
    }
    return ensure_records(c);
}

// Everything a frame needs that can block the device — the counting pre-pass that sizes the buffers, allocations,
// graph capture, instantiation and upload — without rendering. Optional; slpr_render does the same on demand. Needed
// when several contexts of ONE process render bands that wait for each other (a spinning exchange kernel of one
// context would otherwise sit in front of another context's cudaMalloc).
extern "C" int slpr_prepare(slpr_ctx *c) {
    int rc = render_checks(c, "slpr_prepare");
    if (rc) return rc;
    if (!(c->flags & SLPR_FLAG_NO_GRAPH)) {
        cudaGraphExec_t ge = nullptr;
        int n = 0;
        if ((rc = frame_graph(c, &ge, &n))) return rc;
        CU(cudaGraphUpload(ge, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return SLPR_OK;
}

extern "C" int slpr_render(slpr_ctx *c) {
    int rc = render_checks(c, "slpr_render");
    if (rc) return rc;
    k_set_params<<<1, 1, 0, c->stream>>>(c->d_params, c->hp);
    ++c->launches;
    if (c->flags & SLPR_FLAG_NO_GRAPH) {
        int l = 0;
        rc = enqueue_frame(c, c->stream, true, l);
        if (rc) return rc;
        c->launches += l;
        c->launches_per_frame = l;
        c->stage_times_valid = true;
    } else {
        cudaGraphExec_t ge = nullptr;
        int n = 0;
        if ((rc = frame_graph(c, &ge, &n))) return rc;
        CU(cudaGraphLaunch(ge, c->stream));
        c->launches += n;
        c->stage_times_valid = false;
    }
    c->frame_pending = true;
    c->frame_done = false;
    return SLPR_OK;
}

// Wait for the frame; if it overflowed the fragment capacity, grow and render it again.
// Which sort is cheaper for the frame just rendered? Microsecond models fitted to the B200 measurements
// in profiles/README.md: the radix sort costs a fixed ~8 us per pass plus ~8.3 ps per fragment and pass; the
// segmented sort streams the pairs once (~9.4 ps per fragment) but pays ~25 us when warps have to sort paths
// of 129..512 fragments and ~45 us per wave of 2 x #SM paths of 513..4096 fragments (one block each).
// Small scenes with a few long paths are therefore sorted by radix, big scenes of small paths segmented.
static bool segmented_sort_pays(const slpr_ctx *c, const FrameCounters &k) {
    if (k.stat_huge) return false;
    const double nf = (double)k.n_fragments;
    const double t_radix = 15.0 + nf * 3.7e-6 + c->passes * (8.0 + nf * 8.3e-6);
    const double waves = std::ceil((double)k.stat_big / (2.0 * c->num_sms));
    const double t_seg = 12.0 + nf * 9.4e-6 + (k.stat_mid ? 25.0 : 0.0) + (k.stat_big ? 40.0 + 45.0 * waves : 0.0);
    // hysteresis: leave the current mode only for a clear win
    return c->radix_mode ? (t_seg * 1.25 < t_radix) : (t_seg < t_radix * 1.25);
}

// After a whole frame: which sort and which coverage pass suit this scene and view, from the frame's counters
// (used from the next frame on; a change re-captures the graph).
static int settle_modes(slpr_ctx *c, const FrameCounters &k) {
    if (k.fix_missed) return fail(SLPR_ERR_STATE, "internal: a piece started below its start parameter without ending below it");
    if (!(c->flags & (SLPR_FLAG_RADIX_SORT | SLPR_FLAG_SEGMENTED_SORT))) {
        const bool seg = segmented_sort_pays(c, k);
        if (seg == c->radix_mode) {  // the other sort suits this scene and view better
            c->radix_mode = !seg;
            invalidate_graphs(c);
        }
    }
    if (choose_fill_mode(c, k.n_fragments) != c->fill_fused) {
        c->fill_fused = !c->fill_fused;
        invalidate_graphs(c);
    }
    if (choose_long_mode(c, k) != c->long_mode) {
        c->long_mode = !c->long_mode;
        invalidate_graphs(c);
    }
    return SLPR_OK;
}

// SLPR_FLAG_BLEND: did the frame ask for more list nodes than the pool holds? Then it lacks some translucent
// coverage and must be rendered again; grow_nodes() makes the pool fit that frame (plus a quarter).
static bool nodes_short(const slpr_ctx *c, const FrameCounters &k, uint32_t pool) {
    return (c->flags & SLPR_FLAG_BLEND) && (k.n_blend_nodes < 0 || (uint32_t)k.n_blend_nodes > pool);
}

static int grow_nodes(slpr_ctx *c, const FrameCounters &k) {
    const long long need = k.n_blend_nodes;
    if (need < 0 || need >= (1ll << 30)) return fail(SLPR_ERR_INVALID, "frame needs %lld blend nodes; limit is 2^30", need);
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_nodes); c->d_nodes = nullptr;
    c->node_cap = (uint32_t)std::min<long long>(need + need / 4 + 65536, (1ll << 30));
    CU(cudaMalloc(&c->d_nodes, (size_t)c->node_cap * sizeof(uint2)));
    invalidate_graphs(c);
    return SLPR_OK;
}

// Exact bands with the device-side exchange: a band cannot redo a frame on its own (the other bands have consumed
// what it published), so a void frame is reported and the caller renders it again on every band with a new seq.
static int finish_peer_frame(slpr_ctx *c) {
    CU(cudaStreamSynchronize(c->stream));
    c->frame_pending = false;
    const FrameCounters &k = *c->h_ctr;
    if (k.overflow) {
        const long long nf = k.n_fragments;
        if (nf < 0 || nf >= (1ll << 29) - (1ll << 26)) return fail(SLPR_ERR_INVALID, "band has %lld fragments; limit is 2^29", nf);
        int rc = alloc_capacity(c, (int)std::min<long long>(nf + nf / 4 + 65536, (1ll << 29) - 1));
        if (rc) return rc;
        return fail(SLPR_ERR_RETRY, "band %d outgrew its fragment buffers (now %d): render the frame again on every band", c->peers.me, c->cap);
    }
    if (k.stat_huge && !c->radix_mode) {
        c->radix_mode = true;
        invalidate_graphs(c);
        return fail(SLPR_ERR_RETRY, "band %d has a path too long for the segmented sort (now radix): render the frame again on every band", c->peers.me);
    }
    if (nodes_short(c, k, c->node_cap)) {
        int rc = grow_nodes(c, k);
        if (rc) return rc;
        return fail(SLPR_ERR_RETRY, "band %d outgrew its blend node pool (now %u): render the frame again on every band", c->peers.me, c->node_cap);
    }
    if (k.band_void == 2) return fail(SLPR_ERR_STATE, "band %d: another band did not publish its winding sums within the time-out", c->peers.me);
    if (k.band_void) return fail(SLPR_ERR_RETRY, "band %d: the frame was void on another band (%d): render it again on every band", c->peers.me, k.band_void);
    if (k.fix_missed) return fail(SLPR_ERR_STATE, "internal: a piece started below its start parameter without ending below it");
    if (choose_fill_mode(c, k.n_fragments) != c->fill_fused) {  // (the sort mode of a band only ever moves to radix, above)
        c->fill_fused = !c->fill_fused;
        invalidate_graphs(c);
    }
    if (choose_long_mode(c, k) != c->long_mode) {
        c->long_mode = !c->long_mode;
        invalidate_graphs(c);
    }
    c->frame_done = true;
    return SLPR_OK;
}

static int finish_frame(slpr_ctx *c) {
    if (!c->frame_pending && !c->frame_done) return fail(SLPR_ERR_STATE, "no frame has been rendered");
    CU(cudaSetDevice(c->device));
    if (c->peers.n_bands > 0 && !c->x_sums && c->frame_pending) return finish_peer_frame(c);
    if (!c->frame_pending) {  // finished before (a second read-back, or slpr_draw_records): its counters were dealt with then
        CU(cudaStreamSynchronize(c->stream));
        return SLPR_OK;
    }
    for (int attempt = 0; attempt < 4; ++attempt) {
        CU(cudaStreamSynchronize(c->stream));
        c->frame_pending = false;
        if (c->h_ctr->sort_fallback && !c->radix_mode && !c->h_ctr->overflow) {
            // a path outgrew the segmented sort: this scene uses the radix sort from now on; redo the frame
            c->radix_mode = true;
            invalidate_graphs(c);
            int rc2 = slpr_render(c);
            if (rc2) return rc2;
            continue;
        }
        if (!c->h_ctr->overflow && nodes_short(c, *c->h_ctr, c->node_cap)) {
            int rc2 = grow_nodes(c, *c->h_ctr);
            if (!rc2) rc2 = slpr_render(c);
            if (rc2) return rc2;
            continue;
        }
        if (!c->h_ctr->overflow) {
            int rc = settle_modes(c, *c->h_ctr);
            if (rc) return rc;
            c->frame_done = true;
            return SLPR_OK;
        }
        const long long nf = c->h_ctr->n_fragments;
        if (nf < 0 || nf >= (1ll << 29) - (1ll << 26)) return fail(SLPR_ERR_INVALID, "frame has %lld fragments; limit is 2^29", nf);
        int rc = alloc_capacity(c, (int)std::min<long long>(nf + nf / 4 + 65536, (1ll << 29) - 1));
        if (rc) return rc;
        rc = slpr_render(c);
        if (rc) return rc;
    }
    return fail(SLPR_ERR_STATE, "frame kept overflowing its fragment buffers");
}

extern "C" int slpr_synchronize(slpr_ctx *c) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (c->frame_pending) return finish_frame(c);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return SLPR_OK;
}

extern "C" int slpr_readback(slpr_ctx *c, uint8_t *rgba, size_t stride_bytes) {
    if (!c || !rgba) return fail(SLPR_ERR_INVALID, "slpr_readback: null argument");
    if (stride_bytes < (size_t)c->W * 4) return fail(SLPR_ERR_INVALID, "slpr_readback: stride smaller than a row");
    int rc = finish_frame(c);
    if (rc) return rc;
    const uint8_t *fb = c->target ? c->target : c->fb_cur;
    const size_t stride = c->target ? c->target_stride : c->fb_stride;
    CU(copy_frame_to_host(c, rgba, stride_bytes, fb, stride, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SLPR_OK;
}

// Stage 5 alone: the reference's draw call (scanline_rasterizer.cpp:611-656: clear to white, LINE_LIST over output_buf
// through scanlinepr.vert / .frag) over a caller-supplied list of draw records in the reference's own format
// (yx, width, fill_info, frag_index) — what workdir/test_data*.csv hold. The frame is then read with slpr_readback /
// slpr_framebuffer like a rendered one. Records must lie on the 2 x 2 fragment grid (x, y and width even), as every
// record the path emits does: the coverage grid works in those cells.
extern "C" int slpr_draw_records(slpr_ctx *c, const int32_t *records, uint64_t n_records) {
    if (!c || (!records && n_records)) return fail(SLPR_ERR_INVALID, "slpr_draw_records: null argument");
    if (n_records >= (1ull << 30)) return fail(SLPR_ERR_INVALID, "slpr_draw_records: too many records");
    if (c->ss != 1 || (c->flags & SLPR_FLAG_BLEND)) return fail(SLPR_ERR_STATE, "slpr_draw_records: not with SLPR_FLAG_AA4 / SLPR_FLAG_BLEND");
    if (c->hp.cull || c->peers.n_bands > 0) return fail(SLPR_ERR_STATE, "slpr_draw_records: not in band mode");
    for (uint64_t i = 0; i < n_records; ++i) {
        const int32_t yx = records[4 * i], w = records[4 * i + 1];
        if (((yx & 0xFFFF) | (yx >> 16) | w) & 1 || w < 0)
            return fail(SLPR_ERR_INVALID, "slpr_draw_records: record %llu is off the 2 x 2 fragment grid", (unsigned long long)i);
    }
    CU(cudaSetDevice(c->device));
    if (c->frame_pending) {
        int rc = finish_frame(c);
        if (rc) return rc;
    }
    if (!c->d_temp) {
        int rc = alloc_capacity(c, 1 << 14);
        if (rc) return rc;
    }
    int4 *d_rec = nullptr;
    CU(cudaMalloc(&d_rec, std::max<size_t>(n_records, 1) * sizeof(int4)));
    cudaStream_t s = c->stream;
    const int n = (int)n_records;
    const BlendList bl{c->d_heads, c->d_nodes, c->node_cap};
    uint8_t *fb = c->target ? c->target : c->fb_cur;
    const size_t stride = c->target ? c->target_stride : c->fb_stride;
    const int wide = c->num_sms * 8;
    cudaError_t e = cudaMemcpyAsync(d_rec, records, n_records * sizeof(int4), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_temp, 0, c->temp_bytes, s);  // the frame's counters: nothing void
    if (e == cudaSuccess) e = cudaMemcpyAsync(&c->d_ctr->n_records, &n, sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        k_set_params<<<1, 1, 0, s>>>(c->d_params, c->hp);
        k_fill_cells<false><<<wide, 256, 0, s>>>(c->d_params, c->d_ctr, c->cap, d_rec, c->d_cells, c->cw, bl);
        k_resolve<false, false><<<wide, 256, 0, s>>>(c->d_params, d_rec, c->d_finfo, c->d_cells, c->cw, fb, stride, bl);
        c->launches += 3;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_rec);
    if (e != cudaSuccess) return fail(SLPR_ERR_CUDA, "slpr_draw_records: %s", cudaGetErrorString(e));
    c->frame_pending = false;
    c->frame_done = true;
    return SLPR_OK;
}

extern "C" int slpr_render_to_host(slpr_ctx *c, const float rows[16], uint8_t *rgba, size_t stride_bytes) {
    int rc = slpr_set_mvp(c, rows);
    if (rc) return rc;
    if ((rc = slpr_render(c))) return rc;
    return slpr_readback(c, rgba, stride_bytes);
}

// Pipelined end-to-end path: frame i renders into one of two framebuffers while frame i-1 is still
// being copied to the host on a second stream. The pixels of a submitted frame are valid in `rgba`
// after slpr_wait_host() (or after the second-next submit returns).
// A frame can turn out invalid only after it ran (its fragments outgrew the buffers sized from earlier frames,
// or one of its paths outgrew the segmented sort — both common in an animation that zooms in). Its counters
// are copied into the slot right behind its kernels; they are looked at when the slot comes round again and in
// slpr_wait_host, and an invalid frame is then rendered again, synchronously, into the caller's buffer.
static bool slot_invalid(const slpr_ctx *c, const slpr_ctx::PipeSlot &p) {
    return p.in_flight && (p.h->overflow || (p.h->sort_fallback && !p.was_radix) || nodes_short(c, *p.h, p.node_cap));
}

static int pipe_recover(slpr_ctx *c, int first_slot) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    c->frame_pending = false;
    c->copy_pending[0] = c->copy_pending[1] = false;
    uint8_t *const fb_keep = c->fb_cur;
    for (int i = 0; i < 2; ++i) {  // older frame first
        slpr_ctx::PipeSlot &p = c->pslot[(first_slot + i) & 1];
        const bool bad = slot_invalid(c, p);
        p.in_flight = false;
        if (!bad) continue;
        c->fb_cur = ((first_slot + i) & 1) ? c->d_fb2 : c->d_fb;
        int rc = slpr_set_mvp(c, p.rows);
        if (!rc) rc = slpr_render(c);
        if (!rc) rc = finish_frame(c);  // grows the buffers / switches the sort until the frame is whole
        if (rc) { c->fb_cur = fb_keep; return rc; }
        CU(copy_frame_to_host(c, p.rgba, p.stride, c->fb_cur, c->fb_stride, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        ++c->pipe_redone;
    }
    c->fb_cur = fb_keep;
    return SLPR_OK;
}

extern "C" int slpr_submit_to_host(slpr_ctx *c, const float rows[16], uint8_t *rgba, size_t stride_bytes) {
    if (!c || !rows || !rgba) return fail(SLPR_ERR_INVALID, "slpr_submit_to_host: null argument");
    if (stride_bytes < (size_t)c->W * 4) return fail(SLPR_ERR_INVALID, "slpr_submit_to_host: stride smaller than a row");
    if (c->target) return fail(SLPR_ERR_STATE, "slpr_submit_to_host: not available with an external target");
    CU(cudaSetDevice(c->device));
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&c->d_fb2, c->fb_stride * c->H));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&c->ev_rendered[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
            CU(cudaMallocHost(&c->pslot[i].h, sizeof(FrameCounters)));
        }
    }
    const int k = (int)(c->pipe_frame++ & 1u);
    slpr_ctx::PipeSlot &slot = c->pslot[k];
    if (slot.in_flight) {  // the frame submitted two calls ago: finished long since; was it whole?
        CU(cudaEventSynchronize(c->ev_rendered[k]));
        if (slot_invalid(c, slot)) {
            int rc = pipe_recover(c, k);
            if (rc) return rc;
        } else if (slot.was_radix == c->radix_mode) {  // (not while a mode change is still working its way through the pipeline)
            int rc = settle_modes(c, *slot.h);
            if (rc) return rc;
        }
        slot.in_flight = false;
    }
    if (c->copy_pending[k]) {  // the copy that last read this framebuffer must be done before it is rendered into again
        CU(cudaStreamWaitEvent(c->stream, c->ev_copied[k], 0));
        c->copy_pending[k] = false;
    }
    c->fb_cur = k ? c->d_fb2 : c->d_fb;
    int rc = slpr_set_mvp(c, rows);
    if (!rc) rc = slpr_render(c);
    if (rc) return rc;
    memcpy(slot.rows, rows, sizeof slot.rows);
    slot.rgba = rgba; slot.stride = stride_bytes; slot.was_radix = c->radix_mode; slot.node_cap = c->node_cap; slot.in_flight = true;
    CU(cudaMemcpyAsync(slot.h, c->d_ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev_rendered[k], c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_rendered[k], 0));
    CU(copy_frame_to_host(c, rgba, stride_bytes, c->fb_cur, c->fb_stride, c->copy_stream));
    CU(cudaEventRecord(c->ev_copied[k], c->copy_stream));
    c->copy_pending[k] = true;
    return SLPR_OK;
}

extern "C" int slpr_wait_host(slpr_ctx *c) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    if (!c->copy_stream) return c->frame_pending ? finish_frame(c) : SLPR_OK;
    // every submitted frame whole and in its host buffer: redo the (at most two) frames still unchecked
    int rc = pipe_recover(c, (int)(c->pipe_frame & 1u));
    if (!rc) c->frame_done = true;
    return rc;
}

extern "C" uint64_t slpr_pipeline_redone(slpr_ctx *c) { return c ? c->pipe_redone : 0; }

extern "C" int slpr_framebuffer(slpr_ctx *c, void **dev_rgba, size_t *stride_bytes) {
    if (!c || !dev_rgba || !stride_bytes) return fail(SLPR_ERR_INVALID, "slpr_framebuffer: null argument");
    *dev_rgba = c->target ? c->target : c->fb_cur;
    *stride_bytes = c->target ? c->target_stride : c->fb_stride;
    return SLPR_OK;
}

extern "C" int slpr_get_counts(slpr_ctx *c, uint32_t *nf, uint32_t *nof, uint32_t *nspan) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    int rc = finish_frame(c);
    if (rc) return rc;
    if (nf) *nf = (uint32_t)c->h_ctr->n_fragments;
    if (nof) *nof = (uint32_t)c->h_ctr->n_out_frag;
    if (nspan) *nspan = (uint32_t)c->h_ctr->n_span;
    return SLPR_OK;
}

extern "C" int slpr_debug_copy(slpr_ctx *c, int which, void *dst, size_t bytes) {
    if (!c || !dst) return fail(SLPR_ERR_INVALID, "slpr_debug_copy: null argument");
    int rc = finish_frame(c);
    if (rc) return rc;
    const size_t nf = (size_t)c->h_ctr->n_fragments, no = (size_t)c->h_ctr->n_records;
    const void *src = nullptr;
    size_t avail = 0;
    bool tap = false;
    switch (which) {
        case SLPR_TAP_TRANSFORMED_POS: src = c->d_tpos; avail = (size_t)c->np * 8; break;
        case SLPR_TAP_PATH_VISIBLE: src = c->d_pvis; avail = (size_t)c->P * 4; break;
        case SLPR_TAP_CUT_CACHE: src = c->d_cut; avail = (size_t)c->nc * 20; break;
        case SLPR_TAP_CURVE_COUNT: src = c->d_count; avail = (size_t)c->nc * 4; break;
        case SLPR_TAP_CURVE_OFFSET: src = c->d_offset; avail = ((size_t)c->nc + 1) * 4; break;
        case SLPR_TAP_INTERSECTION: src = c->d_inter; avail = nf * 8; tap = true; break;
        case SLPR_TAP_KEY: src = c->t_key32; avail = (nf + 1) * 4; tap = true; break;
        case SLPR_TAP_PATH: src = c->t_path; avail = nf * 4; tap = true; break;
        case SLPR_TAP_WINDING: src = c->t_wind; avail = nf * 4; tap = true; break;
        case SLPR_TAP_SORTED_KEY: src = c->t_skey32; avail = nf * 4; tap = true; break;
        case SLPR_TAP_SORTED_INDEX: src = c->t_sidx; avail = nf * 4; tap = true; break;
        case SLPR_TAP_WINDING_SCAN: src = c->d_wn; avail = (nf + 1) * 4; tap = true; break;
        case SLPR_TAP_FLAGS: src = c->t_flags; avail = 2 * nf * 4; tap = true; break;
        case SLPR_TAP_FLAG_SCAN: src = c->t_scan3; avail = (2 * nf + 1) * 4; tap = true; break;
        case SLPR_TAP_RECORDS:
            if (!want_records(c)) return fail(SLPR_ERR_STATE, "slpr_debug_copy: the draw records need SLPR_FLAG_RECORDS (or SLPR_FLAG_TAPS)");
            src = c->d_rec; avail = no * 16; break;
        case SLPR_TAP_SEGMENTS: src = c->d_seg_tap; avail = ((size_t)c->P + 1) * 4; break;
        default: return fail(SLPR_ERR_INVALID, "slpr_debug_copy: unknown buffer %d", which);
    }
    if (tap && !(c->flags & SLPR_FLAG_TAPS)) return fail(SLPR_ERR_STATE, "slpr_debug_copy: buffer %d needs SLPR_FLAG_TAPS", which);
    if (bytes > avail) return fail(SLPR_ERR_INVALID, "slpr_debug_copy: asked %zu bytes of buffer %d, it has %zu", bytes, which, avail);
    if (bytes) CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return SLPR_OK;
}

extern "C" int slpr_stage_ms(slpr_ctx *c, float *ms, int n) {
    if (!c || !ms) return fail(SLPR_ERR_INVALID, "slpr_stage_ms: null argument");
    if (!c->stage_times_valid) return fail(SLPR_ERR_STATE, "slpr_stage_ms: needs a frame rendered with SLPR_FLAG_NO_GRAPH");
    int rc = finish_frame(c);
    if (rc) return rc;
    for (int i = 0; i < n && i < SLPR_STAGE_COUNT; ++i) CU(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
    return SLPR_OK;
}

extern "C" int slpr_sort_info(slpr_ctx *c, uint32_t *key_bits, uint32_t *passes, uint32_t *key_bytes) {
    if (!c || !c->scene_loaded) return fail(SLPR_ERR_STATE, "slpr_sort_info: no scene loaded");
    if (key_bits) *key_bits = (uint32_t)c->key_bits;
    if (passes) *passes = (uint32_t)c->passes;
    if (key_bytes) *key_bytes = 8;
    return SLPR_OK;
}

// ---- exact row bands across GPUs (bands.cuh) ---------------------------------------------------
extern "C" int slpr_band_exchange_ints(slpr_ctx *c, size_t *ints_per_band) {
    if (!c || !ints_per_band) return fail(SLPR_ERR_INVALID, "slpr_band_exchange_ints: null argument");
    if (!c->scene_loaded) return fail(SLPR_ERR_STATE, "slpr_band_exchange_ints: no scene loaded");
    *ints_per_band = 3 * (size_t)c->P;
    return SLPR_OK;
}

extern "C" int slpr_set_band_exchange(slpr_ctx *c, int32_t *dev_sums, const int32_t *dev_gathered, int n_bands, int band) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (!c->scene_loaded) return fail(SLPR_ERR_STATE, "slpr_set_band_exchange: no scene loaded");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    invalidate_graphs(c);
    free_exchange(c);
    if (!dev_sums && !dev_gathered) return SLPR_OK;  // back to independent bands / full frames
    if (!dev_sums || !dev_gathered || n_bands < 1 || band < 0 || band >= n_bands)
        return fail(SLPR_ERR_INVALID, "slpr_set_band_exchange: need both buffers and 0 <= band < n_bands");
    if (c->flags & SLPR_FLAG_NO_GRAPH) return fail(SLPR_ERR_UNSUPPORTED, "slpr_set_band_exchange: not available with SLPR_FLAG_NO_GRAPH");
    const size_t P = std::max<size_t>(c->P, 1);
    CU(cudaMalloc(&c->x_d, P * 4));
    CU(cudaMalloc(&c->x_e, (P + 4) * 4));
    CU(cudaMalloc(&c->x_corr, 2 * P * 4));
    c->x_scan_bytes = 256 + (P / SCAN_TILE_MIN + 2) * 8;
    CU(cudaMalloc(&c->x_scan_temp, c->x_scan_bytes));
    CU(cudaEventCreateWithFlags(&c->x_event, cudaEventDisableTiming));
    c->x_sums = dev_sums;
    c->x_gathered = dev_gathered;
    c->x_ranks = n_bands;
    c->x_rank = band;
    return SLPR_OK;
}

// One captured graph of `enqueue` on the context's stream, re-captured when `valid` is false.
template <class F>
static int launch_graph(slpr_ctx *c, cudaGraphExec_t &ge, bool &valid, int &n_launches, F enqueue) {
    if (!valid) {
        if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int l = 0;
        int rc = enqueue(l);
        cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(SLPR_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(SLPR_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        n_launches = l;
        valid = true;
    }
    CU(cudaGraphLaunch(ge, c->stream));
    c->launches += n_launches;
    return SLPR_OK;
}

extern "C" int slpr_render_band_begin(slpr_ctx *c) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (!c->x_sums) return fail(SLPR_ERR_STATE, "slpr_render_band_begin: call slpr_set_band_exchange first");
    CU(cudaSetDevice(c->device));
    if (c->cap == 0) {
        int rc = size_buffers_from_count(c);
        if (rc) return rc;
    }
    for (int attempt = 0; attempt < 4; ++attempt) {
        { int rc = ensure_records(c); if (rc) return rc; }
        k_set_params<<<1, 1, 0, c->stream>>>(c->d_params, c->hp);
        ++c->launches;
        int rc = launch_graph(c, c->gexec_a, c->graph_a_valid, c->launches_a, [&](int &l) {
            int r = enqueue_fragments(c, c->stream, false, l);
            if (r) return r;
            CU(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
            return (int)SLPR_OK;
        });
        if (rc) return rc;
        CU(cudaEventRecord(c->x_event, c->stream));
        // the sort does not need the other bands: it runs while the caller exchanges the sums
        rc = launch_graph(c, c->gexec_s, c->graph_s_valid, c->launches_s, [&](int &l) { return enqueue_sort(c, c->stream, false, l); });
        if (rc) return rc;
        // The exchange happens outside the library, so capacity and sort mode are settled here, before it.
        // Only the sums (and counters) are waited for; the sort is still running when this returns.
        CU(cudaEventSynchronize(c->x_event));
        if (c->h_ctr->overflow) {
            const long long nf = c->h_ctr->n_fragments;
            if (nf < 0 || nf >= (1ll << 29) - (1ll << 26)) return fail(SLPR_ERR_INVALID, "band has %lld fragments; limit is 2^29", nf);
            rc = alloc_capacity(c, (int)std::min<long long>(nf + nf / 4 + 65536, (1ll << 29) - 1));
            if (rc) return rc;
            continue;
        }
        if (c->h_ctr->stat_huge && !c->radix_mode) {  // a path too long for the segmented sort (k_path_segments): sort again by radix
            CU(cudaStreamSynchronize(c->stream));
            c->radix_mode = true;
            invalidate_graphs(c);
            continue;
        }
        c->band_begun = true;
        return SLPR_OK;
    }
    return fail(SLPR_ERR_STATE, "band kept overflowing its fragment buffers");
}

extern "C" int slpr_render_band_end(slpr_ctx *c) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (!c->x_sums || !c->band_begun) return fail(SLPR_ERR_STATE, "slpr_render_band_end: no band in flight (slpr_render_band_begin first)");
    CU(cudaSetDevice(c->device));
    c->band_begun = false;
    const bool second = !c->target && c->fb_cur == c->d_fb2 && c->d_fb2;
    slpr_ctx::TargetGraph *tg = nullptr;
    if (c->target) {
        for (auto &t : c->tgraph)
            if (t.valid && t.target == c->target && t.stride == c->target_stride) tg = &t;
        if (!tg) {
            tg = (c->tgraph[0].used <= c->tgraph[1].used) ? &c->tgraph[0] : &c->tgraph[1];
            tg->valid = false;
            tg->target = c->target;
            tg->stride = c->target_stride;
        }
        tg->used = ++c->tgraph_clock;
    }
    cudaGraphExec_t &ge = tg ? tg->ge : (second ? c->gexec2 : c->gexec);
    bool &valid = tg ? tg->valid : (second ? c->graph2_valid : c->graph_valid);
    int rc = launch_graph(c, ge, valid, tg ? tg->launches : c->launches_per_frame, [&](int &l) { return enqueue_back(c, c->stream, false, l); });
    if (rc) return rc;
    c->stage_times_valid = false;
    c->frame_pending = true;
    c->frame_done = false;
    return SLPR_OK;
}

// ---- exact row bands, device-side sparse exchange ---------------------------------------------------------
extern "C" int slpr_band_mailbox(slpr_ctx *c, void **dev_ptr, size_t *bytes) {
    if (!c || !dev_ptr) return fail(SLPR_ERR_INVALID, "slpr_band_mailbox: null argument");
    CU(cudaSetDevice(c->device));
    if (!c->p_box) {
        CU(cudaMalloc(&c->p_box, sizeof(BandMailbox)));
        CU(cudaMemset(c->p_box, 0, sizeof(BandMailbox)));
        CU(cudaMalloc(&c->p_list, (size_t)XB_CAP * sizeof(BandEntry)));
        CU(cudaMalloc(&c->p_tab_path, (size_t)XB_TOTAL * 4));
        CU(cudaMalloc(&c->p_tab_cum, (size_t)XB_TOTAL * 4));
        CU(cudaMalloc(&c->p_tab_n, (size_t)XB_TOTAL * 4));
        CU(cudaMalloc(&c->p_tab_z, (size_t)XB_TOTAL * 4));
        CU(cudaFuncSetAttribute(k_band_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XB_MERGE_SMEM));
    }
    *dev_ptr = c->p_box;
    if (bytes) *bytes = sizeof(BandMailbox);
    return SLPR_OK;
}

extern "C" int slpr_alloc_device(slpr_ctx *c, size_t bytes, void **dev_ptr) {
    if (!c || !dev_ptr || !bytes) return fail(SLPR_ERR_INVALID, "slpr_alloc_device: null argument");
    CU(cudaSetDevice(c->device));
    void *p = nullptr;
    CU(cudaMalloc(&p, bytes));
    c->dev_allocs.push_back(p);
    *dev_ptr = p;
    return SLPR_OK;
}

extern "C" int slpr_ipc_export(slpr_ctx *c, void *dev_ptr, unsigned char handle[64]) {
    if (!c || !dev_ptr || !handle) return fail(SLPR_ERR_INVALID, "slpr_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return SLPR_OK;
}

extern "C" int slpr_ipc_import(slpr_ctx *c, const unsigned char handle[64], void **dev_ptr) {
    if (!c || !dev_ptr || !handle) return fail(SLPR_ERR_INVALID, "slpr_ipc_import: null argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_mapped.push_back(p);
    *dev_ptr = p;
    return SLPR_OK;
}

extern "C" int slpr_set_band_peers(slpr_ctx *c, int n_bands, int band, int root, void *const *mailboxes) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    invalidate_graphs(c);
    if (n_bands == 0) { c->peers = BandPeers{}; return SLPR_OK; }
    if (!c->p_box) return fail(SLPR_ERR_STATE, "slpr_set_band_peers: call slpr_band_mailbox first");
    if (c->x_sums) return fail(SLPR_ERR_STATE, "slpr_set_band_peers: a host-driven exchange is configured (slpr_set_band_exchange)");
    if (!mailboxes || n_bands < 1 || n_bands > XB_MAX_BANDS || band < 0 || band >= n_bands || root < -1 || root >= n_bands)
        return fail(SLPR_ERR_INVALID, "slpr_set_band_peers: need 1 <= n_bands <= %d, 0 <= band < n_bands, -1 <= root < n_bands and the mailboxes", XB_MAX_BANDS);
    if (mailboxes[band] != (void *)c->p_box) return fail(SLPR_ERR_INVALID, "slpr_set_band_peers: mailboxes[band] must be this context's own mailbox");
    BandPeers p{};
    for (int i = 0; i < n_bands; ++i) {
        if (!mailboxes[i]) return fail(SLPR_ERR_INVALID, "slpr_set_band_peers: mailbox %d is null", i);
        p.box[i] = reinterpret_cast<BandMailbox *>(mailboxes[i]);
    }
    p.n_bands = n_bands; p.me = band; p.root = root;
    c->peers = p;
    return SLPR_OK;
}

extern "C" int slpr_render_band(slpr_ctx *c, uint32_t frame_seq) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (c->peers.n_bands <= 0) return fail(SLPR_ERR_STATE, "slpr_render_band: call slpr_set_band_peers first");
    c->hp.frame_seq = (int)frame_seq;
    const int slot = (int)(frame_seq & 1u);
    if (c->band_push_pending[slot]) {  // the push that last read this slot's buffer must be over before it is rendered into again
        CU(cudaSetDevice(c->device));
        CU(cudaStreamWaitEvent(c->stream, c->ev_band_pushed[slot], 0));
        c->band_push_pending[slot] = false;
    }
    return slpr_render(c);
}

// Gather by copy engine: the band just rendered (into the current target, addressed like a full frame) travels to
// `dst_frame` — the root's peer-mapped frame buffer — on a second stream, so that the transfer over NVLink overlaps
// the next frame's kernels and uses no SM; the "pixels are in place" flag for the root follows the copy.
extern "C" int slpr_band_push(slpr_ctx *c, uint32_t frame_seq, int root_band, void *dst_frame, size_t dst_stride) {
    if (!c || !dst_frame) return fail(SLPR_ERR_INVALID, "slpr_band_push: null argument");
    if (c->peers.n_bands <= 0 || root_band < 0 || root_band >= c->peers.n_bands) return fail(SLPR_ERR_STATE, "slpr_band_push: needs slpr_set_band_peers and a valid root band");
    if (dst_stride < (size_t)c->W * 4) return fail(SLPR_ERR_INVALID, "slpr_band_push: stride smaller than a row");
    CU(cudaSetDevice(c->device));
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    const int slot = (int)(frame_seq & 1u);
    if (!c->ev_band_rendered[slot]) {
        CU(cudaEventCreateWithFlags(&c->ev_band_rendered[slot], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_band_pushed[slot], cudaEventDisableTiming));
    }
    const uint8_t *src = c->target ? c->target : c->fb_cur;
    const size_t src_stride = c->target ? c->target_stride : c->fb_stride;
    const size_t row0 = (size_t)(c->H - (uint32_t)c->hp.band_y1 / c->ss);  // image rows [H - y1, H - y0) hold scanline rows [y0, y1)
    const size_t rows = (size_t)(c->hp.band_y1 - c->hp.band_y0) / c->ss;
    CU(cudaEventRecord(c->ev_band_rendered[slot], c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_band_rendered[slot], 0));
    const size_t row_bytes = (size_t)c->W * 4;
    if (src_stride == row_bytes && dst_stride == row_bytes)
        CU(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(dst_frame) + row0 * dst_stride, src + row0 * src_stride, rows * row_bytes, cudaMemcpyDeviceToDevice, c->copy_stream));
    else
        CU(cudaMemcpy2DAsync(reinterpret_cast<uint8_t *>(dst_frame) + row0 * dst_stride, dst_stride, src + row0 * src_stride, src_stride, row_bytes, rows,
                             cudaMemcpyDeviceToDevice, c->copy_stream));
    k_band_done_seq<<<1, 1, 0, c->copy_stream>>>(frame_seq, root_band, c->peers);
    ++c->launches;
    CU(cudaEventRecord(c->ev_band_pushed[slot], c->copy_stream));
    c->band_push_pending[slot] = true;
    CU(cudaGetLastError());
    return SLPR_OK;
}

extern "C" int slpr_band_wait_gather(slpr_ctx *c, uint32_t frame_seq) {
    if (!c) return fail(SLPR_ERR_INVALID, "null context");
    if (c->peers.n_bands <= 0) return fail(SLPR_ERR_STATE, "slpr_band_wait_gather: needs a configured exchange (slpr_set_band_peers)");
    CU(cudaSetDevice(c->device));
    k_band_wait_done<<<1, 32, 0, c->stream>>>(frame_seq, c->peers, c->d_ctr);
    ++c->launches;
    CU(cudaGetLastError());
    return SLPR_OK;
}

__global__ void k_count_diff(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t n, unsigned long long *__restrict__ out) {
    unsigned long long d = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d += a[i] != b[i];
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
    if ((threadIdx.x & 31) == 0 && d) atomicAdd(out, d);
}

// How many 32-bit words (= RGBA8 pixels) differ between two device buffers: compares an assembled frame with the
// same frame rendered whole without moving a gigabyte to the host. Synchronous.
extern "C" int slpr_debug_diff_u32(slpr_ctx *c, const void *dev_a, const void *dev_b, size_t n_words, uint64_t *n_diff) {
    if (!c || !dev_a || !dev_b || !n_diff) return fail(SLPR_ERR_INVALID, "slpr_debug_diff_u32: null argument");
    CU(cudaSetDevice(c->device));
    unsigned long long *d = nullptr;
    CU(cudaMalloc(&d, 8));
    CU(cudaMemsetAsync(d, 0, 8, c->stream));
    k_count_diff<<<c->num_sms * 8, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(dev_a), reinterpret_cast<const uint32_t *>(dev_b), n_words, d);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(SLPR_ERR_CUDA, "slpr_debug_diff_u32: %s", cudaGetErrorString(e));
    *n_diff = h;
    return SLPR_OK;
}

// ---- pinned host memory next to the GPU ---------------------------------------------------------------------
// The end-to-end path moves a frame (33 MB at 4K) to host memory every ~1 ms per GPU; with eight ranks on a
// two-socket host, buffers that all sit on one socket's memory put half of the GPUs' traffic on the inter-socket
// link (e2e efficiency 0.41 at 8 GPUs). This allocator places the pages on the NUMA node the GPU
// hangs off (sysfs numa_node of its PCI function; mbind(2), no libnuma needed) and pins them for DMA.
static int gpu_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
    for (char *p = bus; *p; ++p) *p = (char)tolower(*p);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

extern "C" int slpr_host_alloc(slpr_ctx *c, size_t bytes, void **host_ptr, int *numa_node_out) {
    if (!c || !host_ptr || !bytes) return fail(SLPR_ERR_INVALID, "slpr_host_alloc: null argument");
    CU(cudaSetDevice(c->device));
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t len = align_up(bytes, page);
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return fail(SLPR_ERR_CUDA, "slpr_host_alloc: mmap of %zu bytes failed", len);
    int node = gpu_numa_node(c->device), bound = -1;  // -1: the GPU's node is unknown; -2: mbind refused
#ifdef SYS_mbind
    if (node >= 0 && node < 64) {
        unsigned long mask = 1ul << node;
        bound = (syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, &mask, 65ul, 0u) == 0) ? node : -2;
    }
#endif
    memset(p, 0, len);  // fault the pages in under the policy
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) { munmap(p, len); return fail(SLPR_ERR_CUDA, "slpr_host_alloc: cudaHostRegister failed: %s", cudaGetErrorString(e)); }
    c->host_allocs.push_back({p, len});
    *host_ptr = p;
    if (numa_node_out) *numa_node_out = bound;
    return SLPR_OK;
}

extern "C" int slpr_host_free(slpr_ctx *c, void *host_ptr) {
    if (!c || !host_ptr) return fail(SLPR_ERR_INVALID, "slpr_host_free: null argument");
    for (size_t i = 0; i < c->host_allocs.size(); ++i)
        if (c->host_allocs[i].first == host_ptr) {
            cudaHostUnregister(host_ptr);
            munmap(host_ptr, c->host_allocs[i].second);
            c->host_allocs.erase(c->host_allocs.begin() + (long)i);
            return SLPR_OK;
        }
    return fail(SLPR_ERR_INVALID, "slpr_host_free: not a pointer from slpr_host_alloc");
}

extern "C" int slpr_sort_mode(slpr_ctx *c, int *mode) {
    if (!c || !mode) return fail(SLPR_ERR_INVALID, "slpr_sort_mode: null argument");
    *mode = c->radix_mode ? 1 : 0;
    return SLPR_OK;
}

extern "C" int slpr_fill_mode(slpr_ctx *c, int *fused) {
    if (!c || !fused) return fail(SLPR_ERR_INVALID, "slpr_fill_mode: null argument");
    *fused = c->fill_fused ? 1 : 0;
    return SLPR_OK;
}

extern "C" int slpr_walk_info(slpr_ctx *c, uint32_t *n_pieces) {
    if (!c || !n_pieces) return fail(SLPR_ERR_INVALID, "slpr_walk_info: null argument");
    int rc = finish_frame(c);
    if (rc) return rc;
    *n_pieces = (uint32_t)c->h_ctr->n_pieces;
    return SLPR_OK;
}

extern "C" int slpr_long_walk_info(slpr_ctx *c, int *mode_on, uint32_t *n_long_pieces) {
    if (!c) return fail(SLPR_ERR_INVALID, "slpr_long_walk_info: null argument");
    int rc = finish_frame(c);
    if (rc) return rc;
    if (mode_on) *mode_on = c->long_mode ? 1 : 0;  // what the NEXT frame will use
    if (n_long_pieces) *n_long_pieces = (uint32_t)c->h_ctr->n_long;
    return SLPR_OK;
}

extern "C" uint64_t slpr_launch_count(slpr_ctx *c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------------------------------------
// stand-alone primitives
static int prim_temp(slpr_ctx *c, size_t bytes) {
    if (bytes > c->prim_temp_bytes) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_prim_temp);
        c->d_prim_temp = nullptr;
        CU(cudaMalloc(&c->d_prim_temp, bytes));
        c->prim_temp_bytes = bytes;
    }
    CU(cudaMemsetAsync(c->d_prim_temp, 0, bytes, c->stream));
    return SLPR_OK;
}

extern "C" int slpr_scan_i32(slpr_ctx *c, const int32_t *in, int32_t *out, uint64_t n) {
    if (!c || !in || !out) return fail(SLPR_ERR_INVALID, "slpr_scan_i32: null argument");
    if (n >= (1ull << 40)) return fail(SLPR_ERR_INVALID, "slpr_scan_i32: n too large");
    CU(cudaSetDevice(c->device));
    const size_t tiles = n / SCAN_TILE_MIN + 2;
    int rc = prim_temp(c, 256 + tiles * 8);
    if (rc) return rc;
    ScanI32Op op{in, out, (long long)n, nullptr, 0, nullptr};
    ScanTemp t{reinterpret_cast<unsigned long long *>(c->d_prim_temp + 256), reinterpret_cast<int *>(c->d_prim_temp)};
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (n >= (1ull << 22) && aligned && c->scan_tma_blocks_per_sm > 0) {
        // big arrays: the TMA-pipelined kernel; its round-robin tiles need every CTA resident
        const long long nchunks = (long long)((n + ST_TILE - 1) / ST_TILE);
        const int grid = (int)std::min<long long>(nchunks, (long long)c->num_sms * std::min(c->scan_tma_blocks_per_sm, SLPR_ST_CTAS));
        k_scan_tma<<<grid, ST_THREADS, ST_SMEM_BYTES, c->stream>>>(op, t);
    } else
        k_lookback_scan<ScanI32Op><<<grid_for(c, (long long)(n / scan_tile<ScanI32Op>() + 1), 1, ScanI32Op::MIN_BLOCKS), SCAN_THREADS, 0, c->stream>>>(op, t);
    ++c->launches;
    CU(cudaGetLastError());
    return SLPR_OK;
}


extern "C" int slpr_sort_pairs(slpr_ctx *c, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp,
                               uint64_t n, uint32_t key_bits, int *result_in_tmp) {
    if (!c || !keys || !vals || !keys_tmp || !vals_tmp) return fail(SLPR_ERR_INVALID, "slpr_sort_pairs: null argument");
    if (n >= (1ull << 30)) return fail(SLPR_ERR_INVALID, "slpr_sort_pairs: n must be < 2^30");
    if (key_bits == 0 || key_bits > 64) return fail(SLPR_ERR_INVALID, "slpr_sort_pairs: key_bits must be in [1,64]");
    CU(cudaSetDevice(c->device));
    const int passes = (int)(key_bits + 7) / 8;
    const int tiles_cap = (int)(n / RS_TILE + 2);
    const size_t o_hist = 256, o_lb = o_hist + (size_t)RS_MAX_PASSES * RS_BINS * 4;
    int rc = prim_temp(c, o_lb + (size_t)passes * tiles_cap * RS_BINS * 4);
    if (rc) return rc;
    SortCount cnt{nullptr, (long long)n, 0};
    SortTemp st{reinterpret_cast<uint32_t *>(c->d_prim_temp + o_hist), reinterpret_cast<uint32_t *>(c->d_prim_temp + o_lb),
                reinterpret_cast<int *>(c->d_prim_temp), tiles_cap};
    k_radix_hist<<<c->num_sms * 4, RH_THREADS, 0, c->stream>>>(keys, cnt, passes, st.hist);
    k_radix_hist_scan<<<passes, RS_BINS, 0, c->stream>>>(st.hist);
    uint64_t *kb[2] = {keys, keys_tmp};
    uint32_t *vb[2] = {vals, vals_tmp};
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        k_onesweep<<<c->num_sms * RS_BLOCKS_PER_SM, RS_THREADS, RS_SMEM_BYTES, c->stream>>>(kb[cur], vb[cur], kb[cur ^ 1], vb[cur ^ 1], cnt, p, 8 * p, st);
        cur ^= 1;
    }
    c->launches += 2 + passes;
    if (result_in_tmp) *result_in_tmp = cur;
    CU(cudaGetLastError());
    return SLPR_OK;
}
