/*
 * slpr.h — C ABI of the B200-native scanline path renderer (libslpr.so).
 *
 * This is the drop-in boundary for the compute path of galaxysailing/VkScanlinePR. Each entry
 * point names the reference interface it replaces (file:line relative to the reference root).
 * Plain pointers and sizes only: no C++ or torch types cross this boundary, no exceptions.
 * Every function returns an int status (SLPR_OK = 0) unless noted; the message for the last
 * failure on the calling thread is slpr_last_error().
 *
 * Threading: a context is one GPU + one stream and is not thread-safe; independent contexts are.
 * There is no CPU fallback: if no CUDA device is usable, slpr_create() fails loudly.
 */
#ifndef SLPR_H_
#define SLPR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SLPR_API __declspec(dllexport)
#else
#define SLPR_API __attribute__((visibility("default")))
#endif

typedef struct slpr_ctx slpr_ctx;

enum {
    SLPR_OK = 0,
    SLPR_ERR_INVALID = 1,     /* bad argument */
    SLPR_ERR_CUDA = 2,        /* CUDA runtime error (message has the CUDA string) */
    SLPR_ERR_STATE = 3,       /* call order violated (e.g. render before load_scene) */
    SLPR_ERR_UNSUPPORTED = 4, /* feature not built */
    SLPR_ERR_IO = 5,          /* file / parse error (host scene helpers) */
    SLPR_ERR_RETRY = 6        /* exact bands: the frame is void on some band; render it again on EVERY band (new frame_seq) */
};

enum {
    /* Keep every reference-format intermediate buffer (fragment_data planes etc.) so that
     * slpr_debug_copy() can return them for parity checks. Costs extra HBM traffic. */
    SLPR_FLAG_TAPS = 1u << 0,
    /* Evaluate LERP / dot / the cubic's coefficients with fused multiply-adds, as a Vulkan driver's compiler may (GLSL
     * allows contraction unless `precise`). The policy — which a*b + c shapes are fused — is defined in
     * oracle/oracle.c (orc_set_contract_fma) and matched bit for bit; the default (off) is the reading the executed
     * SPIR-V fixtures pin. Results differ from the default by rounding only; the bisection issues 14 instead of 21
     * instructions per step. */
    SLPR_FLAG_CONTRACT_FMA = 1u << 1,
    /* Launch kernels directly instead of replaying the captured CUDA graph of the frame. */
    SLPR_FLAG_NO_GRAPH = 1u << 2,
    /* Sort selection. Default: the context picks per scene and view from the path sizes of the last frame —
     * the one-pass segmented sort (csrc/segsort.cuh) for many small paths, the onesweep LSD radix sort
     * (csrc/radix.cuh) for small scenes with long paths; a path of more than 4096 fragments always means radix.
     * RADIX_SORT: always radix. SEGMENTED_SORT: segmented whenever no path exceeds 4096 fragments. */
    SLPR_FLAG_RADIX_SORT = 1u << 3,
    SLPR_FLAG_SEGMENTED_SORT = 1u << 4,
    /* Stage-5 coverage. Default: by frame size (slpr_fill_mode). FUSED_FILL: the span kernel always marks the cells
     * of its draw records itself; SEPARATE_FILL: always a separate pass over the records. Same pixels. */
    SLPR_FLAG_FUSED_FILL = 1u << 5,
    SLPR_FLAG_SEPARATE_FILL = 1u << 6,
    /* Order in which the monotone pieces are walked (a scheduling choice, results are identical): by default pieces
     * are walked longest first, and on scenes of more than two million curves the short ones also window by window
     * of consecutive curves (less HBM traffic). WINDOWED_WALK forces eight windows on any scene (tests). */
    SLPR_FLAG_WINDOWED_WALK = 1u << 7,
    /* Keep the reference's draw records (output_buf, gen_merged_fragment_and_span.comp:77,102) readable through
     * slpr_debug_copy(SLPR_TAP_RECORDS). Without it a big frame never writes them: the span kernel marks the
     * coverage grid directly (16 B per record of HBM traffic saved). Implied by SLPR_FLAG_TAPS. */
    SLPR_FLAG_RECORDS = 1u << 8,
    /* SURVEY section 8 f-1, beyond the reference: QUADRIC (3 points) and ARC (0x13: rational quadratic, 3 points + a weight per curve,
     * slpr_set_curve_weights) curves get real arithmetic — monotonic cuts, bisection crossings, fragment end points — in
     * place of the reference's `// TODO` arms (make_intersection_0.comp:141-144,273-276; make_intersection_1.comp:386-391;
     * gen_fragment.comp:66-69). The arithmetic is defined in oracle/oracle.c (orc_set_full_rvg) and matched bit for bit.
     * Scenes of lines and cubics render exactly as without the flag. Use with slpr_vg_load_rvg_full. */
    SLPR_FLAG_FULL_RVG = 1u << 9,
    /* SURVEY section 8 f-3, beyond the reference (which has no antialiasing: README.md:5, gen_fragment.comp:202): four coverage
     * samples per pixel. The pipeline runs at four times the frame's size (the matrix rows are scaled by 4, exactly), so a
     * pixel is 2 x 2 of its coverage cells; each cell holds the top-most path covering it (ordered, opaque compositing
     * per sample, as the reference composites per 2 x 2 block) and the pixel is their box-filtered average. Same result
     * as the reference path rendered at 4x and averaged over 4 x 4 blocks. Width and height at most 8191. */
    SLPR_FLAG_AA4 = 1u << 10,
    /* Never walk long monotone pieces chain by chain (csrc/walk.cuh, k_long_chains / k_long_emit; by default the context
     * turns that on by itself when a frame has few pieces of 62 or more crossings). A scheduling choice: same results. */
    SLPR_FLAG_NO_LONG_WALK = 1u << 11,
    /* SURVEY section 8 f-3, beyond the reference (which composites opaque: blendEnable = VK_FALSE, scanline_rasterizer.cpp:893-895,
     * so a fill's alpha byte only lands in the framebuffer): fills with 0 < alpha < 255 are composited "source over" in
     * path order onto the white clear colour, per coverage cell — with SLPR_FLAG_AA4 per sample, then averaged. Each
     * channel is (src * a + dst * (255 - a) + 127) / 255 in integers; alpha 255 overwrites, alpha 0 leaves the pixel
     * alone; the frame's alpha stays 255. Checked bit for bit against the oracle's orc_set_blend(1). Opaque paths keep the
     * atomicMax coverage marks; translucent ones append per-cell list nodes from a pool that grows when a frame needs
     * more (the frame is then rendered again, like a fragment overflow; with band peers: SLPR_ERR_RETRY). */
    SLPR_FLAG_BLEND = 1u << 12
};

/* Buffers that slpr_debug_copy() can return. Layouts are the reference's:
 * int32 unless noted; sizes in elements. */
enum {
    SLPR_TAP_TRANSFORMED_POS = 0, /* float[2*n_points]        transform_pos.comp:85            */
    SLPR_TAP_PATH_VISIBLE = 1,    /* int[n_paths]             transform_pos.comp:71-82         */
    SLPR_TAP_CUT_CACHE = 2,       /* float[5*n_curves]        make_intersection_0.comp:368-372 */
    SLPR_TAP_CURVE_COUNT = 3,     /* int[n_curves]            make_intersection_0.comp:410     */
    SLPR_TAP_CURVE_OFFSET = 4,    /* int[n_curves+1]          scan #1, SR.cpp:338-358          */
    SLPR_TAP_INTERSECTION = 5,    /* int[2*nf] (curve,tbits)  make_intersection_1.comp:361-375 */
    SLPR_TAP_KEY = 6,             /* plane 0 before sort      gen_fragment.comp:221            */
    SLPR_TAP_PATH = 7,            /* plane 2                  gen_fragment.comp:223            */
    SLPR_TAP_WINDING = 8,         /* plane 4 (unsorted delta) gen_fragment.comp:224            */
    SLPR_TAP_SORTED_KEY = 9,      /* plane 0 after sort       naive_seg_sort_pairs.comp        */
    SLPR_TAP_SORTED_INDEX = 10,   /* plane 1 after sort                                        */
    SLPR_TAP_WINDING_SCAN = 11,   /* int[nf+1] plane 3 after shuffle + scan #2, SR.cpp:479-506 */
    SLPR_TAP_FLAGS = 12,          /* int[2*nf] [frag|span]    mark_merged_fragment_and_span.comp:92-93 */
    SLPR_TAP_FLAG_SCAN = 13,      /* int[2*nf+1] plane 6      scan #3, SR.cpp:545-573          */
    SLPR_TAP_RECORDS = 14,        /* int[4*(n_out_frag+n_span)] output_buf, gen_merged_fragment_and_span.comp:77,102 */
    SLPR_TAP_SEGMENTS = 15        /* int[n_paths+1] sort segment table, gen_fragment.comp:226-244 */
};

/* Stage slots reported by slpr_stage_ms(). */
enum {
    SLPR_STAGE_TRANSFORM = 0,   /* k_transform                                  */
    SLPR_STAGE_MONOTONIZE = 1,  /* k_monotonize_count                           */
    SLPR_STAGE_SCAN1 = 2,       /* look-back scan of the curve counts           */
    SLPR_STAGE_INTERSECT = 3,   /* k_piece_emit: set-up of the monotone pieces (MI1:266-308)      */
    SLPR_STAGE_FRAGMENT = 4,    /* k_walk + k_piece_fix: intersection walk + fragment generation  */
    SLPR_STAGE_SORT_HIST = 5,   /* k_segments (+ k_radix_hist + k_radix_hist_scan in radix mode) */
    SLPR_STAGE_SORT_PASSES = 6, /* radix mode: all k_onesweep launches; segmented mode: k_segsort_warp + k_segsort_block */
    SLPR_STAGE_WIND_SCAN = 7,   /* (folded into SPAN_EMIT: always ~0)           */
    SLPR_STAGE_SPAN_EMIT = 8,   /* k_spans: winding scan + mark + flag scan + record emit */
    SLPR_STAGE_FILL_CELLS = 9,  /* k_fill_cells                                 */
    SLPR_STAGE_RESOLVE = 10,    /* k_resolve                                    */
    SLPR_STAGE_COUNT = 11
};

SLPR_API const char *slpr_last_error(void);
SLPR_API const char *slpr_version(void);

/* Replaces ScanlineVGRasterizer::initialize(window,w,h) (scanline_rasterizer.cpp:42-55) and the
 * Vulkan bring-up behind it (vulkan/vk_vg_rasterizer.cpp:16-94). Headless: there is no window.
 * Returns NULL on failure (no CUDA device, bad size, unsupported flag). */
SLPR_API slpr_ctx *slpr_create(int device, uint32_t width, uint32_t height, uint32_t flags);
SLPR_API void slpr_destroy(slpr_ctx *ctx);

/* Run all work of this context on a caller-owned cudaStream_t (e.g. a torch.cuda.Stream) instead of
 * the context's own stream. NULL goes back to the internal stream (so the legacy default stream,
 * whose handle is NULL, cannot be borrowed: create a real stream). */
SLPR_API int slpr_set_stream(slpr_ctx *ctx, void *cuda_stream);

/* Replaces the 7 staging uploads at the end of ScanlineVGRasterizer::loadVG
 * (scanline_rasterizer.cpp:120-146); the arguments are exactly loadVG's flat arrays
 * (scanline/vk_vg_data.h:14-35). Host pointers; copied before return. */
SLPR_API int slpr_load_scene(slpr_ctx *ctx,
                             const float *pos_xy, const uint32_t *pos_path, uint32_t n_points,
                             const uint32_t *curve_pos_map, const uint32_t *curve_type,
                             const uint32_t *curve_path, uint32_t n_curves,
                             const uint32_t *fill_rule, const uint32_t *fill_rgba8, uint32_t n_paths);

/* SLPR_FLAG_FULL_RVG only: the weight of the middle control point of every ARC curve (rational quadratic with weights
 * (1, w, 1), Euclidean control point; entries of other curves are ignored). float[n_curves], host pointer, copied.
 * All weights are 1 until this is called. */
SLPR_API int slpr_set_curve_weights(slpr_ctx *ctx, const float *curve_weight, uint32_t n_curves);

/* Replaces ScanlineVGRasterizer::setMVP (scanline_rasterizer.cpp:191-197): rows m0..m3 of
 * TransPosIn (scanline/compute_ubo.h:9-13), i.e. x' = dot((x,y,0,1), m0) / dot((x,y,0,1), m3). */
SLPR_API int slpr_set_mvp(slpr_ctx *ctx, const float rows[16]);

/* New (multi-GPU row bands): render only scanline rows [y_begin, y_end), both even.
 * (0, height) restores the full frame. Rows are the reference's scanline rows (y up). */
SLPR_API int slpr_set_band(slpr_ctx *ctx, uint32_t y_begin, uint32_t y_end);

/* Render into a caller-owned device buffer (RGBA8, top-left origin, `stride_bytes` per image
 * row) instead of the context's framebuffer — used to write bands straight into a peer-mapped
 * frame on GPU0. NULL restores the internal framebuffer. */
SLPR_API int slpr_set_target(slpr_ctx *ctx, void *dev_rgba, size_t stride_bytes);

/* Replaces ScanlineVGRasterizer::render()/drawFrame() (scanline_rasterizer.cpp:57-65,282-696):
 * the whole frame is enqueued asynchronously on the context's stream, with no host round trip. */
SLPR_API int slpr_render(slpr_ctx *ctx);

/* Optional: do everything a frame needs that can block the device (buffer sizing, allocations, graph capture and
 * upload for the current target) without rendering; slpr_render otherwise does it on first use. */
SLPR_API int slpr_prepare(slpr_ctx *ctx);

/* Headless replacement for acquire/present (vulkan/vk_vg_rasterizer.cpp:424-447): waits for the
 * frame and copies the RGBA8 image (top-left origin, bytes R,G,B,A) to host memory. */
SLPR_API int slpr_readback(slpr_ctx *ctx, uint8_t *rgba, size_t stride_bytes);

/* set_mvp + render + readback in one call: the end-to-end path bench.py times as `e2e`. */
SLPR_API int slpr_render_to_host(slpr_ctx *ctx, const float rows[16], uint8_t *rgba, size_t stride_bytes);

/* Pipelined variant of slpr_render_to_host for frame sequences: returns as soon as the frame and
 * its device-to-host copy are enqueued (two framebuffers, a second stream for the copies), so the
 * copy of frame i overlaps the rendering of frame i+1. `rgba` (pinned memory for a truly
 * asynchronous copy) holds the frame after slpr_wait_host(); use two host buffers alternately. */
SLPR_API int slpr_submit_to_host(slpr_ctx *ctx, const float rows[16], uint8_t *rgba, size_t stride_bytes);
SLPR_API int slpr_wait_host(slpr_ctx *ctx);
/* A frame of a sequence can outgrow the fragment buffers sized from earlier frames (or a path the on-chip sort);
 * that is only known once it ran. The pipelined path checks every frame's device counters when its slot comes
 * round again and in slpr_wait_host, and renders such a frame again into the caller's buffer (the reference
 * sizes its buffers with a host read-back in the middle of every frame instead, scanline_rasterizer.cpp:355-369).
 * Returns how many frames had to be rendered twice so far. */
SLPR_API uint64_t slpr_pipeline_redone(slpr_ctx *ctx);

/* Pinned host memory for the frames of slpr_submit_to_host / slpr_readback, placed on the NUMA node the context's
 * GPU is attached to (mbind; *numa_node receives the node, or -1 if the placement could not be applied — the
 * memory is pinned either way; -2: mbind was refused). With one rank per GPU on a two-socket host this keeps every GPU's device-to-host
 * traffic off the inter-socket link. Freed by slpr_host_free or with the context. */
SLPR_API int slpr_host_alloc(slpr_ctx *ctx, size_t bytes, void **host_ptr, int *numa_node);
SLPR_API int slpr_host_free(slpr_ctx *ctx, void *host_ptr);

/* Stage 5 alone — the reference's draw call over output_buf (scanline_rasterizer.cpp:611-656; scanlinepr.vert:19-46,
 * scanlinepr.frag): clears the frame to white and draws `n_records` draw records given in the reference's own format,
 * four int32 each (yx = y << 16 | x, width, fill_info, frag_index — the rows of workdir/test_data*.csv), later records
 * over earlier ones. Blocking; the frame is then read with slpr_readback / slpr_framebuffer. Records must lie on the
 * 2 x 2 fragment grid (x, y, width even — true of every record the path emits). Needs no scene. Not with SLPR_FLAG_AA4,
 * SLPR_FLAG_BLEND or a band. New entry point: lets the reference's record dumps be drawn and compared directly. */
SLPR_API int slpr_draw_records(slpr_ctx *ctx, const int32_t *records, uint64_t n_records);

/* Device pointer of the last rendered frame (RGBA8) and its row stride. Does not synchronise. */
SLPR_API int slpr_framebuffer(slpr_ctx *ctx, void **dev_rgba, size_t *stride_bytes);

/* Waits for the frame; the three host read-backs of drawFrame (scanline_rasterizer.cpp:356,578-580). */
SLPR_API int slpr_get_counts(slpr_ctx *ctx, uint32_t *n_fragments, uint32_t *n_out_fragments, uint32_t *n_spans);

/* Waits for the frame and copies one intermediate buffer (SLPR_TAP_*) to host memory. Needs
 * SLPR_FLAG_TAPS for the planes the fast path does not materialise. `bytes` must not exceed
 * the buffer's size for the last frame. Replaces drawDebug() (scanline_rasterizer.cpp:698-802). */
SLPR_API int slpr_debug_copy(slpr_ctx *ctx, int which, void *dst, size_t bytes);

/* Per-stage GPU milliseconds of the last frame rendered with direct launches (SLPR_FLAG_NO_GRAPH);
 * fills min(n, SLPR_STAGE_COUNT) slots. New: the reference has no timers. */
SLPR_API int slpr_stage_ms(slpr_ctx *ctx, float *ms, int n);

/* ---- Exact row bands across GPUs (new; csrc/bands.cuh) --------------------------------
 * The reference's winding scan is one unsegmented prefix sum over all fragments (SR.cpp:479-506), so a band of
 * rows needs the winding sums of the other bands' fragments that sort before its own. Per frame and per band:
 *   slpr_render_band_begin   transform .. fragments of this band, then its sort; leaves 3 * n_paths int32
 *                            (per-path sums of the winding deltas: normal rows | outside the frame | row 0) in
 *                            dev_sums and returns as soon as THEY are complete (capacity growth and sort
 *                            selection are settled here); the sort is still running on the context's stream
 *   (caller)                 all-gather of dev_sums over the bands into dev_gathered [n_bands][3 * n_paths],
 *                            bands ordered by rows — on another stream, so that it overlaps the sort; the
 *                            context's stream must wait for it before the next call (NCCL all-gather)
 *   slpr_render_band_end     corrections, spans, draw records, pixels of the band (asynchronous, like slpr_render)
 * slpr_set_band first; slpr_set_band_exchange(ctx, NULL, NULL, 0, 0) returns to independent bands, which are
 * exact only for scenes whose per-path winding sums vanish. */
SLPR_API int slpr_band_exchange_ints(slpr_ctx *ctx, size_t *ints_per_band);
SLPR_API int slpr_set_band_exchange(slpr_ctx *ctx, int32_t *dev_sums, const int32_t *dev_gathered, int n_bands, int band);
SLPR_API int slpr_render_band_begin(slpr_ctx *ctx);
SLPR_API int slpr_render_band_end(slpr_ctx *ctx);

/* ---- Exact row bands, device-side exchange (new in round 2; csrc/bands.cuh second half) -----------------------
 * No host hand-over and no collective: every (path, row) of a closed path has winding deltas that cancel, so only
 * the few paths with a residue carry information. A band stores its non-zero per-path sums straight into a mailbox
 * in every other band's HBM (peer-mapped over NVLink; CUDA IPC between one-process-per-GPU ranks), followed by a
 * system-scope flag; every band merges what it received into a sorted break-point table that k_spans consults.
 * The whole band frame is ONE captured graph, like slpr_render. Set-up, once per context:
 *   slpr_band_mailbox      allocates this band's mailbox (plain cudaMalloc) and returns its device pointer
 *   slpr_ipc_export/import turn it into a 64-byte CUDA IPC handle / map another process's allocation on this
 *                          device (peer access is enabled on demand). Same-process contexts pass pointers directly.
 *   slpr_set_band_peers    mailboxes[n_bands] as addressable from this device, bands ordered by rows; `root` is the
 *                          band that collects the frame (its mailbox also receives the "pixels are in place" flags)
 * Per frame, on every band, the same increasing frame_seq:
 *   slpr_set_band / slpr_set_target (e.g. the root's frame buffer, allocated with slpr_alloc_device and mapped
 *   through IPC: k_resolve then stores this band's pixels straight into the root's HBM), slpr_render_band(seq);
 *   on the root slpr_band_wait_gather(seq) orders its stream after the arrival of every band's pixels.
 * slpr_synchronize (or any call that finishes the frame) returns SLPR_ERR_RETRY when the frame was void on some
 * band (a band outgrew its buffers, met a path too long for the segmented sort, or had more than 2048 residue
 * paths): the state that caused it is already adjusted; render the frame again on every band with a new frame_seq.
 * A band that never publishes makes the others give up after 2 s (SLPR_ERR_STATE) instead of hanging. */
SLPR_API int slpr_band_mailbox(slpr_ctx *ctx, void **dev_ptr, size_t *bytes);
SLPR_API int slpr_alloc_device(slpr_ctx *ctx, size_t bytes, void **dev_ptr); /* cudaMalloc, freed with the context */
SLPR_API int slpr_ipc_export(slpr_ctx *ctx, void *dev_ptr, unsigned char handle[64]);
SLPR_API int slpr_ipc_import(slpr_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
SLPR_API int slpr_set_band_peers(slpr_ctx *ctx, int n_bands, int band, int root, void *const *mailboxes);
SLPR_API int slpr_render_band(slpr_ctx *ctx, uint32_t frame_seq);
SLPR_API int slpr_band_wait_gather(slpr_ctx *ctx, uint32_t frame_seq);
/* Gather by copy engine instead of direct stores (slpr_set_band_peers with root = -1 switches the in-graph
 * "pixels are in place" flag off): the band just rendered into the current target travels to dst_frame — the
 * root's peer-mapped frame buffer, addressed like a full frame — on a second stream, overlapping the next frame's
 * kernels and using no SM (896 MB per 16K frame into one GPU's NVLink ingress take longer than a band's kernels);
 * the flag for the root follows the copy. A later slpr_render_band into the same slot (frame_seq & 1) waits for it. */
SLPR_API int slpr_band_push(slpr_ctx *ctx, uint32_t frame_seq, int root_band, void *dst_frame, size_t dst_stride);
/* Number of differing 32-bit words (RGBA8 pixels) between two device buffers (checks an assembled frame against
 * the same frame rendered whole, on the device). Waits for the context's stream. */
SLPR_API int slpr_debug_diff_u32(slpr_ctx *ctx, const void *dev_a, const void *dev_b, size_t n_words, uint64_t *n_diff);

/* Sort geometry of the last frame: key bits, number of 8-bit radix passes, bytes of one key. */
SLPR_API int slpr_sort_info(slpr_ctx *ctx, uint32_t *key_bits, uint32_t *passes, uint32_t *key_bytes);
/* Which sort the context uses: 0 = segmented one-pass sort (csrc/segsort.cuh), 1 = onesweep radix
 * sort (SLPR_FLAG_RADIX_SORT, or a path had more than 4096 fragments). */
SLPR_API int slpr_sort_mode(slpr_ctx *ctx, int *mode);
/* Where the cells of the draw records are marked (stage 5, scanlinepr.vert/.frag): 1 = inside the span kernel
 * (big frames), 0 = by a separate pass over the records (small frames). Chosen per scene and view from the
 * fragment count; the pixels are the same either way. */
SLPR_API int slpr_fill_mode(slpr_ctx *ctx, int *fused);
/* Long pieces (62 or more grid crossings) of the last frame, and whether the next frame walks them chain by chain
 * (csrc/walk.cuh: k_long_chains / k_long_emit; chosen per scene and view from this count, SLPR_FLAG_NO_LONG_WALK). */
SLPR_API int slpr_long_walk_info(slpr_ctx *ctx, int *mode_on, uint32_t *n_long_pieces);
/* Monotone pieces walked by the last frame (64-byte piece records read by k_walk). */
SLPR_API int slpr_walk_info(slpr_ctx *ctx, uint32_t *n_pieces);

/* Stand-alone device primitives over caller-owned DEVICE buffers (the two roofline-graded
 * kernels, exposed for micro-benchmarks and property tests).
 * slpr_scan_i32: out[i] = sum_{j<i} in[j] for i in [0,n] (naive_scan.comp:27-73 semantics).
 * slpr_sort_pairs: stable LSD radix sort of (u64 key, u32 value) on key bits [0,key_bits)
 * (replaces naive_seg_sort_pairs.comp:26-97). Both run on the context's stream, asynchronously. */
SLPR_API int slpr_scan_i32(slpr_ctx *ctx, const int32_t *dev_in, int32_t *dev_out, uint64_t n);
SLPR_API int slpr_sort_pairs(slpr_ctx *ctx, uint64_t *dev_keys, uint32_t *dev_vals,
                             uint64_t *dev_keys_tmp, uint32_t *dev_vals_tmp, uint64_t n,
                             uint32_t key_bits, int *result_in_tmp);
SLPR_API int slpr_synchronize(slpr_ctx *ctx);
/* Number of kernels this library has launched on this context so far (graph nodes count). */
SLPR_API uint64_t slpr_launch_count(slpr_ctx *ctx);

/* ---------------------------------------------------------------------------------------------
 * Host scene front end: RVG text -> VGContainer -> the 7 flat arrays.
 * Replaces RVG::load (core/vg/rvg.cpp:9-255) and the flattening half of loadVG
 * (scanline_rasterizer.cpp:67-118). Behaviour follows the reference parser, quirks included.
 * ------------------------------------------------------------------------------------------ */
typedef struct slpr_vg slpr_vg;

typedef struct slpr_scene_view { /* pointers owned by the slpr_vg; valid until slpr_vg_free */
    const float *pos_xy; const uint32_t *pos_path; uint32_t n_points;
    const uint32_t *curve_pos_map, *curve_type, *curve_path; uint32_t n_curves;
    const uint32_t *fill_rule, *fill_rgba8; uint32_t n_paths;
    float viewport[4], window[4];
    const float *curve_weight; /* float[n_curves] from slpr_vg_load_rvg_full (middle weight of ARC curves), else NULL */
} slpr_scene_view;

SLPR_API slpr_vg *slpr_vg_load_rvg(const char *path);
/* a complete reader (rational arcs `A`, `Q`, `H`/`V`, relative commands, per-element transforms, every contour
 * closed, gradient paints as their average colour) instead of the reference parser's behaviour. Its QUADRIC / ARC
 * curves need a context created with SLPR_FLAG_FULL_RVG and slpr_set_curve_weights(view.curve_weight). */
SLPR_API slpr_vg *slpr_vg_load_rvg_full(const char *path);
/* Build a container from VGContainer-shaped arrays (vg_container.h:21-87). */
SLPR_API slpr_vg *slpr_vg_from_arrays(const float *pos_xy, uint32_t n_points,
                                      const uint32_t *curve_pos, const uint32_t *curve_type, uint32_t n_curves,
                                      const uint32_t *path_curve, const uint32_t *fill_rule,
                                      const float *fill_color_rgba, const float *fill_opacity, uint32_t n_paths);
SLPR_API int slpr_vg_flatten(slpr_vg *vg, slpr_scene_view *out);
/* Raw container arrays, for comparing the parser with the reference's (tests). */
SLPR_API int slpr_vg_container(slpr_vg *vg, const float **pos_xy, uint32_t *n_points,
                               const uint32_t **curve_pos, const uint32_t **curve_type, uint32_t *n_curves,
                               const uint32_t **path_curve, const uint32_t **fill_rule,
                               const float **fill_color, const float **fill_opacity, uint32_t *n_paths);
SLPR_API void slpr_vg_free(slpr_vg *vg);

#ifdef __cplusplus
}
#endif
#endif /* SLPR_H_ */
