/*
 * oracle.h — CPU restatement of the VkScanlinePR compute shaders (TEST INFRASTRUCTURE ONLY).
 *
 * This library is the checker, never the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it. The shipped renderer
 * (vkscanlinepr_b200/csrc, libslpr.so) never links or calls anything in oracle/.
 *
 * PARITY STATUS: pinned against the reference's own shipped binaries, by interpretation. The
 * reference implements this path only as GLSL/SPIR-V (no CPU code, no tests, no golden outputs for
 * its scenes) and no Vulkan loader/ICD exists in the build container or on the GPU box, so the
 * shaders cannot run on a driver. Instead oracle/spirv_exec.py (a SPIR-V interpreter written for
 * this purpose) executes the prebuilt workdir/shaders/<dir>/spv/<name>.comp.spv along drawFrame's dispatch
 * sequence (tools/make_spirv_golden.py); every buffer of those frames is committed under
 * tests/golden/spirv_*.npz and this restatement reproduces all of them bit for bit
 * (tests/test_spirv_golden.py: 5 scenes incl. a 4-cut cubic, QUADRIC-typed curves, clipping on
 * all edges, a tiger subset with a winding residue). Caveats: the interpreter evaluates fp32
 * without FMA contraction and sums OpDot left to right (a real driver may differ),
 * and the fixed-function line raster of stage 5 is restated from the Vulkan rules, not executed —
 * its vertex and fragment shaders ARE executed (tools/make_stage5_golden.py, on the reference's own
 * record dumps; tests/test_stage5_golden.py: orc_fill equals their output rasterized by those rules).
 * Also pinned: (1) scene flattening (loadVG + RVG parser) against the reference's own parser
 * compiled from its sources (oracle/_ref, see oracle/Makefile); (2) the output record format /
 * ordering invariants against workdir/test_data.csv and test_data3.csv.
 *
 * All citations are file:line relative to /root/reference/.
 *   TP   = workdir/shaders/scanline/compute/transform_pos.comp
 *   MI0  = workdir/shaders/scanline/compute/make_intersection_0.comp
 *   MI1  = workdir/shaders/scanline/compute/make_intersection_1.comp
 *   GF   = workdir/shaders/scanline/compute/gen_fragment.comp
 *   SHUF = workdir/shaders/scanline/compute/shuffle_fragment.comp
 *   MARK = workdir/shaders/scanline/compute/mark_merged_fragment_and_span.comp
 *   GEN  = workdir/shaders/scanline/compute/gen_merged_fragment_and_span.comp
 *   SCAN = workdir/shaders/common/naive_scan.comp
 *   SORT = workdir/shaders/common/naive_seg_sort_pairs.comp
 *   VERT = workdir/shaders/scanline/surface/scanlinepr.vert
 *   SR   = VkScanlinePR/src/core/scanline/scanline_rasterizer.cpp
 *
 * Arithmetic policy: IEEE fp32, no FMA contraction (compile with -ffp-contract=off),
 * correctly rounded sqrt and division, float->int by truncation with saturation.
 */
#ifndef SLPR_ORACLE_H_
#define SLPR_ORACLE_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TP:33-85. tpos_out: float[2*n_points]; path_visible: int32[n_paths], OR-accumulated
 * (caller zeroes it per frame; race-free OR). rows = m0..m3 (16 floats). */
void orc_transform(uint32_t n_points, const float *pos, const uint32_t *pos_path,
                   const float *rows, int width, int height,
                   float *tpos_out, int32_t *path_visible);

/* MI0:226-410. cut_cache_out: float[5*n_curves] (slot 4 = bits of n_cuts; slots >= n_cuts
 * are written as 0 here, uninitialised in the reference). count_out: int32[n_curves]. */
void orc_monotonize_count(uint32_t n_curves, const uint32_t *curve_type,
                          const uint32_t *curve_pos_map, const uint32_t *curve_path,
                          const float *tpos, const int32_t *path_visible,
                          int width, int height, float *cut_cache_out, int32_t *count_out);

/* SCAN:27-73 semantics: out[i] = sum_{j<i} in[j], i in [0,n]; in-place allowed. */
void orc_exclusive_scan(int64_t n, const int32_t *in, int32_t *out);

/* MI1:217-447. offsets = scanned counts (n_curves+1). inter_out: int32[2*nf] as (curve, tbits). */
void orc_intersect(uint32_t n_curves, const uint32_t *curve_type,
                   const uint32_t *curve_pos_map, const uint32_t *curve_path,
                   const float *tpos, const int32_t *path_visible,
                   const float *cut_cache, const int32_t *offsets,
                   int width, int height, int32_t *inter_out);

/* GF:90-246. key/idx/path/wind: int32[nf] (planes 0,1,2,4); seg: int32[n_paths+1] (plane 3). */
void orc_gen_fragment(int32_t nf, uint32_t n_paths, const int32_t *inter,
                      const uint32_t *curve_path, const uint32_t *curve_pos_map,
                      const uint32_t *curve_type, const float *tpos,
                      int width, int height,
                      int32_t *key, int32_t *idx, int32_t *path, int32_t *wind, int32_t *seg);

/* SORT:26-97 result: each [seg[p],seg[p+1]) ascending by signed (key, idx). In place.
 * (The reference's O(n^2) odd-even transposition is replaced by qsort on the same total
 * order; the result is identical by construction.) */
void orc_seg_sort(uint32_t n_paths, const int32_t *seg, int32_t *key, int32_t *idx);

/* The literal O(n^2) odd-even transposition network of SORT:53-97 for small cross-checks. */
void orc_seg_sort_literal(uint32_t n_paths, const int32_t *seg, int32_t *key, int32_t *idx);

/* SHUF:17-27: out[i] = wind[idx[i]]. */
void orc_shuffle(int32_t nf, const int32_t *idx, const int32_t *wind, int32_t *out);

/* MARK:22-94. flags_out: int32[2*nf] = [frag flags | span flags]. wn = scanned winding. */
void orc_mark(int32_t nf, const int32_t *key, const int32_t *path, const int32_t *wn,
              const uint32_t *fill_rule, int width, int height, int32_t *flags_out);

/* GEN:33-103. scan3: int32[2*nf+1] exclusive scan of flags. out: int32[4*(n_out_frag+n_span)]. */
void orc_emit(int32_t nf, const int32_t *key, const int32_t *path, const int32_t *flags,
              const int32_t *scan3, const uint32_t *fill_info, int32_t n_out_frag,
              int32_t *out);

/* VERT:19-46 + SR:611-656,883-895: records -> RGBA8, top-left origin,
 * image row = H-1-scanline row, clear white, opaque, later record wins. rgba: uint8[4*W*H]. */
void orc_fill(int64_t n_records, const int32_t *records, int width, int height, uint8_t *rgba);

/* Whole frame, all intermediates kept. Caller frees with orc_frame_free. */
typedef struct orc_frame {
    int32_t n_fragments, n_out_frag, n_span;
    uint32_t n_points, n_curves, n_paths;
    int width, height;
    float   *tpos;          /* 2*n_points */
    int32_t *path_visible;  /* n_paths */
    float   *cut_cache;     /* 5*n_curves */
    int32_t *curve_count;   /* n_curves */
    int32_t *curve_offset;  /* n_curves+1 */
    int32_t *inter;         /* 2*nf */
    int32_t *key, *idx, *path, *wind, *seg; /* unsorted planes; seg n_paths+1 */
    int32_t *skey, *sidx;   /* sorted */
    int32_t *swind;         /* shuffled deltas */
    int32_t *wn;            /* nf+1 */
    int32_t *flags;         /* 2*nf */
    int32_t *scan3;         /* 2*nf+1 */
    int32_t *records;       /* 4*(n_out_frag+n_span) */
    uint8_t *rgba;          /* 4*W*H (NULL if do_fill==0) */
    double   ms[8];         /* per-stage wall ms: tp, mi0, scan1, mi1, gf, sort, span(shuf..emit), fill */
} orc_frame;

orc_frame *orc_render(uint32_t n_points, const float *pos, const uint32_t *pos_path,
                      uint32_t n_curves, const uint32_t *curve_pos_map,
                      const uint32_t *curve_type, const uint32_t *curve_path,
                      uint32_t n_paths, const uint32_t *fill_rule, const uint32_t *fill_info,
                      const float *rows, int width, int height, int do_fill);
void orc_frame_free(orc_frame *f);

/* Host flattening of SR.cpp:67-171 colour quantisation: rgba float[4], opacity -> RGBA8 word. */
uint32_t orc_quantise_colour(const float *rgba, float opacity);

/* SURVEY section 8 f-1, off by default: real QUADRIC / ARC arithmetic in place of the reference's TODO arms (definition in
 * oracle.c next to interp_general). curve_weight: float[n_curves], the middle weight of ARC curves (others ignored);
 * the pointer must stay valid while the mode is on. Process-wide. */
void orc_set_full_rvg(int on, const float *curve_weight);

/* SURVEY App. D.1, off by default: evaluate the shaders with a*b + c contracted into fused multiply-adds (policy in oracle.c
 * next to lerpf). NOT pinned by anything of the reference's: no Vulkan driver here to say what it contracts. Process-wide. */
void orc_set_contract_fma(int on);

/* SURVEY section 8 f-3, off by default: stage 5 composites the records "source over" in path order (integer arithmetic defined in
 * oracle.c next to orc_fill) instead of the reference's opaque overwrite. Process-wide. */
void orc_set_blend(int on);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
