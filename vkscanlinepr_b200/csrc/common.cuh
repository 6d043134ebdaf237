// common.cuh — shared device helpers and frame-wide structures for libslpr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slpr {

constexpr int FRAG_SIZE = 2;
constexpr uint32_t T_LINE = 0x02u, T_QUADRIC = 0x03u, T_CUBIC = 0x04u, T_ARC = 0x13u;
constexpr int CUBIC_ITERATION_NUMBER = 24;  // make_intersection_1.comp:7
constexpr int NUM_SMS_B200 = 148;

// Everything a frame's kernels need that changes between frames or is only known on the device.
// Lives in device memory so that a captured CUDA graph can be replayed with new matrices/counts.
struct FrameParams {
    float rows[16];  // m0..m3 of TransPosIn (compute_ubo.h:9-13)
    int width, height;
    int band_y0, band_y1;  // scanline rows [y0,y1) this context renders
    int cull;              // 1 when the band is a strict subset of the frame
    int stat_mid;      // paths of 129..512 fragments   } k_path_segments, every frame and in both sort modes:
    int stat_big;      // paths of 513..4096 fragments  } the host picks the sort from them
    int stat_huge;     // paths of more than 4096 fragments (radix sort only)
    int n_live;        // band mode: curves whose path comes near the band (k_band_live)
    int frame_seq;     // exact bands with the device-side exchange: which frame this is (mailbox slot = seq & 1)
    int pad[2];
};

// Device-resident counters (zeroed at the start of every frame).
struct FrameCounters {
    int n_fragments;   // total of scan #1 (SR.cpp:356)
    int n_out_frag;    // SR.cpp:578
    int n_span;        // SR.cpp:579-580
    int overflow;      // set when n_fragments exceeded the allocated capacity
    int n_records;     // n_out_frag + n_span
    int wn_total;      // last element of the winding scan (residue diagnostic)
    int sort_fallback; // segmented sort met a path with more than SEG_BLOCK_MAX fragments
    int n_big_segments;  // paths queued for k_segsort_block
    int n_pieces;      // monotone pieces walked this frame (k_piece_emit)
    int stat_mid;      // paths of 129..512 fragments   } k_path_segments, every frame and in both sort modes:
    int stat_big;      // paths of 513..4096 fragments  } the host picks the sort from them
    int stat_huge;     // paths of more than 4096 fragments (radix sort only)
    int n_live;        // band mode: curves whose path comes near the band (k_band_live)
    int n_fix;         // pieces whose predecessor's boundary fragment must be redone (k_walk -> k_piece_fix)
    int fix_missed;    // invariant check of k_walk's conditional boundary stores (always 0)
    int n_band_entries;  // exact bands: paths of this band with a non-zero winding sum (k_band_sums_sparse)
    int n_band_bp;       // exact bands: break points in the merged correction table (k_band_merge)
    int n_long;          // the longest monotone pieces, laid out first (k_piece_emit: 62 records or more, and shorter ones while few): walked chain by chain when long_mode is on
    int n_live_paths;    // band mode: paths that can reach the band (k_band_paths / k_path_cull), listed for the sort
    int band_void;       // exact bands: 1 = another band's frame was void, 2 = a band never published (timeout),
                         //              3 = more entries than the merge table holds; the frame must be rendered again
    int n_top;           // pieces in the top length bucket (62 records or more): the first n_top of the record array
    int n_blend_nodes;   // SLPR_FLAG_BLEND: (cell, translucent path) pairs asked for this frame; more than the node buffer
                         //              holds = the frame lacks some and is rendered again with a bigger buffer
};

// SLPR_FLAG_BLEND (beyond the reference, which overwrites: blendEnable = VK_FALSE, SR.cpp:893-895): fills with
// 0 < alpha < 255 are composited "source over" in path order. Opaque paths keep marking the coverage grid with
// atomicMax (the top-most opaque path hides everything below it); a translucent path appends one node (its priority,
// next) to the cell's list; the resolve kernel composites, in ascending priority, the nodes above the opaque top.
struct BlendList {
    uint32_t *heads;  // per coverage cell: 0 or node index + 1
    uint2 *nodes;     // (priority = path + 1 or record index + 1, next)
    uint32_t cap;
};

__device__ __forceinline__ void blend_append(const BlendList &bl, size_t cell, uint32_t node, uint32_t prio) {
    if (node < bl.cap) {
        const uint32_t prev = atomicExch(bl.heads + cell, node + 1u);
        bl.nodes[node] = make_uint2(prio, prev);
    }
}

// dst, src: R,G,B,A bytes (low to high). The integer arithmetic of oracle/oracle.c blend_over(); the frame stays opaque.
__device__ __forceinline__ uint32_t blend_over(uint32_t dst, uint32_t src) {
    const uint32_t a = src >> 24;
    if (a == 255u) return src;
    if (a == 0u) return dst;
    uint32_t out = 0xFF000000u;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const uint32_t s = (src >> (8 * ch)) & 0xFFu, d = (dst >> (8 * ch)) & 0xFFu;
        out |= ((s * a + d * (255u - a) + 127u) / 255u) << (8 * ch);
    }
    return out;
}

// The back half of a frame (winding prefix, spans, coverage) must not touch the sorted buffers of a frame the
// host is going to render again: one whose fragments outgrew the buffers (overflow; nothing was generated) or one
// with a path too long for the segmented sort (sort_fallback: the sorted buffers are partly unwritten).
__device__ __forceinline__ bool frame_void(const FrameCounters *ctr, int capacity) {
    return ctr->overflow != 0 || ctr->n_fragments > capacity || ctr->sort_fallback != 0 || ctr->band_void != 0;
}

// Key geometry for the compact 64-bit sort key (path | row rank | cell x), see DESIGN.md.
struct KeyLayout {
    int bits_x, bits_y, bits_path;
    int ny;  // number of cell rows = (H+1)/2
};

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
// GLSL int(x): truncation; cvt.rzi saturates and maps NaN to 0 (the oracle pins the same).
__device__ __forceinline__ int f2i(float x) { return __float2int_rz(x); }

// LERP(a,b,t) = a + t*(b-a) with three separately rounded operations (no FMA contraction);
// make_intersection_0.comp:10, make_intersection_1.comp:12, gen_fragment.comp:10.
__device__ __forceinline__ float lerpf(float a, float b, float t) {
    return __fadd_rn(a, __fmul_rn(t, __fsub_rn(b, a)));
}

// ---- SLPR_FLAG_CONTRACT_FMA: the other legitimate reading of the shaders. GLSL lets a compiler
// contract a*b + c into one fused multiply-add unless the expression is `precise`; whether a given Vulkan driver does
// is not pinned by anything in the reference. Policy of the contracted mode, defined in oracle/oracle.c
// (orc_set_contract_fma) and mirrored here: (1) LERP(a,b,t) = fma(t, b - a, a); (2) dot(vec4(x,y,0,1), m) accumulated
// left to right, x*m.x then one fma per further term; (3) make_intersection_0's a = fma(3, x1 - x2, x3 - x0) and the
// discriminant fma(B, B, -(A*C)); (4) the f-1 arc evaluator's numerator as an fma chain. Nothing else has the a*b + c
// shape. FMA = false is the default, pinned by the executed SPIR-V fixtures.
template <bool FMA>
__device__ __forceinline__ float madd_t(float a, float b, float c) {  // a * b + c
    return FMA ? __fmaf_rn(a, b, c) : __fadd_rn(__fmul_rn(a, b), c);
}
template <bool FMA>
__device__ __forceinline__ float lerp_t(float a, float b, float t) {
    return FMA ? __fmaf_rn(t, __fsub_rn(b, a), a) : __fadd_rn(a, __fmul_rn(t, __fsub_rn(b, a)));
}

// make_intersection_1.comp:76-79, gen_fragment.comp:54-57
__device__ __forceinline__ int float2int_rd(float x) { return x >= 0.0f ? f2i(x) : f2i(__fsub_rn(x, 1.0f)); }

// make_intersection_0.comp:14-20
__device__ __forceinline__ bool path_invisible(int mask) {
    uint32_t m = (uint32_t)mask;
    return ((m & 0x11111000u) == 0) || ((m & 0x01101011u) == 0) || ((m & 0x00011111u) == 0) ||
           ((m & 0x11010110u) == 0);
}

template <bool FMA = false>
__device__ __forceinline__ float cubic_eval(float p0, float p1, float p2, float p3, float t) {
    float q0 = lerp_t<FMA>(p0, p1, t), q1 = lerp_t<FMA>(p1, p2, t), q2 = lerp_t<FMA>(p2, p3, t);
    float l0 = lerp_t<FMA>(q0, q1, t), l1 = lerp_t<FMA>(q1, q2, t);
    return lerp_t<FMA>(l0, l1, t);
}

// interpolateGeneralCurve: make_intersection_0.comp:131-165 (dflt 1.0f), make_intersection_1.comp:81-148 (dflt 0.0f)
__device__ __forceinline__ float interp_general(uint32_t type, float t, float p0, float p1, float p2, float p3,
                                                float dflt) {
    if (type == T_LINE) return lerpf(p0, p1, t);
    if (type == T_CUBIC) return cubic_eval<false>(p0, p1, p2, p3, t);
    return dflt;
}

// ---- SURVEY section 8 f-1 (SLPR_FLAG_FULL_RVG; NOT in the reference, whose QUADRIC / ARC arms are `// TODO`): the arithmetic is
// defined in oracle/oracle.c next to interp_general and mirrored here operation for operation.
template <bool FMA = false>
__device__ __forceinline__ float eval_quadric(float p0, float p1, float p2, float t) {
    const float q0 = lerp_t<FMA>(p0, p1, t), q1 = lerp_t<FMA>(p1, p2, t);
    return lerp_t<FMA>(q0, q1, t);
}
// rational quadratic with weights (1, w, 1), Euclidean control point
template <bool FMA = false>
__device__ __forceinline__ float eval_arc(float p0, float p1, float p2, float w, float t) {
    const float u = __fsub_rn(1.0f, t);
    const float b0 = __fmul_rn(u, u), b2 = __fmul_rn(t, t);
    const float b1 = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, t), u), w);
    const float D = __fadd_rn(__fadd_rn(b0, b1), b2);
    const float N = madd_t<FMA>(b2, p2, madd_t<FMA>(b1, p1, __fmul_rn(b0, p0)));
    return __fdiv_rn(N, D);
}
// interp_general with the TODO arms filled in (p3 carries the weight of an ARC)
template <bool FMA = false>
__device__ __forceinline__ float interp_full(uint32_t type, float t, float p0, float p1, float p2, float p3, float dflt, bool full) {
    if (type == T_LINE) return lerp_t<FMA>(p0, p1, t);
    if (type == T_CUBIC) return cubic_eval<FMA>(p0, p1, p2, p3, t);
    if (full && type == T_QUADRIC) return eval_quadric<FMA>(p0, p1, p2, t);
    if (full && type == T_ARC) return eval_arc<FMA>(p0, p1, p2, p3, t);
    return dflt;
}

// streaming 128-bit accesses (bypass L1 allocation for data touched once; not .nc so that
// in-place use is well defined)
__device__ __forceinline__ int4 ld_stream(const int4 *p) {
    int4 r;
    asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(int4 *p, const int4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w) : "memory");
}

// Tile-status words of the decoupled look-backs: GPU-scope relaxed accesses (a `volatile` access
// compiles to .STRONG.SYS, i.e. system scope, which is needlessly expensive on a two-die part).
__device__ __forceinline__ unsigned long long ld_status64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_status32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace slpr
