// spans.cuh — everything after the sort, as two streaming look-back passes:
//   pass A  shuffle_fragment.comp:17-27 + scan #2 (SR.cpp:479-506): the winding delta travels in the
//           two top bits of the sorted value, so the gather disappears; the scan is global and
//           unsegmented exactly like the reference's.
//   pass B  mark_merged_fragment_and_span.comp:22-94 + scan #3 (SR.cpp:545-573) +
//           gen_merged_fragment_and_span.comp:33-103: flags are computed on the fly from the sorted
//           keys and the winding scan, both flag counts are scanned in one 2x31-bit packed value,
//           and the draw records are written straight from the scan's store step.
#pragma once
#include "geom.cuh"
#include "scan.cuh"

namespace slpr {

struct WindScanOp {
    using Aux = NoAux;
    static constexpr int VECS = 4, MIN_BLOCKS = 4;
    const uint32_t *sval;  // sorted values: index | (delta+1) << 30
    int *wn;               // [nf+1] exclusive winding scan (plane 3 after scan #2)
    int *sidx_tap;         // optional: plane 1 after sort
    FrameCounters *ctr;
    int capacity;
    __device__ long long count() const {
        const int nf = ctr->n_fragments;
        return nf > capacity ? -1 : nf;
    }
    __device__ void load(long long i, long long n, unsigned long long x[4], Aux &) const {
        uint32_t v[4];
        if (i + 3 < n) {
            const int4 q = ld_stream(reinterpret_cast<const int4 *>(sval + i));
            v[0] = (uint32_t)q.x; v[1] = (uint32_t)q.y; v[2] = (uint32_t)q.z; v[3] = (uint32_t)q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (i + k < n) ? sval[i + k] : (1u << 30);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = (uint32_t)((int)(v[k] >> 30) - 1);  // zero-extended int32 delta
        if (sidx_tap) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k < n) sidx_tap[i + k] = (int)(v[k] & 0x3FFFFFFFu);
        }
    }
    __device__ void store(long long i, long long n, const unsigned long long e[4], const unsigned long long *,
                          const Aux &) const {
        if (i + 3 < n) {
            st_stream(reinterpret_cast<int4 *>(wn + i), make_int4((int)e[0], (int)e[1], (int)e[2], (int)e[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k < n) wn[i + k] = (int)e[k];
        }
    }
    __device__ void finish(long long n, unsigned long long total) const {
        wn[n] = (int)total;
        ctr->wn_total = (int)total;
    }
};

struct SpanTaps {
    int *skey32;  // plane 0 after sort
    int *flags;   // [2*nf] = [frag | span]                     (MARK:92-93)
    int *scan3;   // [2*nf+1]; second half needs + n_out_frag   (k_scan3_fixup)
};

struct SpanAux {
    uint32_t xy[4];     // per element: cell x (low 16) | row (high 16), only meaningful when flagged
    uint32_t xprev[4];  // per element: previous fragment's cell x + 2, clamped at 0 (span start)
    uint32_t path[4];
};

struct SpanEmitOp {
    using Aux = SpanAux;
    static constexpr int VECS = 2, MIN_BLOCKS = 2;  // fat per-element state: smaller tile, 2 blocks/SM
    const uint64_t *skey;  // sorted compact keys
    const int *wn;         // exclusive winding scan
    const uint32_t *fill_rule, *fill_info;
    int4 *records;         // output_buf (GEN:77,102)
    FrameCounters *ctr;
    KeyLayout L;
    int width, height;
    int capacity;
    SpanTaps taps;

    __device__ long long count() const {
        const int nf = ctr->n_fragments;
        return nf > capacity ? -1 : nf;
    }

    __device__ __forceinline__ void decode(uint64_t k, uint32_t &path, int &x, int &y, uint32_t &yk) const {
        const uint32_t xk = (uint32_t)(k & ((1ull << L.bits_x) - 1));
        yk = (uint32_t)((k >> L.bits_x) & ((1ull << L.bits_y) - 1));
        path = (uint32_t)(k >> (L.bits_x + L.bits_y));
        if (yk == (uint32_t)(L.ny - 1)) { x = 0x7FFF; y = 0x7FFF; }  // invalid key decodes to (32767, 32767), MARK:41-42
        else { x = (int)xk * 2 - FRAG_SIZE; y = (yk == (uint32_t)L.ny) ? 0 : (int)(yk + 1) * 2; }
    }

    __device__ void load(long long i, long long n, unsigned long long x[4], Aux &a) const {
        uint64_t k[5];  // k[0] = key[i-1]
        k[0] = (i > 0 && i - 1 < n) ? skey[i - 1] : 0ull;
        int w[4];
        if (i + 3 < n) {
            const int4 q0 = ld_stream(reinterpret_cast<const int4 *>(skey + i));
            const int4 q1 = ld_stream(reinterpret_cast<const int4 *>(skey + i + 2));
            k[1] = ((uint64_t)(uint32_t)q0.y << 32) | (uint32_t)q0.x;
            k[2] = ((uint64_t)(uint32_t)q0.w << 32) | (uint32_t)q0.z;
            k[3] = ((uint64_t)(uint32_t)q1.y << 32) | (uint32_t)q1.x;
            k[4] = ((uint64_t)(uint32_t)q1.w << 32) | (uint32_t)q1.z;
            const int4 qw = ld_stream(reinterpret_cast<const int4 *>(wn + i));
            w[0] = qw.x; w[1] = qw.y; w[2] = qw.z; w[3] = qw.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                k[j + 1] = (i + j < n) ? skey[i + j] : 0ull;
                w[j] = (i + j < n) ? wn[i + j] : 0;
            }
        }
        uint32_t p0; int x0, y0; uint32_t yk0;
        decode(k[0], p0, x0, y0, yk0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t p1; int x1, y1; uint32_t yk1;
            decode(k[j + 1], p1, x1, y1, yk1);
            uint32_t frag = 0, span = 0;
            if (i + j < n) {
                const bool oob = (x1 < 0 || y1 < 0 || x1 >= width || y1 >= height);  // MARK:45,69
                if (i + j == 0) {
                    frag = oob ? 0u : 1u;  // MARK:44-51
                } else {
                    frag = (!oob && (p0 != p1 || k[j] != k[j + 1])) ? 1u : 0u;  // MARK:69-77
                    const uint32_t rule = fill_rule[p1];
                    const bool wn_flag = ((rule == 0) && (w[j] != 0)) || ((rule == 1) && ((w[j] & 1) != 0));  // MARK:82
                    span = (y0 == y1 && (x0 + FRAG_SIZE) < x1 && p0 == p1 && wn_flag) ? 1u : 0u;             // MARK:84
                }
                if (taps.flags) {
                    taps.flags[i + j] = (int)frag;
                    taps.flags[n + i + j] = (int)span;
                    uint32_t pp;
                    taps.skey32[i + j] = unpack_key32(L, k[j + 1], pp);
                }
            }
            x[j] = (unsigned long long)frag | ((unsigned long long)span << 31);
            a.xy[j] = ((uint32_t)y1 << 16) | ((uint32_t)x1 & 0xFFFFu);
            a.xprev[j] = (uint32_t)max(0, x0 + FRAG_SIZE);  // GEN:88-92
            a.path[j] = p1;
            p0 = p1; x0 = x1; y0 = y1;
        }
    }

    __device__ void store(long long i, long long n, const unsigned long long e[4], const unsigned long long *x,
                          const Aux &a) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j >= n) break;
            const int frag_before = (int)(e[j] & 0x7FFFFFFFu), span_before = (int)((e[j] >> 31) & 0x7FFFFFFFu);
            const int frag = (int)(x[j] & 1u), span = (int)((x[j] >> 31) & 1u);
            if (taps.scan3) {
                taps.scan3[i + j] = frag_before;
                taps.scan3[n + i + j] = span_before;  // + n_out_frag, added by k_scan3_fixup
            }
            if (frag | span) {
                const int fill = (int)fill_info[a.path[j]];
                const int oi = frag_before + span_before;  // GEN:64-66
                if (frag) records[oi] = make_int4((int)a.xy[j], 2, fill, frag_before + 1);  // GEN:77 (frag_index is inclusive)
                if (span) {
                    const int xs = (int)a.xprev[j];
                    const int xe = (int)(a.xy[j] & 0xFFFFu);  // span flag implies a valid key, so x1 >= 0
                    records[oi + frag] = make_int4((int)((a.xy[j] & 0xFFFF0000u) | (uint32_t)xs), xe - xs, fill, 0);  // GEN:102
                }
            }
        }
    }

    __device__ void finish(long long n, unsigned long long total) const {
        const int nfrag = (int)(total & 0x7FFFFFFFu), nspan = (int)((total >> 31) & 0x7FFFFFFFu);
        ctr->n_out_frag = nfrag;
        ctr->n_span = nspan;
        ctr->n_records = nfrag + nspan;
        if (taps.scan3) { taps.scan3[n] = 0; taps.scan3[2 * n] = nspan; }  // fixed up with + n_out_frag
    }
};

// scan3[nf + i] += n_out_frag for i in [0, nf] (tap only): turns the two separate counts into the
// reference's single scan over the concatenated [frag | span] flag array (SR.cpp:545-573).
__global__ void k_scan3_fixup(const FrameCounters *__restrict__ ctr, int capacity, int *__restrict__ scan3) {
    const int nf = ctr->n_fragments;
    if (nf > capacity) return;
    const int add = ctr->n_out_frag;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nf; i += gridDim.x * blockDim.x) scan3[nf + i] += add;
}

}  // namespace slpr
