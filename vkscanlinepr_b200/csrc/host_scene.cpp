// host_scene.cpp — host scene front end behind the C ABI (include/slpr.h, slpr_vg_*):
//   * RVG text -> container, with the behaviour of the reference parser
//     (VkScanlinePR/src/core/vg/rvg.cpp:9-255), quirks included: only M, L, C, Z
//     and fL do anything; lower-case commands, A and anything else end the command loop without
//     consuming operands; unclosed contours get a closing LINE to the first point of the LAST
//     M-contour; a paint that is not `solid` leaves the default colour (0,0,0,1); when a token
//     cannot be read as the opacity the stream fails and parsing stops for the rest of the file.
//   * container -> the 7 flat arrays + colour quantisation of ScanlineVGRasterizer::loadVG
//     (VkScanlinePR/src/core/scanline/scanline_rasterizer.cpp:67-118), without mutating the input.
// No reference source is copied: this is an independent restatement checked against the
// reference's own parser compiled in place (oracle/_ref, tests/test_scene_frontend.py).
#include <stdint.h>
#include <stdio.h>

#include <math.h>
#include <stdlib.h>

#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/slpr.h"

namespace {

thread_local std::string g_scene_err;

struct Container {  // Galaxysailing::VGContainer as flat vectors (vg_container.h:21-87)
    float vp[4] = {0, 0, 0, 0}, win[4] = {0, 0, 0, 0};
    std::vector<float> pos;            // x,y interleaved
    std::vector<uint32_t> curve_pos;   // CurveData::posIndices
    std::vector<uint32_t> curve_type;  // CurveData::curveType
    std::vector<uint32_t> path_curve;  // PathData::curveIndices
    std::vector<uint32_t> fill_rule;
    std::vector<float> fill_color;  // rgba per path
    std::vector<float> fill_opacity;
    std::vector<float> curve_weight;  // full loader only: middle weight of ARC curves (1 for the others); empty otherwise

    size_t n_points() const { return pos.size() / 2; }
    void new_path() {  // vg_container.h:49-56
        path_curve.push_back((uint32_t)curve_pos.size());
        fill_rule.push_back(0);
        const float dflt[4] = {0.f, 0.f, 0.f, 1.f};
        fill_color.insert(fill_color.end(), dflt, dflt + 4);
        fill_opacity.push_back(0.0f);
    }
    void add_curve(uint32_t type, const float (*p)[2], int npts) {  // newCurve + addCurve, vg_container.h:58-85
        curve_pos.push_back((uint32_t)n_points());
        curve_type.push_back(type);
        for (int i = 0; i < npts; ++i) { pos.push_back(p[i][0]); pos.push_back(p[i][1]); }
    }
};

struct Flat {
    std::vector<uint32_t> pos_path, curve_path, fill_info;
};

}  // namespace

struct slpr_vg {
    Container c;
    Flat f;
    bool flattened = false;
};

namespace {

// A point token is "x,y" possibly with ':' separators (rvg.cpp:38-57): both become blanks, then "%f %f".
void read_point(std::istream &in, float out[2]) {
    std::string tok;
    in >> tok;
    for (char &ch : tok)
        if (ch == ',' || ch == ':') ch = ' ';
    float x = 0.f, y = 0.f;
    sscanf(tok.c_str(), "%f %f", &x, &y);
    out[0] = x; out[1] = y;
}

void parse_rvg(std::istream &in, Container &vg) {
    std::string tok;
    float a[2], b[2];
    // header (rvg.cpp:33-83): "viewport p p", "window p p", then two tokens ("scene", "dyn_identity")
    in >> tok; read_point(in, a); read_point(in, b);
    vg.vp[0] = a[0]; vg.vp[1] = a[1]; vg.vp[2] = b[0]; vg.vp[3] = b[1];
    in >> tok; read_point(in, a); read_point(in, b);
    vg.win[0] = a[0]; vg.win[1] = a[1]; vg.win[2] = b[0]; vg.win[3] = b[1];
    in >> tok; in >> tok;

    long long last_path_points = -1;
    while (in >> tok) {  // one `element` per iteration (rvg.cpp:113-254)
        if (tok.size() >= 2 && tok[0] == '/' && tok[1] == '/') { std::getline(in, tok); continue; }
        if (last_path_points != (long long)vg.n_points()) {  // a new path only if the last one got points
            last_path_points = (long long)vg.n_points();
            vg.new_path();
        }
        const size_t path = vg.fill_rule.size() - 1;
        in >> tok;  // "element"
        in >> tok;  // fill rule
        if (tok == "ofill") vg.fill_rule[path] = 1;
        else if (tok == "nzfill") vg.fill_rule[path] = 0;
        else { in >> tok; continue; }
        in >> tok;  // "dyn_concrete"
        float skip[2];
        read_point(in, skip); read_point(in, skip); read_point(in, skip);

        float p[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        size_t contour_begin = vg.n_points();
        bool closed = false;
        while (in >> tok) {  // command loop (rvg.cpp:159-203)
            const bool is_fl = (tok == "fL");
            if (!is_fl && tok.size() != 1) break;
            const char cmd = tok[0];
            if (!is_fl && std::string("MmZzLlCcAa").find(cmd) == std::string::npos) break;
            if (is_fl) { read_point(in, p[0]); continue; }
            if (cmd == 'M') { read_point(in, p[0]); contour_begin = vg.n_points(); closed = false; }
            else if (cmd == 'L') {
                read_point(in, p[1]);
                vg.add_curve(0x02, p, 2);
                p[0][0] = p[1][0]; p[0][1] = p[1][1];
            } else if (cmd == 'C') {
                read_point(in, p[1]); read_point(in, p[2]); read_point(in, p[3]);
                vg.add_curve(0x04, p, 4);
                p[0][0] = p[3][0]; p[0][1] = p[3][1];
            } else if (cmd == 'Z') closed = true;
            // m z l c A a: accepted, nothing consumed (rvg.cpp:196-201)
        }
        if (!closed && vg.n_points() > contour_begin) {  // implicit close (rvg.cpp:210-223)
            const float fx = vg.pos[2 * contour_begin], fy = vg.pos[2 * contour_begin + 1];
            const float lx = vg.pos[vg.pos.size() - 2], ly = vg.pos[vg.pos.size() - 1];
            if (fx != lx || fy != ly) {
                const float q[2][2] = {{lx, ly}, {fx, fy}};
                vg.add_curve(0x02, q, 2);
            }
        }
        in >> tok;  // "dyn_paint" (the loop left "dyn_identity" in tok)
        float opacity = 0.f;
        in >> opacity;
        vg.fill_opacity[path] = opacity;
        in >> tok;
        if (tok == "solid") {  // rvg.cpp:233-251
            in >> tok;
            float *col = &vg.fill_color[4 * path];
            if (tok.compare(0, 4, "rgba") == 0 && tok.size() >= 6) {
                std::string body = tok.substr(5, tok.size() - 6);
                for (char &ch : body) if (ch == ',') ch = ' ';
                std::istringstream iss(body);
                iss >> col[0] >> col[1] >> col[2] >> col[3];
            } else if (tok.compare(0, 3, "rgb") == 0 && tok.size() >= 6) {
                std::string body = tok.substr(4, tok.size() - 6);  // the reference's off-by-one length; ')' stops the parse
                for (char &ch : body) if (ch == ',') ch = ' ';
                std::istringstream iss(body);
                iss >> col[0] >> col[1] >> col[2];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// a COMPLETE reader of the RVG files the reference ships (and of the format's other path commands),
// next to the quirk-faithful one above. What the reference's parser drops is kept:
//   * `A x1,y1,w x2,y2` — rational quadratic segments (control point in homogeneous coordinates) become ARC
//     curves (0x13, three points, Euclidean control point) with a per-curve weight; an arc whose control point is at
//     infinity (w = 0: the half ellipses of car.rvg) or has a negative weight is split at t = 1/2 into two arcs of
//     positive weight sqrt((1 + w) / 2);
//   * `Q` (QUADRIC, 0x03, three points), `H` / `V`, the lower-case relative forms of every command, `Z` per contour;
//   * the element's own transform (dyn_identity | dyn_affine([a,b,c],[d,e,f]) | dyn_translation | dyn_scaling |
//     dyn_rotation) is applied to its control points (an affine map keeps an arc's weights);
//   * every contour is closed (a LINE back to its first point), not only the last one;
//   * gradient paints (the scanline pipeline fills with one colour per path, as the reference does) are replaced by
//     the average colour of their ramp.
// Each `element` is one line of the file. Curves of type QUADRIC / ARC need SLPR_FLAG_FULL_RVG to be rendered.
struct Affine {
    float a = 1, b = 0, c = 0, d = 0, e = 1, f = 0;
    void apply(float &x, float &y) const {
        const float nx = a * x + b * y + c, ny = d * x + e * y + f;
        x = nx; y = ny;
    }
};

// numbers of a token like "12.5,-3e-2,0.99" or "(1,2)" or "rgba(1,0,0,1)": every maximal run that strtof accepts
std::vector<float> numbers_in(const std::string &tok, size_t from = 0) {
    std::vector<float> v;
    const char *s = tok.c_str() + from;
    while (*s) {
        if ((*s >= '0' && *s <= '9') || *s == '-' || *s == '+' || *s == '.') {
            char *end = nullptr;
            const float x = strtof(s, &end);
            if (end != s) { v.push_back(x); s = end; continue; }
        }
        ++s;
    }
    return v;
}

void parse_rvg_full(std::istream &in, Container &vg) {
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::vector<std::string> tok;
        for (std::string t; ls >> t;) tok.push_back(t);
        if (tok.empty() || (tok[0].size() >= 2 && tok[0][0] == '/' && tok[0][1] == '/')) continue;
        if (tok[0] == "viewport" || tok[0] == "window") {
            float *dst = tok[0] == "viewport" ? vg.vp : vg.win;
            std::vector<float> n;
            for (size_t i = 1; i < tok.size(); ++i) { const std::vector<float> m = numbers_in(tok[i]); n.insert(n.end(), m.begin(), m.end()); }
            for (size_t i = 0; i < 4 && i < n.size(); ++i) dst[i] = n[i];
            continue;
        }
        size_t i = 0;
        while (i < tok.size() && tok[i] != "element") ++i;
        if (i + 1 >= tok.size()) continue;  // "scene dyn_identity" and anything else that is not an element
        const std::string rule = tok[i + 1];
        if (rule != "ofill" && rule != "nzfill") continue;
        // path data starts after the token that ends with ':' (the dyn_concrete header)
        size_t k = i + 2;
        while (k < tok.size() && tok[k].back() != ':') ++k;
        ++k;
        struct Seg { uint32_t type; float p[4][2]; float w; };
        std::vector<Seg> segs;
        float cur[2] = {0, 0}, start[2] = {0, 0};
        bool open = false;  // the current contour has segments and was not closed
        auto close_contour = [&]() {
            if (open && (cur[0] != start[0] || cur[1] != start[1])) {
                Seg s{0x02, {{cur[0], cur[1]}, {start[0], start[1]}, {0, 0}, {0, 0}}, 1.f};
                segs.push_back(s);
            }
            open = false;
        };
        auto point = [&](const std::string &t, bool rel, float out[2], float *w = nullptr) {
            const std::vector<float> n = numbers_in(t);
            out[0] = (n.size() > 0 ? n[0] : 0.f) + (rel ? cur[0] : 0.f);
            out[1] = (n.size() > 1 ? n[1] : 0.f) + (rel ? cur[1] : 0.f);
            if (w) *w = n.size() > 2 ? n[2] : 1.f;
        };
        for (; k < tok.size(); ++k) {
            const std::string &t = tok[k];
            if (t.compare(0, 4, "dyn_") == 0) break;
            if (t.size() != 1) continue;  // stray operand
            const char cmd = t[0];
            const bool rel = cmd >= 'a' && cmd <= 'z';
            const char C = rel ? (char)(cmd - 32) : cmd;
            if (C == 'M' && k + 1 < tok.size()) {
                close_contour();
                point(tok[++k], rel, cur);
                start[0] = cur[0]; start[1] = cur[1];
            } else if (C == 'Z') {
                close_contour();
                cur[0] = start[0]; cur[1] = start[1];
            } else if ((C == 'L' || C == 'H' || C == 'V') && k + 1 < tok.size()) {
                float q[2] = {cur[0], cur[1]};
                if (C == 'L') point(tok[++k], rel, q);
                else {
                    const std::vector<float> n = numbers_in(tok[++k]);
                    const float v = n.empty() ? 0.f : n[0];
                    if (C == 'H') q[0] = v + (rel ? cur[0] : 0.f); else q[1] = v + (rel ? cur[1] : 0.f);
                }
                Seg s{0x02, {{cur[0], cur[1]}, {q[0], q[1]}, {0, 0}, {0, 0}}, 1.f};
                segs.push_back(s); open = true;
                cur[0] = q[0]; cur[1] = q[1];
            } else if (C == 'Q' && k + 2 < tok.size()) {
                float c1[2], e[2];
                point(tok[++k], rel, c1); point(tok[++k], rel, e);
                Seg s{0x03, {{cur[0], cur[1]}, {c1[0], c1[1]}, {e[0], e[1]}, {0, 0}}, 1.f};
                segs.push_back(s); open = true;
                cur[0] = e[0]; cur[1] = e[1];
            } else if (C == 'C' && k + 3 < tok.size()) {
                float c1[2], c2[2], e[2];
                point(tok[++k], rel, c1); point(tok[++k], rel, c2); point(tok[++k], rel, e);
                Seg s{0x04, {{cur[0], cur[1]}, {c1[0], c1[1]}, {c2[0], c2[1]}, {e[0], e[1]}}, 1.f};
                segs.push_back(s); open = true;
                cur[0] = e[0]; cur[1] = e[1];
            } else if ((C == 'A' || C == 'R') && k + 2 < tok.size()) {  // rational quadratic: homogeneous control point (X, Y, W)
                float h[2], W = 1.f, e[2];
                const std::vector<float> n = numbers_in(tok[++k]);
                h[0] = n.size() > 0 ? n[0] : 0.f; h[1] = n.size() > 1 ? n[1] : 0.f; W = n.size() > 2 ? n[2] : 1.f;
                if (rel) { h[0] += cur[0] * W; h[1] += cur[1] * W; }
                point(tok[++k], rel, e);
                if (W > 1e-6f) {
                    Seg s{0x13, {{cur[0], cur[1]}, {h[0] / W, h[1] / W}, {e[0], e[1]}, {0, 0}}, W};
                    segs.push_back(s);
                } else if (1.f + W > 1e-6f) {  // control point at infinity or beyond: two arcs of weight sqrt((1 + W) / 2)
                    const float ws = sqrtf((1.f + W) * 0.5f), inv = 1.f / (1.f + W);
                    const float a0[2] = {(cur[0] + h[0]) * inv, (cur[1] + h[1]) * inv};  // Euclidean control of the first half
                    const float a1[2] = {(e[0] + h[0]) * inv, (e[1] + h[1]) * inv};      // ... of the second half
                    const float m[2] = {(cur[0] + 2.f * h[0] + e[0]) * 0.5f * inv, (cur[1] + 2.f * h[1] + e[1]) * 0.5f * inv};
                    Seg s0{0x13, {{cur[0], cur[1]}, {a0[0], a0[1]}, {m[0], m[1]}, {0, 0}}, ws};
                    Seg s1{0x13, {{m[0], m[1]}, {a1[0], a1[1]}, {e[0], e[1]}, {0, 0}}, ws};
                    segs.push_back(s0); segs.push_back(s1);
                } else {  // degenerate weights: a straight segment
                    Seg s{0x02, {{cur[0], cur[1]}, {e[0], e[1]}, {0, 0}, {0, 0}}, 1.f};
                    segs.push_back(s);
                }
                open = true;
                cur[0] = e[0]; cur[1] = e[1];
            }
        }
        close_contour();
        // the element's transform
        Affine xf;
        if (k < tok.size()) {
            const std::string &t = tok[k];
            const std::vector<float> n = numbers_in(t, 4);
            if (t.compare(0, 10, "dyn_affine") == 0 && n.size() >= 6) { xf.a = n[0]; xf.b = n[1]; xf.c = n[2]; xf.d = n[3]; xf.e = n[4]; xf.f = n[5]; }
            else if (t.compare(0, 15, "dyn_translation") == 0 && n.size() >= 2) { xf.c = n[0]; xf.f = n[1]; }
            else if (t.compare(0, 11, "dyn_scaling") == 0 && n.size() >= 1) { xf.a = n[0]; xf.e = n.size() > 1 ? n[1] : n[0]; }
            else if (t.compare(0, 12, "dyn_rotation") == 0 && n.size() >= 1) {
                const float r = n[0] * 3.14159265358979f / 180.f;
                xf.a = cosf(r); xf.b = -sinf(r); xf.d = sinf(r); xf.e = cosf(r);
            }
            ++k;
        }
        // paint: dyn_paint <opacity> solid rgba(..) | ... gradient ... dyn_ramp <t:rgba(..)>* ...
        float opacity = 1.f, col[4] = {0, 0, 0, 1};
        while (k < tok.size() && tok[k] != "dyn_paint") ++k;
        if (k + 2 < tok.size()) {
            opacity = strtof(tok[k + 1].c_str(), nullptr);
            if (tok[k + 2] == "solid" && k + 3 < tok.size()) {
                const std::vector<float> n = numbers_in(tok[k + 3], tok[k + 3].find('('));
                for (size_t q = 0; q < 4 && q < n.size(); ++q) col[q] = n[q];
                if (n.size() == 3) col[3] = 1.f;
            } else {  // gradient: average of the ramp's stops ("t:rgba(r,g,b,a)")
                float sum[4] = {0, 0, 0, 0};
                int stops = 0;
                for (size_t q = k + 3; q < tok.size(); ++q) {
                    const size_t at = tok[q].find(":rgb");
                    if (at == std::string::npos) continue;
                    const std::vector<float> n = numbers_in(tok[q], tok[q].find('(', at));
                    if (n.size() >= 3) {
                        for (int c = 0; c < 3; ++c) sum[c] += n[(size_t)c];
                        sum[3] += n.size() > 3 ? n[3] : 1.f;
                        ++stops;
                    }
                }
                if (stops) for (int c = 0; c < 4; ++c) col[c] = sum[c] / (float)stops;
            }
        }
        if (segs.empty()) continue;
        vg.new_path();
        const size_t path = vg.fill_rule.size() - 1;
        vg.fill_rule[path] = rule == "ofill" ? 1u : 0u;
        vg.fill_opacity[path] = opacity;
        for (int c = 0; c < 4; ++c) vg.fill_color[4 * path + c] = col[c];
        for (Seg &s : segs) {
            const int npts = s.type == 0x02 ? 2 : s.type == 0x04 ? 4 : 3;
            for (int q = 0; q < npts; ++q) xf.apply(s.p[q][0], s.p[q][1]);
            vg.add_curve(s.type, s.p, npts);
            vg.curve_weight.push_back(s.w);
        }
    }
}

}  // namespace

extern "C" {

// slpr_last_error() lives in slpr.cu; scene errors are reported through the same accessor by
// storing them with this hook.
void slpr_internal_set_error(const char *msg);

slpr_vg *slpr_vg_load_rvg(const char *path) {
    if (!path) { slpr_internal_set_error("slpr_vg_load_rvg: null path"); return nullptr; }
    std::ifstream in(path);
    if (!in.is_open()) {  // rvg.cpp:13-15 throws std::runtime_error here
        slpr_internal_set_error((std::string("RVG::load can't open file \"") + path + "\"").c_str());
        return nullptr;
    }
    slpr_vg *vg = new slpr_vg();
    parse_rvg(in, vg->c);
    return vg;
}

slpr_vg *slpr_vg_load_rvg_full(const char *path) {
    if (!path) { slpr_internal_set_error("slpr_vg_load_rvg_full: null path"); return nullptr; }
    std::ifstream in(path);
    if (!in.is_open()) {
        slpr_internal_set_error((std::string("RVG::load can't open file \"") + path + "\"").c_str());
        return nullptr;
    }
    slpr_vg *vg = new slpr_vg();
    parse_rvg_full(in, vg->c);
    return vg;
}

slpr_vg *slpr_vg_from_arrays(const float *pos_xy, uint32_t n_points, const uint32_t *curve_pos, const uint32_t *curve_type,
                             uint32_t n_curves, const uint32_t *path_curve, const uint32_t *fill_rule,
                             const float *fill_color_rgba, const float *fill_opacity, uint32_t n_paths) {
    if ((n_points && !pos_xy) || (n_curves && (!curve_pos || !curve_type)) ||
        (n_paths && (!path_curve || !fill_rule || !fill_color_rgba || !fill_opacity))) {
        slpr_internal_set_error("slpr_vg_from_arrays: null array with non-zero count");
        return nullptr;
    }
    slpr_vg *vg = new slpr_vg();
    Container &c = vg->c;
    c.pos.assign(pos_xy, pos_xy + 2 * (size_t)n_points);
    c.curve_pos.assign(curve_pos, curve_pos + n_curves);
    c.curve_type.assign(curve_type, curve_type + n_curves);
    c.path_curve.assign(path_curve, path_curve + n_paths);
    c.fill_rule.assign(fill_rule, fill_rule + n_paths);
    c.fill_color.assign(fill_color_rgba, fill_color_rgba + 4 * (size_t)n_paths);
    c.fill_opacity.assign(fill_opacity, fill_opacity + n_paths);
    return vg;
}

// loadVG's flattening (scanline_rasterizer.cpp:67-118): walk paths -> curves -> points.
int slpr_vg_flatten(slpr_vg *vg, slpr_scene_view *out) {
    if (!vg || !out) { slpr_internal_set_error("slpr_vg_flatten: null argument"); return SLPR_ERR_INVALID; }
    Container &c = vg->c;
    const uint32_t n_paths = (uint32_t)c.path_curve.size(), n_curves = (uint32_t)c.curve_pos.size();
    const uint32_t n_points = (uint32_t)c.n_points();
    if (!vg->flattened) {
        Flat &f = vg->f;
        f.pos_path.assign(n_points, 0);
        f.curve_path.assign(n_curves, 0);
        f.fill_info.assign(n_paths, 0);
        uint32_t expect_curve = 0, expect_point = 0;
        for (uint32_t pi = 0; pi < n_paths; ++pi) {
            const uint32_t cb = c.path_curve[pi], ce = (pi + 1 != n_paths) ? c.path_curve[pi + 1] : n_curves;
            // colour: a *= opacity; rgba *= 255; truncate to u8; the word is 0 when alpha is 0 (SR.cpp:107-118)
            float col[4] = {c.fill_color[4 * pi], c.fill_color[4 * pi + 1], c.fill_color[4 * pi + 2], c.fill_color[4 * pi + 3]};
            col[3] = col[3] * c.fill_opacity[pi];
            uint32_t word = 0;
            for (int k = 0; k < 4; ++k) word |= (uint32_t)(uint8_t)(col[k] * 255.0f) << (8 * k);
            f.fill_info[pi] = (word & 0xFF000000u) ? word : 0u;
            if (cb > ce || ce > n_curves || cb != expect_curve) {
                slpr_internal_set_error("slpr_vg_flatten: path curve ranges must tile [0, n_curves) in order");
                return SLPR_ERR_INVALID;
            }
            for (uint32_t ci = cb; ci < ce; ++ci) {
                const uint32_t pb = c.curve_pos[ci], pe = (ci + 1 != n_curves) ? c.curve_pos[ci + 1] : n_points;
                if (pb > pe || pe > n_points || pb != expect_point) {
                    slpr_internal_set_error("slpr_vg_flatten: curve point ranges must tile [0, n_points) in order");
                    return SLPR_ERR_INVALID;
                }
                f.curve_path[ci] = pi;
                for (uint32_t k = pb; k < pe; ++k) f.pos_path[k] = pi;
                expect_point = pe;
            }
            expect_curve = ce;
        }
        vg->flattened = true;
    }
    out->pos_xy = c.pos.data(); out->pos_path = vg->f.pos_path.data(); out->n_points = n_points;
    out->curve_pos_map = c.curve_pos.data(); out->curve_type = c.curve_type.data();
    out->curve_path = vg->f.curve_path.data(); out->n_curves = n_curves;
    out->fill_rule = c.fill_rule.data(); out->fill_rgba8 = vg->f.fill_info.data(); out->n_paths = n_paths;
    for (int k = 0; k < 4; ++k) { out->viewport[k] = c.vp[k]; out->window[k] = c.win[k]; }
    out->curve_weight = c.curve_weight.size() == n_curves && n_curves ? c.curve_weight.data() : nullptr;
    return SLPR_OK;
}

int slpr_vg_container(slpr_vg *vg, const float **pos_xy, uint32_t *n_points, const uint32_t **curve_pos,
                      const uint32_t **curve_type, uint32_t *n_curves, const uint32_t **path_curve,
                      const uint32_t **fill_rule, const float **fill_color, const float **fill_opacity, uint32_t *n_paths) {
    if (!vg) { slpr_internal_set_error("slpr_vg_container: null argument"); return SLPR_ERR_INVALID; }
    Container &c = vg->c;
    if (pos_xy) *pos_xy = c.pos.data();
    if (n_points) *n_points = (uint32_t)c.n_points();
    if (curve_pos) *curve_pos = c.curve_pos.data();
    if (curve_type) *curve_type = c.curve_type.data();
    if (n_curves) *n_curves = (uint32_t)c.curve_pos.size();
    if (path_curve) *path_curve = c.path_curve.data();
    if (fill_rule) *fill_rule = c.fill_rule.data();
    if (fill_color) *fill_color = c.fill_color.data();
    if (fill_opacity) *fill_opacity = c.fill_opacity.data();
    if (n_paths) *n_paths = (uint32_t)c.path_curve.size();
    return SLPR_OK;
}

void slpr_vg_free(slpr_vg *vg) { delete vg; }

}  // extern "C"
