// slpr_rasterizer.hpp — header-only C++ mirror of the reference's renderer interface over the C ABI
// (include/slpr.h). A user of galaxysailing/VkScanlinePR switches by replacing
//     ScanlineVGRasterizer            (VkScanlinePR/src/core/scanline/scanline_rasterizer.h)
// with
//     Galaxysailing::CudaVGRasterizer (this file)
// and keeping the call sequence of ScanlineVGApplication::run() (app/vg_app.cpp:146-165):
//     initialize(nullptr, w, h); loadVG(container); loop { setMVP(transpose(camera.mv())); render(); }
// plus the headless readback() the north star adds in place of acquire/present.
//
// Inside the reference tree, define SLPR_WITH_REFERENCE_HEADERS before including this file: the
// class then derives from the reference's own Galaxysailing::VGRasterizer (core/rasterizer.h:9-20)
// and takes its Galaxysailing::VGContainer (core/vg/vg_container.h:21-87). Stand-alone, the same
// names are declared here with the same members, and a minimal glm-compatible vec/mat is used when
// <glm/glm.hpp> is not on the include path. Errors are std::runtime_error, as in the reference
// (scanline_rasterizer.cpp:219,224; vg/rvg.cpp:14).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "slpr.h"

#if defined(SLPR_WITH_REFERENCE_HEADERS)
#include "core/rasterizer.h"  // brings vg/vg_container.h and glm
#include "core/vg/rvg.h"
#else
#if __has_include(<glm/glm.hpp>)
#include <glm/glm.hpp>
#else
namespace glm {  // the three types the interface needs, layout-compatible with glm's
struct vec2 { float x = 0, y = 0; vec2() = default; vec2(float a, float b) : x(a), y(b) {}
              bool operator!=(const vec2 &o) const { return x != o.x || y != o.y; } };
struct vec4 { union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; float v[4]; };
              vec4() : x(0), y(0), z(0), w(0) {} vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
              float &operator[](int i) { return v[i]; } const float &operator[](int i) const { return v[i]; } };
struct mat4 { vec4 col[4];  // column-major like glm: m[i] is column i
              mat4() : mat4(1.f) {}
              explicit mat4(float d) { for (int i = 0; i < 4; ++i) col[i][i] = d; }  // glm::mat4(1.0f) = identity
              vec4 &operator[](int i) { return col[i]; } const vec4 &operator[](int i) const { return col[i]; } };
inline mat4 transpose(const mat4 &m) { mat4 t; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) t[i][j] = m[j][i]; return t; }
}  // namespace glm
#endif

namespace Galaxysailing {

// core/vg/vg_container.h:9-19
enum class CurveType { NONE = 0x00, LINE = 0x02, QUADRIC = 0x03, CUBIC = 0x04, ARC = 0x13 };
enum class FillRule { NON_ZERO = 0, EVEN_ODD = 1 };

// core/vg/vg_container.h:21-87 (same members, same helper methods)
struct VGContainer {
    glm::vec4 vp, win;
    struct PointData { std::vector<glm::vec2> pos; };
    struct CurveData { std::vector<uint32_t> posIndices; std::vector<CurveType> curveType; int curveIndex = -1; };
    struct PathData {
        std::vector<uint32_t> curveIndices;
        std::vector<FillRule> fillRule;
        std::vector<glm::vec4> fillColor;
        std::vector<float> fillOpacity;
        int pathIndex = -1;
    };
    PathData pathData;
    CurveData curveData;
    PointData pointData;

    void newPath() {
        ++pathData.pathIndex;
        pathData.curveIndices.push_back((uint32_t)(curveData.curveIndex + 1));
        pathData.fillRule.push_back(FillRule::NON_ZERO);
        pathData.fillColor.push_back(glm::vec4(0, 0, 0, 1));
        pathData.fillOpacity.push_back(0.0f);
    }
    void newCurve() {
        ++curveData.curveIndex;
        curveData.posIndices.push_back((uint32_t)pointData.pos.size());
        curveData.curveType.push_back(CurveType::NONE);
    }
    void addCurve(CurveType ct, glm::vec2 *p) {
        curveData.curveType[(size_t)curveData.curveIndex] = ct;
        const int n = ct == CurveType::LINE ? 2 : ct == CurveType::CUBIC ? 4 : 0;  // QUADRIC / ARC store no points
        for (int i = 0; i < n; ++i) pointData.pos.push_back(p[i]);
    }
};

// core/rasterizer.h:9-20
class VGRasterizer {
public:
    virtual ~VGRasterizer() = default;
    virtual void initialize(void *window, uint32_t w, uint32_t h) = 0;
    virtual void render() = 0;
    virtual void loadVG(std::shared_ptr<VGContainer> vg) = 0;
    virtual void setMVP(const glm::mat4 &m) = 0;
};

// core/vg/rvg.h:11-22 — the parser lives in libslpr.so (csrc/host_scene.cpp, behaviour of vg/rvg.cpp:9-255)
class RVG {
public:
    void load(const std::string &filename) {
        slpr_vg *h = slpr_vg_load_rvg(filename.c_str());
        if (!h) throw std::runtime_error(slpr_last_error());
        try { adopt(h); } catch (...) { slpr_vg_free(h); throw; }
        slpr_vg_free(h);
    }
    std::shared_ptr<VGContainer> getVGContainer() { return _vgContainer; }
    // SURVEY section 8 f-1, beyond the reference: the complete reader (rational arcs, quadratics, relative commands, per-element
    // transforms, gradients as their average colour; slpr_vg_load_rvg_full). QUADRIC / ARC curves carry their three
    // control points here (the reference's addCurve stores none for them) and getCurveWeights() has the arcs' weights:
    // render with CudaVGRasterizer(device, SLPR_FLAG_FULL_RVG) and setCurveWeights().
    void loadFull(const std::string &filename) {
        slpr_vg *h = slpr_vg_load_rvg_full(filename.c_str());
        if (!h) throw std::runtime_error(slpr_last_error());
        try { adopt(h); } catch (...) { slpr_vg_free(h); throw; }
        slpr_vg_free(h);
    }
    const std::vector<float> &getCurveWeights() const { return _curveWeights; }
private:
    void adopt(slpr_vg *h) {
        const float *pos, *col, *op; const uint32_t *cpos, *ctype, *pcur, *frule; uint32_t np, nc, npath;
        slpr_scene_view v;
        if (slpr_vg_container(h, &pos, &np, &cpos, &ctype, &nc, &pcur, &frule, &col, &op, &npath) || slpr_vg_flatten(h, &v))
            throw std::runtime_error(slpr_last_error());
        auto vg = std::make_shared<VGContainer>();
        vg->vp = glm::vec4(v.viewport[0], v.viewport[1], v.viewport[2], v.viewport[3]);
        vg->win = glm::vec4(v.window[0], v.window[1], v.window[2], v.window[3]);
        for (uint32_t i = 0; i < np; ++i) vg->pointData.pos.push_back(glm::vec2(pos[2 * i], pos[2 * i + 1]));
        vg->curveData.posIndices.assign(cpos, cpos + nc);
        for (uint32_t i = 0; i < nc; ++i) vg->curveData.curveType.push_back((CurveType)ctype[i]);
        vg->curveData.curveIndex = (int)nc - 1;
        vg->pathData.curveIndices.assign(pcur, pcur + npath);
        for (uint32_t i = 0; i < npath; ++i) {
            vg->pathData.fillRule.push_back((FillRule)frule[i]);
            vg->pathData.fillColor.push_back(glm::vec4(col[4 * i], col[4 * i + 1], col[4 * i + 2], col[4 * i + 3]));
            vg->pathData.fillOpacity.push_back(op[i]);
        }
        vg->pathData.pathIndex = (int)npath - 1;
        _curveWeights.assign(v.curve_weight ? v.curve_weight : nullptr, v.curve_weight ? v.curve_weight + nc : nullptr);
        _vgContainer = vg;
    }
    std::shared_ptr<VGContainer> _vgContainer;
    std::vector<float> _curveWeights;
};

}  // namespace Galaxysailing
#endif  // !SLPR_WITH_REFERENCE_HEADERS

namespace Galaxysailing {

// Drop-in for ScanlineVGRasterizer (scanline/scanline_rasterizer.h): same four calls, CUDA underneath.
class CudaVGRasterizer : public VGRasterizer {
public:
    explicit CudaVGRasterizer(int device = 0, uint32_t flags = 0) : _device(device), _flags(flags) {}
    // (no `override`: the reference's VGRasterizer declares no virtual destructor, core/rasterizer.h:9-20)
    ~CudaVGRasterizer() { slpr_destroy(_ctx); }
    CudaVGRasterizer(const CudaVGRasterizer &) = delete;
    CudaVGRasterizer &operator=(const CudaVGRasterizer &) = delete;

    // scanline_rasterizer.cpp:42-55. `window` is the reference's GLFWwindow*; headless: must be null.
    void initialize(void *window, uint32_t w, uint32_t h) override {
        if (window) throw std::runtime_error("CudaVGRasterizer is headless: pass window = nullptr and use readback()");
        slpr_destroy(_ctx);
        _ctx = slpr_create(_device, w, h, _flags);
        if (!_ctx) throw std::runtime_error(slpr_last_error());
        _width = w; _height = h;
    }

    // scanline_rasterizer.cpp:67-171. The container is only read (the reference pre-multiplies its
    // colours in place, SR.cpp:108-110; that side effect is not replicated).
    void loadVG(std::shared_ptr<VGContainer> vg) override {
        need_ctx();
        const auto &pt = vg->pointData; const auto &cv = vg->curveData; const auto &pa = vg->pathData;
        const uint32_t np = (uint32_t)pt.pos.size(), nc = (uint32_t)(cv.curveIndex + 1), npath = (uint32_t)(pa.pathIndex + 1);
        std::vector<float> pos(2 * (size_t)np), col(4 * (size_t)npath);
        std::vector<uint32_t> ctype(nc), frule(npath);
        for (uint32_t i = 0; i < np; ++i) { pos[2 * i] = pt.pos[i].x; pos[2 * i + 1] = pt.pos[i].y; }
        for (uint32_t i = 0; i < nc; ++i) ctype[i] = (uint32_t)cv.curveType[i];
        for (uint32_t i = 0; i < npath; ++i) {
            frule[i] = (uint32_t)pa.fillRule[i];
            for (int k = 0; k < 4; ++k) col[4 * i + k] = pa.fillColor[i][k];
        }
        slpr_vg *h = slpr_vg_from_arrays(pos.data(), np, cv.posIndices.data(), ctype.data(), nc, pa.curveIndices.data(),
                                         frule.data(), col.data(), pa.fillOpacity.data(), npath);
        if (!h) throw std::runtime_error(slpr_last_error());
        slpr_scene_view v;
        int rc = slpr_vg_flatten(h, &v);
        if (!rc) rc = slpr_load_scene(_ctx, v.pos_xy, v.pos_path, v.n_points, v.curve_pos_map, v.curve_type, v.curve_path,
                                      v.n_curves, v.fill_rule, v.fill_rgba8, v.n_paths);
        slpr_vg_free(h);
        check(rc);
    }

    // scanline_rasterizer.cpp:191-197: the caller passes transpose(camera.mv()) so that m[i] is ROW i
    // of the transform (vg_app.cpp:159-160); the four "columns" of `m` are therefore m0..m3 of TransPosIn.
    void setMVP(const glm::mat4 &m) override {
        need_ctx();
        float rows[16];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) rows[4 * i + j] = m[i][j];
        check(slpr_set_mvp(_ctx, rows));
    }

    // scanline_rasterizer.cpp:57-65,282-696. Asynchronous: returns once the frame is enqueued.
    void render() override { need_ctx(); check(slpr_render(_ctx)); }

    // ---- headless additions -----------------------------------------------------------------
    // The image a user of the reference sees: RGBA8, top-left origin, `stride` bytes per row.
    void readback(uint8_t *rgba, size_t stride_bytes) { need_ctx(); check(slpr_readback(_ctx, rgba, stride_bytes)); }
    std::vector<uint8_t> readback() {
        std::vector<uint8_t> img((size_t)_width * _height * 4);
        readback(img.data(), (size_t)_width * 4);
        return img;
    }
    // Frame sequences: setMVP + render + copy to (pinned) host memory, the copy of one frame overlapping the
    // rendering of the next; every submitted frame is whole and in its buffer after waitFrames().
    void submitFrame(const glm::mat4 &m, uint8_t *rgba, size_t stride_bytes) {
        need_ctx();
        float rows[16];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) rows[4 * i + j] = m[i][j];
        check(slpr_submit_to_host(_ctx, rows, rgba, stride_bytes));
    }
    void waitFrames() { need_ctx(); check(slpr_wait_host(_ctx)); }
    void setBand(uint32_t y_begin, uint32_t y_end) { need_ctx(); check(slpr_set_band(_ctx, y_begin, y_end)); }
    // SLPR_FLAG_FULL_RVG (f-1): the middle weights of the scene's ARC curves, after loadVG
    void setCurveWeights(const std::vector<float> &w) { need_ctx(); check(slpr_set_curve_weights(_ctx, w.data(), (uint32_t)w.size())); }
    void counts(uint32_t &n_fragments, uint32_t &n_out_fragments, uint32_t &n_spans) {
        need_ctx(); check(slpr_get_counts(_ctx, &n_fragments, &n_out_fragments, &n_spans));
    }
    slpr_ctx *handle() { return _ctx; }
    uint32_t width() const { return _width; }
    uint32_t height() const { return _height; }

private:
    void need_ctx() const { if (!_ctx) throw std::runtime_error("CudaVGRasterizer: initialize() has not been called"); }
    static void check(int rc) { if (rc) throw std::runtime_error(slpr_last_error()); }
    slpr_ctx *_ctx = nullptr;
    int _device;
    uint32_t _flags, _width = 0, _height = 0;
};

}  // namespace Galaxysailing
