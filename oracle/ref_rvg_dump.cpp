// ref_rvg_dump.cpp — harness around the REFERENCE's own RVG parser (compiled in place from
// /root/reference/VkScanlinePR/src/core/vg/rvg.cpp by oracle/Makefile; no reference source is
// copied into this repo). Dumps the VGContainer it produces as a flat little-endian binary so
// tests can pin our RVG loader + flattening against it. Test infrastructure only.
//
// Output layout: "VGC1" u32 n_points, n_curves, n_paths; float vp[4]; float win[4];
//   float pos[2*n_points]; u32 curve_pos[n_curves]; u32 curve_type[n_curves];
//   u32 path_curve[n_paths]; u32 fill_rule[n_paths]; float fill_color[4*n_paths]; float opacity[n_paths]
#include <cassert>
#include <cstdio>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <vector>
#include "vg/rvg.h"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: rvg_dump in.rvg out.vgc\n"); return 2; }
    Galaxysailing::RVG rvg;
    try { rvg.load(argv[1]); } catch (std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    auto vg = rvg.getVGContainer();
    uint32_t np = (uint32_t)vg->pointData.pos.size();
    uint32_t nc = (uint32_t)(vg->curveData.curveIndex + 1);
    uint32_t npath = (uint32_t)(vg->pathData.pathIndex + 1);
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) return 1;
    std::fwrite("VGC1", 1, 4, f);
    std::fwrite(&np, 4, 1, f); std::fwrite(&nc, 4, 1, f); std::fwrite(&npath, 4, 1, f);
    float vp[4] = {vg->vp[0], vg->vp[1], vg->vp[2], vg->vp[3]};
    float win[4] = {vg->win[0], vg->win[1], vg->win[2], vg->win[3]};
    std::fwrite(vp, 4, 4, f); std::fwrite(win, 4, 4, f);
    for (auto& p : vg->pointData.pos) { float xy[2] = {p.x, p.y}; std::fwrite(xy, 4, 2, f); }
    for (uint32_t i = 0; i < nc; ++i) { uint32_t v = vg->curveData.posIndices[i]; std::fwrite(&v, 4, 1, f); }
    for (uint32_t i = 0; i < nc; ++i) { uint32_t v = (uint32_t)vg->curveData.curveType[i]; std::fwrite(&v, 4, 1, f); }
    for (uint32_t i = 0; i < npath; ++i) { uint32_t v = vg->pathData.curveIndices[i]; std::fwrite(&v, 4, 1, f); }
    for (uint32_t i = 0; i < npath; ++i) { uint32_t v = (uint32_t)vg->pathData.fillRule[i]; std::fwrite(&v, 4, 1, f); }
    for (uint32_t i = 0; i < npath; ++i) { auto& c = vg->pathData.fillColor[i]; float v[4] = {c.r, c.g, c.b, c.a}; std::fwrite(v, 4, 4, f); }
    for (uint32_t i = 0; i < npath; ++i) { float v = vg->pathData.fillOpacity[i]; std::fwrite(&v, 4, 1, f); }
    std::fclose(f);
    std::printf("points=%u curves=%u paths=%u\n", np, nc, npath);
    return 0;
}
