// slpr_render.cpp — headless replacement of the reference's main.cpp + ScanlineVGApplication::run()
// (VkScanlinePR/src/main.cpp:8-24, src/app/vg_app.cpp:146-165): same builder-style call sequence,
// CudaVGRasterizer in place of ScanlineVGRasterizer, a PPM file in place of the GLFW window.
//   slpr_render scene.rvg out.ppm [width height [full]]     (full: the complete RVG reader + SLPR_FLAG_FULL_RVG, f-1)
// Build: g++ -std=c++17 -Iinclude tools/slpr_render.cpp -Lvkscanlinepr_b200 -lslpr -Wl,-rpath,$PWD/vkscanlinepr_b200
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "slpr_rasterizer.hpp"

using namespace Galaxysailing;

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s scene.rvg out.ppm [width height]\n", argv[0]); return 2; }
    const uint32_t w = argc > 4 ? (uint32_t)std::atoi(argv[3]) : 1200, h = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 1024;
    const bool full = argc > 5 && std::string(argv[5]) == "full";
    try {
        RVG rvg;
#if !defined(SLPR_WITH_REFERENCE_HEADERS)
        if (full) rvg.loadFull(argv[1]); else
#endif
        rvg.load(argv[1]);                               // app->loadPathFile(...)
        auto vg = rvg.getVGContainer();
        std::cout << "---------- vg load success ---------\n";
        CudaVGRasterizer rast(0, full ? SLPR_FLAG_FULL_RVG : 0);
        rast.initialize(nullptr, w, h);                  // _vgRasterizer->initialize(window, w, h)
        rast.loadVG(vg);                                 // _vgRasterizer->loadVG(_vgContainer)
#if !defined(SLPR_WITH_REFERENCE_HEADERS)
        if (full && !rvg.getCurveWeights().empty()) rast.setCurveWeights(rvg.getCurveWeights());
#endif
        // camera: uniform fit of the file's viewport box, centred; the app passes transpose(camera.mv())
        const float sx = w / (vg->vp[2] - vg->vp[0]), sy = h / (vg->vp[3] - vg->vp[1]), s = std::min(sx, sy);
        glm::mat4 mv(1.0f);                              // identity (glm's default constructor leaves it unset); column-major affine: x' = s*x + tx
        mv[0][0] = s; mv[1][1] = s;
        mv[3][0] = (w - s * (vg->vp[2] - vg->vp[0])) * 0.5f - s * vg->vp[0];
        mv[3][1] = (h - s * (vg->vp[3] - vg->vp[1])) * 0.5f - s * vg->vp[1];
        rast.setMVP(glm::transpose(mv));                 // vg_app.cpp:159-160
        rast.render();                                   // vg_app.cpp:161
        std::vector<uint8_t> img = rast.readback();
        uint32_t nf, nof, ns;
        rast.counts(nf, nof, ns);
        std::printf("fragments=%u merged=%u spans=%u\n", nf, nof, ns);
        FILE *f = std::fopen(argv[2], "wb");
        if (!f) throw std::runtime_error("cannot write output");
        std::fprintf(f, "P6\n%u %u\n255\n", w, h);
        for (size_t i = 0; i < (size_t)w * h; ++i) std::fwrite(&img[4 * i], 1, 3, f);
        std::fclose(f);
    } catch (std::exception &e) {                        // main.cpp:16-23
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
