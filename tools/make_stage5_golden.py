#!/usr/bin/env python
"""Run the REFERENCE's shipped stage-5 shaders — workdir/shaders/scanline/surface/spv/scanlinepr.vert.spv and
scanlinepr.frag.spv — through oracle/spirv_exec.py on the REFERENCE's own draw-record dumps (workdir/test_data.csv,
test_data3.csv: `output_buf` of a real run at the 1200x1024 viewport the vertex shader hard-codes, VERT:39), the way
ScanlineVGRasterizer::drawFrame draws them (scanline_rasterizer.cpp:611-656: LINE_LIST, two vertices per record,
vkCmdDraw(2 * n_records)), and write what the shaders produce to tests/golden/stage5_<tag>.npz:

  position   float32 bits [2n, 4]  gl_Position of vertex 2k + e
  color      float32 bits [n, 4]   fragment shader output for record k (flat: the provoking vertex's fragment_color
                                   through scanlinepr.frag)
  frag_pos   int32 [n, 2]          path_frag_pos of record k (flat varying, location 1)
  sample_mask uint32               gl_SampleMask[0] written by the fragment shader

tests/test_stage5_golden.py pins the oracle's stage 5 (orc_fill) to these: the record decode, the y flip, the colour
unpacking come from the executed shaders; only the fixed-function part (viewport transform, line coverage, UNORM
conversion) is restated from the Vulkan rules there. Runs only where /root/reference exists (a minute of CPU).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import spirv_exec as SX  # noqa: E402

REF = os.environ.get("SLPR_REFERENCE", "/root/reference")
SURF = os.path.join(REF, "workdir", "shaders", "scanline", "surface", "spv")
GOLD = os.path.join(ROOT, "tests", "golden")
BI_POSITION, BI_FRAG_COORD, BI_SAMPLE_MASK, BI_VERTEX_INDEX = 0, 15, 20, 42
F32 = np.float32


def bits(t):
    return [int(np.float32(x).view(np.uint32)) for x in t]


def run_records(rec, log=print):
    """rec int32 [n, 4] -> dict of arrays (see the module docstring)."""
    vert, frag = SX.Module(os.path.join(SURF, "scanlinepr.vert.spv")), SX.Module(os.path.join(SURF, "scanlinepr.frag.spv"))
    texels = np.ascontiguousarray(rec, np.int32).reshape(-1).view(np.uint32)
    rv = SX.Runner(vert, {0: (texels, 0)})  # layout(binding = 0) uniform isamplerBuffer tb_index
    rf = SX.Runner(frag, {})
    n = rec.shape[0]
    position = np.zeros((2 * n, 4), np.uint32)
    color = np.zeros((n, 4), np.uint32)
    frag_pos = np.zeros((n, 2), np.int32)
    masks = set()
    frag_cache = {}
    t = time.time()
    for k in range(n):
        flat = None
        for e in (0, 1):
            ob, ol = rv.run_stage(builtins={BI_VERTEX_INDEX: (2 * k + e,)})
            position[2 * k + e] = bits(ob[BI_POSITION])
            if e == 0:
                flat = ol  # flat varyings: Vulkan's provoking vertex is the first one of the line
            else:
                assert bits(ol[0]) == bits(flat[0]) and ol[1] == flat[1], "both vertices of a record carry the same varyings"
        fp = tuple(SX.s32(x) for x in flat[1])
        frag_pos[k] = fp
        key = (tuple(bits(flat[0])), fp)
        if key not in frag_cache:  # the fragment shader's outputs depend on gl_FragCoord only through a dead value
            fb, fl = rf.run_stage(builtins={BI_FRAG_COORD: (F32(fp[0] + 0.5), F32(fp[1] + 0.5), F32(0), F32(1))},
                                  locations={0: flat[0], 1: flat[1]})
            frag_cache[key] = (bits(fl[0]), fb[BI_SAMPLE_MASK][0])
        color[k], m = frag_cache[key]
        masks.add(m)
    assert len(masks) == 1
    log(f"  {n} records: {rv.instr_count} vertex + {rf.instr_count} fragment instructions, {time.time() - t:.0f}s")
    return dict(position=position, color=color, frag_pos=frag_pos, sample_mask=np.uint32(masks.pop()))


def end_to_end():
    """All eleven shipped shaders on one small scene at the reference's native 1200x1024 viewport: the nine compute
    shaders along drawFrame's dispatch sequence (tools/make_spirv_golden.py run_frame), then the vertex / fragment pair
    over the draw records they produced -> tests/golden/e2e1200.npz (every compute buffer + s5_* = the stage-5 outputs)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import make_spirv_golden as G
    from vkscanlinepr_b200 import scene as S
    sc = S.synth_scene(16, 1200, 1024, 8.0, 30.0, seed=77)
    r = G.run_frame(sc, S.identity_rows(), 1200, 1024, log=lambda *a: None)
    print(f"e2e1200: {sc.n_curves} curves, {r['n_fragments']} fragments, {r['records'].shape[0]} records")
    s5 = run_records(r["records"])
    np.savez_compressed(os.path.join(GOLD, "e2e1200.npz"), **G.scene_arrays(sc), **r, **{"s5_" + k: v for k, v in s5.items()})


def main():
    for tag in ("1", "3"):
        rec = np.load(os.path.join(GOLD, f"ref_records_{tag}.npz"))["records"]
        print(f"stage5_{tag}: {rec.shape[0]} records")
        np.savez_compressed(os.path.join(GOLD, f"stage5_{tag}.npz"), **run_records(rec))
    end_to_end()


if __name__ == "__main__":
    main()
