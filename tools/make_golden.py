#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the REFERENCE itself.

Runs here (build container, /root/reference present), never on the GPU box:
  * <scene>.npz   — the VGContainer that the reference's own RVG parser (oracle/_ref/rvg_dump,
                    compiled in place from VkScanlinePR/src/core/vg/rvg.cpp) produces for each
                    shipped workdir/input/rvg/<scene>.rvg;
  * ref_records_{1,3}.npz — the reference's output_buf dumps workdir/test_data.csv and
                    test_data3.csv (yx, width, fill_info, frag_index per record), the only
                    golden data the reference holds for this path.
Usage: python tools/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vkscanlinepr_b200.scene import Container  # noqa: E402

REF = os.environ.get("SLPR_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
SCENES = ["test", "tiger", "reschart", "drops", "embrace", "car", "chord", "chord-black"]


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    dump = os.path.join(ROOT, "oracle", "_ref", "rvg_dump")
    os.makedirs(OUT, exist_ok=True)
    for s in SCENES:
        with tempfile.NamedTemporaryFile(suffix=".vgc") as t:
            subprocess.check_call([dump, os.path.join(REF, "workdir/input/rvg", s + ".rvg"), t.name],
                                  stdout=subprocess.DEVNULL)
            c = Container.from_vgc(t.name)
        c.to_npz(os.path.join(OUT, s + ".npz"))
        print(f"{s}: points={c.pos.shape[0]} curves={len(c.curve_pos)} paths={len(c.path_curve)}")
    for tag, f in (("1", "test_data.csv"), ("3", "test_data3.csv")):
        rec = np.loadtxt(os.path.join(REF, "workdir", f), delimiter=",", dtype=np.int64).astype(np.int32)
        np.savez_compressed(os.path.join(OUT, f"ref_records_{tag}.npz"), records=rec)
        print(f"{f}: {rec.shape[0]} records")


if __name__ == "__main__":
    main()
