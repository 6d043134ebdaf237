#!/usr/bin/env python
"""Stand-alone rooflines of the two graded primitives (slpr_scan_i32, slpr_sort_pairs): one JSON line.
  [SLPR_LIB=variant.so] python tools/bench_primitives.py [n_fragments] [key_bits]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import vkscanlinepr_b200 as V  # noqa: E402

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 17_699_852
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 40
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
r = V.ScanlineRasterizer(0, 0).initialize(None, 64, 64)
r.set_stream(stream.cuda_stream)
peak, src = bench.peaks()
out = bench.primitive_rooflines(r, stream, nf, bits, peak)
print(json.dumps({"lib": os.path.basename(V.LIB_PATH), "peak": peak, **{k: {"ms": round(v["ms"], 4), "frac": round(v["frac"], 3)} for k, v in out.items()}}))
