#!/usr/bin/env python
"""bench.py — headline benchmark of the scanline path-rendering hot path (BASELINE.json).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload NAME] [--mode frames|bands]

A step is one frame: set_mvp + render of the resident scene (as the reference: loadVG once, then
{setMVP; render} per frame, VkScanlinePR/src/app/vg_app.cpp:146-165). Default workload is
BASELINE cfg3 `synth_1m_4k` (1 048 576 cubic curves in 262 144 paths, 3840x2160): the configuration
the north-star target is quoted on. Metric: Mpixel/s (= W*H / frame time), ms_per_step = ms/frame.

N > 1 (torchrun, one process per GPU): frame-parallel — every rank renders its own K frames of the
resident scene, no data-path collective ("scaling": "weak"). `--mode bands` instead splits ONE
frame into N row bands and gathers them to rank 0 with NCCL (BASELINE cfg4; "strong").
`--mode anim` is BASELINE cfg5: a 256-frame animation (per-frame affine matrices, scene.anim_rows) of the
workload, frame f on rank f mod N; step i of rank r renders frame r + i*N (use --steps 256/N for one cycle).

`--impl reference` times the reference's algorithm on the host cores (the oracle port: the
reference has no CPU implementation and cannot run without Vulkan), same workload and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder, width, height)
    "synth_1m_4k": (lambda S: S.synth_1m_4k(), 3840, 2160),
    "synth_64k_1080p": (lambda S: S.synth_scene(16384, 1920, 1080, 6.0, 30.0, 0x5CA71E01, "synth_64k_1080p"), 1920, 1080),
    "synth_16k": (lambda S: S.synth_16k(), 16384, 16384),
}


def load_workload(name):
    from vkscanlinepr_b200 import scene as S
    if name in WORKLOADS:
        b, W, H = WORKLOADS[name]
        return b(S), S.identity_rows(), W, H
    # shipped scene from the golden fixtures: "<scene>@<W>x<H>"
    scene, _, size = name.partition("@")
    W, H = (int(v) for v in (size or "3840x2160").split("x"))
    c = S.Container.from_npz(os.path.join(ROOT, "tests", "golden", scene + ".npz"))
    return S.flatten_reference(c, scene), S.fit_rows(c.vp, W, H), W, H


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, str(gpu_index)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.idx], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/*_traffic.json)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.endswith("_traffic.json"):
            d = json.load(open(os.path.join(pdir, f)))
            for k, v in d.get("dram_bytes_per_launch", {}).items():  # "k_spans<1>" is k_spans with the fill fused in
                if d.get("workload") == workload and (k == kernel or k.startswith(kernel + "<")):
                    best = (v, f)
    return best


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_frame(sc, rows, W, H, threads=None):
    from oracle import oracle_py as O
    t = time.perf_counter()
    r = O.render(sc, rows, W, H, do_fill=True, threads=threads, keep={"rgba"})
    return time.perf_counter() - t, r


L2_NOTE = "per-frame working set (fragments x 44 B + 33 MB frame) exceeds the 126 MB L2; no flush needed"


def workload_config(args, sc, W, H, world):
    """The `config` object of BOTH arms' lines (the driver compares them): what is rendered and how it is spread over the
    GPUs. What a run found out about the workload (fragment counts, the sort it chose, ...) goes under `run`."""
    bands, anim = args.mode == "bands" and world > 1, args.mode == "anim"
    par = ((("bands%d" % world) + ("+independent" if args.independent_bands else "+nccl-allgather-of-winding-sums")
            + ("+no-gather" if args.no_gather else "+pipelined-nccl-gather")
            + ("" if args.rank0_share == 1.0 else "+rank0-share-%.2f" % args.rank0_share)) if bands
           else (("anim%d-frame-f-on-rank-f-mod-%d" % (args.anim_frames, world)) if anim else ("frames-dp%d" % world)))
    return {"workload": args.workload, "width": W, "height": H, "curves": sc.n_curves, "paths": sc.n_paths, "points": sc.n_points,
            "scene_sha256": sc.sha256()[:16], "parallelism": par, "l2": L2_NOTE}


def run_reference(args):
    """Reference arm: the reference's algorithm (oracle port) on all host cores, full frames."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O
    sc, rows, W, H = load_workload(args.workload)
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would
    # otherwise make the N > 1 reference lines single-threaded; rank 0 alone runs this arm
    cores = host_threads()
    O.lib().orc_set_num_threads(cores)
    cores = O.num_threads()
    t0, _ = cpu_frame(sc, rows, W, H)  # first frame doubles as warm-up and as the size probe
    budget_s = 150.0
    steps = max(1, min(args.steps, int(budget_s / max(t0, 1e-3))))
    # the requested warm-up (the probe frame is its first frame) as far as the budget goes
    warm = max(0, min(args.warmup - 1, int(budget_s / max(t0, 1e-3)) - steps))
    for _ in range(warm):
        cpu_frame(sc, rows, W, H)
    t = time.perf_counter()
    for _ in range(steps):
        cpu_frame(sc, rows, W, H)
    dt = (time.perf_counter() - t) / steps
    mpix = W * H / dt / 1e6
    sample = f"full frames of {args.workload} ({steps} timed, capped for a {budget_s:.0f}s budget; requested {args.steps})"
    print(json.dumps({
        "impl": "reference", "metric": "Mpixel/s", "value": mpix, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm + 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": workload_config(args, sc, W, H, args.gpus),
        "cpu_baseline": {"value": mpix, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mpix, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def primitive_rooflines(r, stream, nf, key_bits, peak, iters=10):
    """The two roofline-graded primitives timed alone through the C ABI (slpr_scan_i32, slpr_sort_pairs) at the
    frame's sizes: algorithmic bytes (scan 8 B/element; radix sort n * [K + p * 2 * (K + 4)]) / time
    / measured peak. CUDA events on the launching stream, one pair per call; inputs are larger than L2 or re-created
    per call (the sort input is copied fresh outside the timed pair: sorting sorted keys would flatter the scatter)."""
    import torch
    out = {}

    def timed(fn, prep=None):
        ms = []
        for i in range(iters + 3):
            if prep:
                prep()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            if i >= 3:
                ms.append(e0.elapsed_time(e1))
        return float(np.median(ms)), float(min(ms))

    for label, n in (("scan_nf", nf), ("scan_2nf", 2 * nf), ("scan_35M", 35_000_000)):
        a = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")
        b = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        med, best = timed(lambda: r.scan_i32(a.data_ptr(), b.data_ptr(), n))
        # slpr_scan_i32 takes the TMA-pipelined kernel from 2^22 elements on (csrc/slpr.cu), the ticketed look-back kernel below
        out[label] = {"kernel": "k_scan_tma" if n >= (1 << 22) else "k_lookback_scan", "n": n, "bytes": 8 * n, "ms": med, "ms_best": best,
                      "achieved": 8 * n / (med * 1e-3) / 1e9, "frac": 8 * n / (med * 1e-3) / 1e9 / peak}
        del a, b
    passes = (key_bits + 7) // 8
    n = nf
    k0 = torch.randint(0, 1 << key_bits, (n,), dtype=torch.int64, device="cuda")
    v0 = torch.arange(n, dtype=torch.int32, device="cuda")
    k, v, kt, vt = (torch.empty_like(k0), torch.empty_like(v0), torch.empty_like(k0), torch.empty_like(v0))

    def prep():
        k.copy_(k0)
        v.copy_(v0)

    med, best = timed(lambda: r.sort_pairs(k.data_ptr(), v.data_ptr(), kt.data_ptr(), vt.data_ptr(), n, key_bits), prep)
    nbytes = n * (8 + passes * 2 * 12)
    out["sort_pairs_nf"] = {"kernel": "k_radix_hist + k_onesweep x %d" % passes, "n": n, "key_bits": key_bits, "passes": passes,
                            "bytes": nbytes, "ms": med, "ms_best": best, "achieved": nbytes / (med * 1e-3) / 1e9,
                            "frac": nbytes / (med * 1e-3) / 1e9 / peak}
    return out


def d2h_probe(W, H, stream, dist, host=None, iters=20):
    """Host-memory ceiling of the end-to-end path: every rank copies one RGBA8 frame device -> pinned host memory,
    back to back, all ranks at once; per-rank GB/s (min / mean over ranks). Nothing is rendered."""
    import torch
    dev = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    if host is None:
        host = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)
    for _ in range(3):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        host.copy_(dev, non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    gbs = iters * W * H * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if dist is None:
        return {"per_rank_gbs_min": gbs, "per_rank_gbs_mean": gbs, "aggregate_gbs": gbs}
    t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
    lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    sm = t.clone(); dist.all_reduce(sm)
    return {"per_rank_gbs_min": float(lo.item()), "per_rank_gbs_mean": float(sm.item()) / dist.get_world_size(),
            "aggregate_gbs": float(sm.item())}


class BandRunner:
    """One frame split into `world` exact row bands, device-side exchange (csrc/bands.cuh second half): per frame ONE
    graph launch per rank, no host synchronisation, no collective call; the few non-zero per-path winding sums travel
    as peer stores over NVLink. gather="peer": every band's k_resolve stores its pixels straight into the root's
    peer-mapped frame buffer (two buffers, alternating); the root orders its stream after the arrival of frame i once
    frame i+1 is enqueued. gather="none": bands stay on their GPUs."""

    def __init__(self, V, PAR, sc, rows, W, H, rank, world, local_rank, dist, stream, rank0_share=1.0):
        self.V, self.PAR, self.W, self.H, self.rank, self.world, self.dist, self.stream = V, PAR, W, H, rank, world, dist, stream
        self.rows = rows
        self.r = V.ScanlineRasterizer(local_rank, 0).initialize(None, W, H)
        self.r.set_stream(stream.cuda_stream)
        self.r.loadVG(sc)
        self.r.setMVP(rows)
        self.band_list = PAR.band_rows(H, world, rank0_share)
        self.r.set_band(*self.band_list[rank])
        self.peers = PAR.connect_band_peers(self.r, dist, rank, world, root=0, frame_bytes=W * H * 4, n_frames=2)
        self.seq = 1
        self.retries = 0
        self.last_slot = 0
        self.mode = "store"
        y0, y1 = self.band_list[rank]
        self.local = None
        if rank != 0:  # gather by copy engine: the band is rendered into a local buffer (addressed like a full frame) and pushed
            self.local = [self.r.alloc_device((y1 - y0) * W * 4) - (H - y1) * W * 4 for _ in range(2)]

    def _set_mode(self, gather):
        # "store": k_resolve writes into the root's peer-mapped frame and the graph ends with the arrival flag;
        # "copy": the flag follows the copy-engine push (slpr_band_push), so the in-graph flag is switched off
        want = "copy" if gather == "copy" else "store"
        if want != self.mode:
            self.r.set_band_peers(self.world, self.rank, -1 if want == "copy" else 0, self.peers["mailboxes"])
            self.mode = want

    def _step(self, gather, first):
        r = self.r
        slot = self.seq & 1
        if gather == "store" or (gather == "copy" and self.rank == 0):
            r.set_target(self.peers["frames"][slot], self.W * 4)
        elif gather == "copy":
            r.set_target(self.local[slot], self.W * 4)
        else:
            r.set_target(0, 0)
        if gather != "none":
            self.last_slot = slot
        r.setMVP(self.rows)
        r.render_band(self.seq)
        if gather == "copy" and self.rank != 0:
            r.band_push(self.seq, 0, self.peers["frames"][slot], self.W * 4)
        if gather != "none" and self.rank == 0 and not first:
            r.band_wait_gather(self.seq - 1)
        self.seq += 1

    def _finish(self, gather):
        if gather != "none" and self.rank == 0:
            self.r.band_wait_gather(self.seq - 1)

    def run(self, K, Wm, gather):
        import torch
        self._set_mode(gather)
        for _ in range(max(Wm, 3)):  # waited for one by one: capacities and sort modes settle (a void frame is redone by all)
            for _ in range(4):
                self._step(gather, True)
                self._finish(gather)
                if not self.PAR.finish_band_frame(self.r, self.dist):
                    break
                self.retries += 1
        self.dist.barrier()
        torch.cuda.synchronize()
        l0 = self.r.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for i in range(K):
            self._step(gather, i == 0)
        self._finish(gather)
        e1.record(self.stream)
        void = self.PAR.finish_band_frame(self.r, self.dist)
        self.dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        lt = torch.tensor([self.r.launch_count() - l0], device="cuda", dtype=torch.int64)
        self.dist.all_reduce(lt)
        nf = torch.tensor([self.r.counts()["n_fragments"]], device="cuda", dtype=torch.int64)
        nf_max = nf.clone(); self.dist.all_reduce(nf_max, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(nf)
        return {"ms_per_step": float(t.item()) / K, "steps": K, "launches": int(lt.item()), "void_frame_in_timed_region": bool(void),
                "fragments_sum_over_bands": int(nf.item()), "fragments_largest_band": int(nf_max.item())}

    def check_against_full_frame(self, sc, K):
        """Rank 0: the same frame rendered whole on this GPU — its time (the 1-GPU reference of the speed-up) and the
        number of pixels in which the assembled frame of the last gathered step differs from it."""
        import torch
        V = self.V
        full = V.ScanlineRasterizer(self.r._device, 0).initialize(None, self.W, self.H)
        full.set_stream(self.stream.cuda_stream)
        full.loadVG(sc)
        full.setMVP(self.rows)
        for _ in range(3):
            full.render()
        full.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(K):
            full.render()
        e1.record(self.stream)
        full.synchronize()
        ms = e0.elapsed_time(e1) / K
        fb, _ = full.framebuffer()
        diff = full.diff_u32(self.peers["frames"][self.last_slot], fb, self.W * self.H)
        nf = full.counts()["n_fragments"]
        full.close()
        return ms, diff, nf

    def close(self):
        self.r.close()


def bands_peer_report(V, PAR, sc, rows, W, H, rank, world, local_rank, dist, stream, K, Wm, workload, rank0_share=1.0, gather=True):
    """Resident and gathered exact bands + the 1-GPU full frame on rank 0: the north star's strong-scaling case."""
    br = BandRunner(V, PAR, sc, rows, W, H, rank, world, local_rank, dist, stream, rank0_share)
    resident = br.run(K, Wm, "none")
    g_store = br.run(K, Wm, "store") if gather else None
    g_copy = br.run(K, Wm, "copy") if gather else None
    gathered = None
    if gather:
        gathered = g_copy if g_copy["ms_per_step"] <= g_store["ms_per_step"] else g_store
    rep = None
    ms1 = diff = nf1 = None
    if rank == 0:
        ms1, diff, nf1 = br.check_against_full_frame(sc, max(3, min(K, 10))) if gather else (None, None, None)
    dist.barrier()
    if rank == 0:
        rep = {"workload": workload, "width": W, "height": H, "curves": sc.n_curves, "paths": sc.n_paths, "bands": world,
               "fragments_full_frame": nf1, "ms_1gpu": ms1, "ms_resident": resident["ms_per_step"],
               "ms_gathered": gathered["ms_per_step"] if gathered else None,
               "speedup_resident": (ms1 / resident["ms_per_step"]) if ms1 else None,
               "speedup_gathered": (ms1 / gathered["ms_per_step"]) if (ms1 and gathered) else None,
               "ms_gathered_direct_stores": g_store["ms_per_step"] if g_store else None,
               "ms_gathered_copy_engine": g_copy["ms_per_step"] if g_copy else None,
               "pixels_differing": diff, "pixels": W * H, "resident": resident, "gathered": gathered, "retried_warmup_frames": br.retries,
               "exchange": "sparse per-path winding sums stored into peer-mapped mailboxes (CUDA IPC over NVLink), no collective, no host sync",
               "gather": "two ways, both into rank 0's peer-mapped frame buffers (two, alternating; the arrival of frame i is awaited after frame i+1 is enqueued): "
                         "direct stores — k_resolve of every band writes straight into rank 0's HBM; copy engine — the band is rendered locally and pushed "
                         "(slpr_band_push) on a second stream while the next frame renders. ms_gathered is the faster one; pixels_differing checks the last gathered frame",
               "timing": "CUDA events on each rank's stream around K back-to-back frames, max over ranks"}
    br.close()
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synth_1m_4k")
    ap.add_argument("--mode", default="frames", choices=["frames", "bands", "anim"])
    ap.add_argument("--anim-frames", type=int, default=256, help="anim mode: length of the animation cycle (BASELINE cfg5: 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--independent-bands", action="store_true", help="bands mode: no winding-sum exchange (exact only without winding residues)")
    ap.add_argument("--rank0-share", type=float, default=1.0, help="bands mode: rank 0 renders this fraction of an equal band (it also receives the gathered frame)")
    ap.add_argument("--no-gather", action="store_true", help="bands mode: leave every band on its GPU (no NCCL gather)")
    ap.add_argument("--no-radix-leg", action="store_true", help="skip timing the radix sort beside the segmented sort")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="bands mode: device-side sparse exchange over peer-mapped mailboxes (default) or the host-driven NCCL all-gather of dense sums")
    ap.add_argument("--no-bands16k", action="store_true", help="N > 1, frames mode: skip the cfg4 row-band leg (bands16k in the JSON line)")
    ap.add_argument("--no-scenes", action="store_true", help="N = 1, frames mode: skip the shipped-scene leg (scenes_4k in the JSON line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import vkscanlinepr_b200 as V
    from vkscanlinepr_b200 import parallel as PAR

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sc, rows, W, H = load_workload(args.workload)
    K, Wm = args.steps, args.warmup
    # a dedicated (non-default) torch stream: the context launches on it, torch events time it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    if args.mode == "bands" and world > 1 and args.exchange == "peer" and not args.independent_bands:
        rep = bands_peer_report(V, PAR, sc, rows, W, H, rank, world, local_rank, dist, stream, K, Wm, args.workload,
                                args.rank0_share, gather=not args.no_gather)
        if rank == 0:
            ms = rep["ms_gathered"] if rep["ms_gathered"] else rep["ms_resident"]
            print(json.dumps({"metric": "Mpixel/s", "value": W * H / (ms * 1e-3) / 1e6, "unit": "Mpixel/s", "n_gpus": world, "steps": K,
                              "warmup": Wm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                              "dtype": "f32+i32", "data": "synthetic",
                              "config": {"workload": args.workload, "width": W, "height": H, "parallelism": "bands%d+peer-exchange%s" % (world, "" if args.no_gather else "+peer-store-gather")},
                              "gpu_launches": (rep["gathered"] or rep["resident"])["launches"], "bands": rep}))
        dist.destroy_process_group()
        return

    def make_ctx(flags):
        r = V.ScanlineRasterizer(local_rank, flags).initialize(None, W, H)
        r.set_stream(stream.cuda_stream)
        r.loadVG(sc)
        r.setMVP(rows)
        return r

    r = make_ctx(0)
    anim = args.mode == "anim"
    if anim:
        from vkscanlinepr_b200 import scene as S
        base64 = np.asarray(rows, np.float64)
        my_frames = list(range(rank, args.anim_frames, world)) or [0]
        anim_mats = [np.ascontiguousarray((S.anim_rows(f, W, H, args.anim_frames).astype(np.float64) @ base64).astype(np.float32))
                     for f in my_frames]
        anim_step = [0]
        # one untimed pass over this rank's frames, waited for frame by frame: the fragment buffers grow to the
        # largest frame of the cycle and the sort mode settles, so no timed frame has to be rendered twice
        for m in anim_mats:
            r.setMVP(m)
            r.render()
            r.synchronize()
        anim_frag = []
    bands = args.mode == "bands" and world > 1
    frame = None
    if bands:
        band_list = PAR.band_rows(H, world, args.rank0_share)
        y0, y1 = band_list[rank]
        r.set_band(y0, y1)
        # Pipelined gather: every rank renders its band straight into frame buffer i & 1 (two captured graphs,
        # slpr_set_target); the band of frame i then travels to rank 0 (one NCCL group of send/recv over NVLink,
        # on NCCL's own stream) while frame i+1 is rendered into the other buffer. A buffer is reused only after
        # the transfer that read / wrote it two frames earlier has finished. Rank 0 receives in place: no copies.
        if rank == 0:
            frames2 = [torch.empty((H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
        else:  # senders only need their own band, twice
            bands2 = [torch.empty((y1 - y0, W, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
        pending = [[], []]
        step_no = [0]
        # exact bands: per-path winding sums are all-gathered between the two halves of the frame (csrc/bands.cuh)
        x_sums = torch.zeros(3 * sc.n_paths, dtype=torch.int32, device="cuda")
        x_gathered = torch.zeros((world, 3 * sc.n_paths), dtype=torch.int32, device="cuda")
        comm_stream = torch.cuda.Stream()
        if not args.independent_bands:
            r.set_band_exchange(x_sums.data_ptr(), x_gathered.data_ptr(), world, rank)

    def band_target(slot):
        """(pointer the context renders to so that its band lands in the buffer of `slot`, the band view)"""
        if rank == 0:
            return frames2[slot].data_ptr(), frames2[slot][H - y1:H - y0]
        # the context addresses a full frame: offset the base so that image rows [H-y1, H-y0) hit the band buffer
        return bands2[slot].data_ptr() - (H - y1) * W * 4, bands2[slot]

    def drain(slot):
        for q in pending[slot]:
            q.wait()
        pending[slot] = []

    def step():
        if anim:
            r.setMVP(anim_mats[anim_step[0] % len(anim_mats)])
            anim_step[0] += 1
        else:
            r.setMVP(rows)
        if bands:
            slot = step_no[0] & 1
            step_no[0] += 1
            drain(slot)  # the transfer of frame i-2 used this buffer
            ptr, view = band_target(slot)
            r.set_target(ptr, W * 4)
            if not args.independent_bands:
                PAR.render_bands_exact(r, x_sums, x_gathered, dist, comm_stream)
            else:
                r.render()
        else:
            r.render()
        if bands and not args.no_gather:
            if rank == 0:
                ops = []
                for g in range(1, world):
                    gy0, gy1 = band_list[g]
                    ops.append(dist.P2POp(dist.irecv, frames2[slot][H - gy1:H - gy0], g))
                pending[slot] = dist.batch_isend_irecv(ops)  # one NCCL group: the receives run concurrently
            else:
                pending[slot] = dist.batch_isend_irecv([dist.P2POp(dist.isend, view, 0)])

    for _ in range(Wm):
        step()
    if bands:
        drain(0)
        drain(1)
    r.synchronize()
    cnt = r.counts()
    info = r.sort_info()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps, device time, max over ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = r.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    if bands:  # every band of every timed frame has arrived before the clock stops
        drain(0)
        drain(1)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = r.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    frames_total = K if bands else K * world
    ms_per_step = ms_total / K
    mpix = frames_total * W * H / (ms_total * 1e-3) / 1e6

    # ---- per-frame distribution (median / p10 / p90), in a second loop of K frames with one event
    #      per frame so that the timed region above stays exactly K back-to-back steps
    frame_dist = None
    if rank == 0 and not bands:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        evs[0].record(stream)
        for i in range(K):
            step()
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        per = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(K)])
        frame_dist = {"median": float(np.median(per)), "p10": float(np.percentile(per, 10)), "p90": float(np.percentile(per, 90)),
                      "max": float(per.max()), "frames": K}
    if dist is not None:
        dist.barrier()

    anim_info = None
    if anim:
        # every frame of this rank once more, waited for: the fragment counts of the cycle
        for m in anim_mats:
            r.setMVP(m)
            r.render()
            anim_frag.append(r.counts()["n_fragments"])
        anim_info = {"frames_in_cycle": args.anim_frames, "frames_of_this_rank": len(anim_mats),
                     "fragments_min": int(min(anim_frag)), "fragments_max": int(max(anim_frag)),
                     "fragments_mean": float(np.mean(anim_frag))}
        anim_step[0] = 0

    band_check = None
    if bands and rank == 0 and not args.no_gather:  # the assembled frame against the same frame rendered whole on this GPU
        full = make_ctx(0)
        ref_frame = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
        full.set_target(ref_frame.data_ptr(), W * 4)
        full.render()
        full.synchronize()
        last = frames2[(step_no[0] - 1) & 1]
        diff = int((last.view(torch.int32) != ref_frame.view(torch.int32)).sum().item())
        band_check = {"pixels_differing": diff, "pixels": W * H,
                      "note": "0 unless the scene has a winding residue (DESIGN.md section 5)"}
        full.close()
        del ref_frame

    # ---- e2e: the public calls with HOST buffers (matrix in, RGBA8 frame out to pinned memory).
    #      Headline: the pipelined frame-sequence call (slpr_submit_to_host: the copy of frame i overlaps the
    #      rendering of frame i+1; every frame's pixels land in host memory). Also the blocking per-frame call.
    # host buffers: pinned, and placed on the GPU's NUMA node by the library (slpr_host_alloc; torch-pinned as a fall-back)
    host_arrays, host_nodes = [], []
    try:
        for _ in range(2):
            a, node = r.host_alloc((H, W, 4))
            host_arrays.append(a)
            host_nodes.append(node)
        host_alloc = "slpr_host_alloc"
    except Exception:
        host_arrays = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
        host_nodes, host_alloc = [-1, -1], "torch pin_memory"
    hosts = [torch.from_numpy(a) for a in host_arrays]
    host_np = host_arrays[0]
    rows_host = np.ascontiguousarray(rows, dtype=np.float32)
    if not bands:
        for _ in range(2):
            r.render_to_host(rows_host, host_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            r.render_to_host(rows_host, host_np)
        barrier()
        sync_s = time.perf_counter() - t0
        def rows_of(i):
            return anim_mats[i % len(anim_mats)] if anim else rows_host
        for i in range(4):
            r.submit_to_host(rows_of(i), host_arrays[i & 1])
        r.wait_host()
        barrier()
        redone0 = r.pipeline_redone()
        t0 = time.perf_counter()
        for i in range(K):
            r.submit_to_host(rows_of(i), host_arrays[i & 1])
        r.wait_host()
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_redone = r.pipeline_redone() - redone0
        if dist is not None:
            t = torch.tensor([e2e_s, sync_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, sync_s = float(t[0].item()), float(t[1].item())
        e2e = {"value": frames_total * W * H / e2e_s / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_s / K * 1e3,
               "h2d_bytes_per_step": 80, "d2h_bytes_per_step": W * H * 4,
               "what": "slpr_submit_to_host per frame + slpr_wait_host: TransPosIn rows from host, render, RGBA8 frame "
                       "to pinned host memory, copy of frame i overlapped with rendering of frame i+1",
               "blocking_call": {"value": frames_total * W * H / sync_s / 1e6, "ms_per_step": sync_s / K * 1e3,
                                 "what": "slpr_render_to_host (set_mvp + render + readback, synchronous per frame)"},
               "frames_rendered_twice": int(e2e_redone),
               "host_buffers": {"allocator": host_alloc, "numa_node": host_nodes[0], "torch_sees_pinned": bool(hosts[0].is_pinned())},
               "d2h_probe": {"torch_pinned": d2h_probe(W, H, stream, dist, None), "gpu_numa_node": d2h_probe(W, H, stream, dist, hosts[0]),
                             "what": "every rank copies one RGBA8 frame device -> pinned host, 20 times back to back, all ranks at once, nothing rendered: "
                                     "the host-side ceiling of the end-to-end path (max-over-ranks timing follows the slowest rank)"}}
        floor = W * H * 4 / (e2e["d2h_probe"]["gpu_numa_node"]["per_rank_gbs_min"] * 1e9) * 1e3
        e2e["d2h_floor_ms_per_step"] = floor  # a frame cannot reach host memory faster than the slowest rank's copy
        e2e["of_d2h_floor"] = floor / e2e["ms_per_step"]
    else:
        e2e = None

    # ---- roofline: per-kernel times from CUDA events around the stages in direct-launch mode, same K
    #      steps, on the launching stream. The headline object is the kernel with the largest share
    #      of the frame; the sort and the scans (the north star's HBM-bound stages) are listed too.
    roof = stage_avg = None
    if rank == 0 and not bands and not anim:
        def stage_times(flags):
            ri = make_ctx(flags)
            for _ in range(3):
                ri.render()
            ri.synchronize()
            acc = {}
            for _ in range(K):
                ri.render()
                for k, v in ri.stage_ms().items():
                    acc[k] = acc.get(k, 0.0) + v
            out = {k: v / K for k, v in acc.items()}, ri.sort_mode(), ri.n_pieces(), ri.sort_info(), ri.fill_fused()
            ri.close()
            return out

        stage_avg, mode, n_pieces, _, fused = stage_times(V.FLAG_NO_GRAPH)
        nf = cnt["n_fragments"]
        nrec = cnt["n_out_frag"] + cnt["n_span"]
        peak, peak_src = peaks()

        def entry(kernel, nbytes, ms, launches=1, note=None):
            gbs = nbytes / (ms * 1e-3) / 1e9
            tr = ncu_traffic(kernel, args.workload) or (None, None)
            e = {"kernel": kernel, "bytes_per_launch": nbytes / launches, "ms_per_launch": ms / launches,
                 "launches_per_step": launches, "achieved": gbs, "frac": gbs / peak, "traffic": tr[0], "traffic_source": tr[1]}
            if note:
                e["note"] = note
            return e

        kb = info["key_bytes"]
        kernels = {
            # 64-byte piece records in, (key, value) pairs and the boundary parameters out
            "walk": entry("k_walk", n_pieces * 64 + nf * (kb + 4) + n_pieces * 9, stage_avg["walk"],
                          note="issue-bound, not HBM-bound: 2/3 of its instructions are the reference's 24-step fp32 "
                               "bisection, reproduced operation for operation (no FMA); see DESIGN.md"),
            "spans": entry("k_spans", nf * (kb + 4) + nrec * 16, stage_avg["span_emit"],
                           note="also marks the cells of every record (stage 5 coverage fused in: ~2.4 L2 atomics per record)"
                           if fused else None),
        }
        if not fused:
            kernels["fill"] = entry("k_fill_cells", nrec * 16, stage_avg["fill_cells"])
        if mode == "segmented":  # one read + one write of every pair
            kernels["sort"] = entry("k_segsort_warp", nf * 2 * (kb + 4), stage_avg["sort_passes"])
        else:                    # 2*(K+4) B per fragment per pass
            kernels["sort"] = entry("k_onesweep", nf * 2 * (kb + 4) * info["passes"], stage_avg["sort_passes"], info["passes"])
        dominant = max(kernels.values(), key=lambda e: e["ms_per_launch"] * e["launches_per_step"])
        # In the frame, scan #1 is k_lookback_scan over the curve counts (4 B read + 4 B written per curve); the
        # winding scan is split: k_wsum + k_wscan only READ the 4-byte values (tile sums; timed here), its in-tile
        # half runs inside k_spans (timed there). Bytes are the ones these kernels move, not 8 B per element.
        n_tiles = (nf + 4095) // 4096
        scan_bytes = 8 * sc.n_curves + 4 * nf + 8 * n_tiles
        scan_ms = stage_avg["scan1"] + stage_avg["wind_scan"]
        roof = {"bound": "hbm", "kernel": dominant["kernel"], "achieved": dominant["achieved"], "peak": peak, "unit": "GB/s",
                "frac": dominant["frac"], "traffic": dominant["traffic"], "traffic_source": dominant["traffic_source"],
                "peak_source": peak_src, "launches_per_step": dominant["launches_per_step"],
                "ms_per_launch": dominant["ms_per_launch"], "bytes_per_launch": dominant["bytes_per_launch"],
                "note": dominant.get("note"),
                "timing": "cudaEvent pairs around each stage, direct-launch mode, averaged over the same K steps",
                "sort_mode": mode, "kernels": kernels,
                "scans": {"what": "in-frame: k_lookback_scan over curve counts (8 B/curve) + k_wsum/k_wscan tile sums (4 B/fragment, "
                                  "read only); the stand-alone scan primitive is under `primitives`",
                          "bytes": scan_bytes, "ms": scan_ms, "gbs": scan_bytes / (scan_ms * 1e-3) / 1e9,
                          "frac": scan_bytes / (scan_ms * 1e-3) / 1e9 / peak,
                          "scan1": {"bytes": 8 * sc.n_curves, "ms": stage_avg["scan1"],
                                    "frac": 8 * sc.n_curves / (stage_avg["scan1"] * 1e-3) / 1e9 / peak},
                          "wind_prefix": {"bytes": 4 * nf + 8 * n_tiles, "ms": stage_avg["wind_scan"],
                                          "frac": (4 * nf + 8 * n_tiles) / (stage_avg["wind_scan"] * 1e-3) / 1e9 / peak}},
                "primitives": primitive_rooflines(r, stream, nf, info["key_bits"], peak)}
        if mode == "segmented" and not args.no_radix_leg:
            # the general sort (paths of any length), timed on the same frame for comparison
            st_r, _, _, info_r, _ = stage_times(V.FLAG_NO_GRAPH | V.FLAG_RADIX_SORT)
            pass_ms = st_r["sort_passes"] / info_r["passes"]
            pass_bytes = nf * 2 * (info_r["key_bytes"] + 4)
            tr = ncu_traffic("k_onesweep", args.workload) or (None, None)
            roof["radix_sort"] = {"kernel": "k_onesweep", "launches_per_step": info_r["passes"], "ms_per_launch": pass_ms,
                                  "bytes_per_launch": pass_bytes, "achieved": pass_bytes / (pass_ms * 1e-3) / 1e9,
                                  "frac": pass_bytes / (pass_ms * 1e-3) / 1e9 / peak, "traffic": tr[0], "traffic_source": tr[1],
                                  "sort_total_ms": st_r["sort_hist"] + st_r["sort_passes"],
                                  "frame_ms_direct": sum(st_r.values())}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_py as O
        O.lib().orc_set_num_threads(host_threads())
        cores = O.num_threads()
        ts = []
        t_budget = time.perf_counter()
        cpu_rows = anim_mats[min(37, len(anim_mats) - 1)] if anim else rows
        while len(ts) < 3 and (time.perf_counter() - t_budget) < 20.0:
            dt, ref = cpu_frame(sc, cpu_rows, W, H)
            ts.append(dt)
        r.render_to_host(np.ascontiguousarray(cpu_rows, dtype=np.float32), host_np)
        ok = bool(np.array_equal(host_np, ref["rgba"])) and ref["n_fragments"] == r.counts()["n_fragments"]
        cpu = {"value": W * H / min(ts) / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": "port",
               "sample": f"{len(ts)} full frame(s) of {args.workload}, best; all {cores} OpenMP threads",
               "ms_per_frame": min(ts) * 1e3, "gpu_frame_matches_oracle": ok}

    sort_now, fused_now, long_now = r.sort_mode(), r.fill_fused(), r.long_walk_info()
    scenes_4k = None
    if world == 1 and args.mode == "frames" and args.workload == "synth_1m_4k" and not args.no_scenes:
        # BASELINE cfg2 in the driver-run line: the reference's shipped scenes at 3840 x 2160 — few, long
        # curves, the opposite regime of the headline — device time per frame (graph replay) and the frame against the oracle.
        scenes_4k = {}
        try:
            from oracle import oracle_py as O2
            for name in ("test", "tiger", "reschart", "drops", "embrace"):
                s2, rows2, W2, H2 = load_workload(name + "@3840x2160")
                r2 = V.ScanlineRasterizer(local_rank, 0).initialize(None, W2, H2)
                r2.set_stream(stream.cuda_stream); r2.loadVG(s2); r2.setMVP(rows2)
                for _ in range(5):
                    r2.render()
                r2.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for _ in range(20):
                    r2.render()
                a1.record(stream); torch.cuda.synchronize()
                entry = {"ms_per_frame": a0.elapsed_time(a1) / 20, "curves": s2.n_curves, "fragments": r2.counts()["n_fragments"],
                         "long_pieces": r2.long_walk_info()[1], "sort": r2.sort_mode()}
                if not args.no_cpu_baseline:
                    entry["identical_to_oracle"] = bool(np.array_equal(r2.readback(), O2.render(s2, rows2, W2, H2, keep={"rgba"})["rgba"]))
                scenes_4k[name] = entry
                r2.close()
        except Exception as e:  # the headline line stands on its own
            scenes_4k["error"] = repr(e)[:300]
    bands16k = None
    if world > 1 and args.mode == "frames" and not args.no_bands16k:
        # The north star's multi-GPU case in the driver-run line: BASELINE cfg4, one 16384 x 16384 frame
        # of 4 M curves in `world` exact row bands, band-resident and gathered to rank 0, against the same frame on one GPU.
        from vkscanlinepr_b200 import scene as S16
        r.close()
        r = None
        try:
            sc16 = S16.synth_16k()
            bands16k = bands_peer_report(V, PAR, sc16, S16.identity_rows(), 16384, 16384, rank, world, local_rank, dist, stream,
                                         min(K, 20), 3, "synth_16k")
        except Exception as e:  # the frame-batch line above stands on its own; say what happened instead of losing it
            bands16k = {"workload": "synth_16k", "error": repr(e)[:500]}
    if rank == 0:
        out = {
            "metric": "Mpixel/s", "value": mpix, "unit": "Mpixel/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if bands else "weak",
            "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "config": workload_config(args, sc, W, H, world),
            "run": {"fragments": cnt["n_fragments"], "records": cnt["n_out_frag"] + cnt["n_span"], "sort": sort_now,
                    "fill": "fused into k_spans" if fused_now else "k_fill_cells", "long_piece_walk": long_now[0], "long_pieces": long_now[1],
                    "sort_key_bits": info["key_bits"], "radix_passes_if_radix": info["passes"], "frame_replay": "cuda-graph"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "stage_ms": stage_avg, "frame_ms": frame_dist,
        }
        if band_check is not None:
            out["bands_vs_full_frame"] = band_check
        if bands16k is not None:
            out["bands16k"] = bands16k
        if scenes_4k is not None:
            out["scenes_4k"] = scenes_4k
        if anim_info is not None:
            out["anim"] = anim_info
        print(json.dumps(out))
    if r is not None:
        r.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
