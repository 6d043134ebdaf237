"""Multi-GPU host logic (new work; the reference is single-device — SURVEY §8e).

The path shards two ways:
  * row bands  — rank r renders scanline rows [y0, y1) of one frame (slpr_set_band). The reference's
    winding scan is one prefix sum over ALL fragments, so the bands exchange one thing: per-path sums
    of their winding deltas (3 * n_paths int32 per band, all-gather; csrc/bands.cuh) between the two
    halves of the frame (render_bands_exact). The finished RGBA8 bands are then gathered to rank 0
    (NCCL send/recv over NVLink, or direct stores into a peer-mapped frame through slpr_set_target).
    Without the exchange (plain render() on a band) the result is exact only for scenes whose per-path
    winding sums vanish;
  * frame batches — frame f goes to rank f mod world; nothing is exchanged.
One process per GPU, torch.distributed for the plumbing. These helpers are backend-agnostic so the
host logic is covered on CPU with gloo (tests/test_parallel_cpu.py).
"""
import numpy as np


def band_rows(height, world, first_share=1.0):
    """Even-aligned scanline-row bands [(y0, y1)] covering [0, height): fragments are 2 px tall, so a band
    edge must be even (slpr_set_band); the last band takes the remainder. `first_share` < 1 gives band 0 that
    fraction of an equal share (the rank that also receives the gathered frame renders less)."""
    if world < 1 or height < 1:
        raise ValueError("world and height must be positive")
    per = (height // world) & ~1
    if per == 0:
        if world > 1:
            raise ValueError(f"height {height} is too small for {world} even-aligned bands")
        return [(0, height)]
    if world == 1 or first_share == 1.0:
        return [(r * per, height if r == world - 1 else (r + 1) * per) for r in range(world)]
    first = max(2, int(per * first_share) & ~1)
    rest = ((height - first) // (world - 1)) & ~1
    edges = [0, first] + [first + k * rest for k in range(1, world - 1)] + [height]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def image_rows(height, y0, y1):
    """Image rows (top-left origin) that hold scanline rows [y0, y1): image row = H-1-scanline row
    (scanlinepr.vert:44), so the band is the contiguous slice [H-y1, H-y0)."""
    return slice(height - y1, height - y0)


def frames_of_rank(n_frames, rank, world):
    """BASELINE cfg5: frame f is rendered by rank f mod world."""
    return list(range(rank, n_frames, world))


def gather_bands(frame, bands, rank, world, dist, dst=0):
    """Collect the bands on `dst`. `frame` is a [H, W, 4] uint8 tensor on every rank (the rank's own band
    rows are valid); rank `dst` receives every other band straight into its rows of `frame`."""
    H = frame.shape[0]
    if world == 1:
        return frame
    if rank == dst:
        reqs = []
        for r in range(world):
            if r == dst:
                continue
            reqs.append(dist.irecv(frame[image_rows(H, *bands[r])], src=r))
        for q in reqs:
            q.wait()
    else:
        dist.send(frame[image_rows(H, *bands[rank])].contiguous(), dst=dst)
    return frame


def scatter_frames(n_frames, world):
    """Which rank renders which frame, as an int array (frame-parallel batches)."""
    return np.arange(n_frames) % world


def render_bands_exact(rasterizer, sums, gathered, dist, comm_stream=None):
    """One exact band of a frame on this rank: fragments + sort, all-gather of the per-path winding sums,
    second half. `sums` [3P] and `gathered` [world, 3P] are int32 tensors on the rasterizer's device,
    registered once with rasterizer.set_band_exchange(sums.data_ptr(), gathered.data_ptr(), world, rank).
    render_band_begin returns when the sums are complete while the sort is still running on the rasterizer's
    stream (torch's current stream); the collective is issued from `comm_stream` so that it overlaps the
    sort, and the current stream waits for it before the second half."""
    import torch
    if gathered.is_cuda and getattr(rasterizer, "_stream_ptr", 0) != torch.cuda.current_stream().cuda_stream:
        # slpr_render_band_end reads `gathered` on the rasterizer's stream; the collective is ordered against torch's
        # CURRENT stream only — they must be the same stream, or the corrections are read while NCCL still writes them
        raise RuntimeError("render_bands_exact: bind the rasterizer to torch's current stream first "
                           "(rasterizer.set_stream(torch.cuda.current_stream().cuda_stream))")
    rasterizer.render_band_begin()
    if comm_stream is not None and gathered.is_cuda:
        with torch.cuda.stream(comm_stream):
            work = dist.all_gather_into_tensor(gathered.view(-1), sums, async_op=True)
        work.wait()  # orders the CURRENT stream after the collective; the host does not block
    else:
        dist.all_gather_into_tensor(gathered.view(-1), sums)
    rasterizer.render_band_end()


def band_corrections(gathered, band):
    """numpy restatement of k_band_other + scan + k_band_corr (csrc/bands.cuh) for tests: gathered is
    [n_bands, 3, P] (normal rows | outside the frame | row 0); returns (corrN, corrZ) of `band`."""
    g = np.asarray(gathered, dtype=np.int64)
    others = np.delete(g, band, axis=0)
    d = others.sum(axis=(0, 1))                                   # total - mine, per path
    e = np.concatenate(([0], np.cumsum(d)[:-1])) if d.size else d  # exclusive scan over paths
    corr_n = e + g[:band, 0].sum(axis=0)
    corr_z = e + others[:, 0].sum(axis=0) + others[:, 1].sum(axis=0)
    return corr_n.astype(np.int32), corr_z.astype(np.int32)


# --------------------------------------------------------------------------------------------------------------
# Exact bands with the device-side sparse exchange (csrc/bands.cuh, second half; include/slpr.h)
def connect_band_peers(rasterizer, dist, rank, world, root=0, frame_bytes=0, n_frames=2, done_flags_in_graph=True):
    """Once per rasterizer: every rank allocates its mailbox, the 64-byte CUDA IPC handles are all-gathered over
    torch.distributed (host plumbing only), every rank maps the others' mailboxes (peer access over NVLink) and
    registers them (slpr_set_band_peers). With `frame_bytes`, the root also allocates `n_frames` frame buffers and
    every rank maps them: a band rendered with set_target(frames[i], stride) stores its pixels straight into the
    root's HBM. Returns {"mailboxes": [...], "frames": [...]} as device pointers valid on this rank's GPU."""
    own, _ = rasterizer.band_mailbox()
    mine = {"box": rasterizer.ipc_export(own) if world > 1 else None, "frames": []}
    frames_local = []
    if frame_bytes and rank == root:
        frames_local = [rasterizer.alloc_device(frame_bytes) for _ in range(n_frames)]
        if world > 1:
            mine["frames"] = [rasterizer.ipc_export(p) for p in frames_local]
    if world > 1:
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
    else:
        everyone = [mine]
    boxes = [own if r == rank else rasterizer.ipc_import(everyone[r]["box"]) for r in range(world)]
    frames = frames_local if rank == root else [rasterizer.ipc_import(h) for h in everyone[root]["frames"]]
    # done_flags_in_graph: bands store their pixels straight into the root's frame (set_target(frames[i])) and the
    # frame's graph ends with the "in place" flag; False: the caller moves bands with band_push (copy engine) instead
    rasterizer.set_band_peers(world, rank, root if done_flags_in_graph else -1, boxes)
    return {"mailboxes": boxes, "frames": frames}


def finish_band_frame(rasterizer, dist=None):
    """Wait for this rank's band of the frame; True if the frame has to be rendered again on EVERY band (some band
    outgrew its buffers / changed its sort / had too many residue paths — SLPR_ERR_RETRY), agreed over all ranks."""
    from . import SlprRetry
    retry = 0
    try:
        rasterizer.synchronize()
    except SlprRetry:
        retry = 1
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([retry], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        retry = int(t.item())
    return bool(retry)


def sparse_band_table(entries_per_band, band):
    """numpy restatement of k_band_merge for tests: entries_per_band[r] = array [[path, a, inv, z], ...] (the non-zero
    per-path sums band r publishes); returns (path, cum, n, z) of `band`: sorted unique paths of the OTHER bands'
    entries, cum = inclusive prefix of D = sum(a + inv + z), n = sum over bands below of a, z = sum of (a + inv)."""
    rows = []
    for r, e in enumerate(entries_per_band):
        if r == band:
            continue
        for p, a, inv, z in np.asarray(e, dtype=np.int64).reshape(-1, 4):
            rows.append((p, a + inv + z, a if r < band else 0, a + inv))
    if not rows:
        return tuple(np.zeros(0, np.int64) for _ in range(4))
    rows = np.array(rows, dtype=np.int64)
    paths = np.unique(rows[:, 0])
    D = np.array([rows[rows[:, 0] == p, 1].sum() for p in paths])
    N = np.array([rows[rows[:, 0] == p, 2].sum() for p in paths])
    Z = np.array([rows[rows[:, 0] == p, 3].sum() for p in paths])
    return paths, np.cumsum(D), N, Z


def sparse_band_lookup(table, path, row0):
    """band_table_lookup (csrc/bands.cuh): the winding correction of a fragment of `path`."""
    paths, cum, n, z = table
    i = int(np.searchsorted(paths, path, side="left"))
    c = int(cum[i - 1]) if i > 0 else 0
    if i < len(paths) and paths[i] == path:
        c += int(z[i] if row0 else n[i])
    return c
