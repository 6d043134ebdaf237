"""Multi-GPU host logic (new work; the reference is single-device — SURVEY §8e).

The path shards two ways and has NO exchange step inside the algorithm:
  * row bands  — rank r renders scanline rows [y0, y1) of one frame (slpr_set_band); the only
    communication is the gather of the finished RGBA8 bands to rank 0 (NCCL send/recv over NVLink,
    or direct stores into a peer-mapped frame through slpr_set_target);
  * frame batches — frame f goes to rank f mod world; nothing is exchanged.
One process per GPU, torch.distributed for the plumbing. These helpers are backend-agnostic so the
host logic is covered on CPU with gloo (tests/test_parallel_cpu.py).
"""
import numpy as np


def band_rows(height, world):
    """Even-aligned scanline-row bands [(y0, y1)] covering [0, height): fragments are 2 px tall, so a band
    edge must be even (slpr_set_band); the last band takes the remainder."""
    if world < 1 or height < 1:
        raise ValueError("world and height must be positive")
    per = (height // world) & ~1
    if per == 0:
        if world > 1:
            raise ValueError(f"height {height} is too small for {world} even-aligned bands")
        return [(0, height)]
    return [(r * per, height if r == world - 1 else (r + 1) * per) for r in range(world)]


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
