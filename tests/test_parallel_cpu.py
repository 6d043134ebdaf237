"""Host logic of the multi-GPU paths on CPU: world_size-2 gloo (the N>1 plumbing of bench.py --mode
bands, and the frame-parallel partition). The per-rank band images come from the oracle here; on a
GPU box the same functions move device tensors over NCCL."""
import os
import socket

import numpy as np
import pytest

import util
from vkscanlinepr_b200 import parallel as PAR
from vkscanlinepr_b200 import scene as S


def test_band_rows_are_even_and_cover():
    for H in (2, 16, 1080, 2160, 16384, 1081):
        for G in (1, 2, 4, 8):
            if H // G < 2 and G > 1:
                with pytest.raises(ValueError):
                    PAR.band_rows(H, G)
                continue
            b = PAR.band_rows(H, G)
            assert len(b) == G and b[0][0] == 0 and b[-1][1] == H
            assert all(b[i][1] == b[i + 1][0] for i in range(G - 1))
            assert all(y0 % 2 == 0 and (y1 % 2 == 0 or y1 == H) and y0 < y1 for y0, y1 in b)


def test_image_rows_flip():
    assert PAR.image_rows(100, 0, 10) == slice(90, 100)
    assert PAR.image_rows(100, 90, 100) == slice(0, 10)


def test_frames_round_robin():
    parts = [PAR.frames_of_rank(256, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(256))
    assert all(f % 8 == r for r, p in enumerate(parts) for f in p)
    assert np.array_equal(PAR.scatter_frames(10, 4), [0, 1, 2, 3, 0, 1, 2, 3, 0, 1])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_py as O
        sc = util.tiny_scene()
        W, H = 96, 80
        full = O.render(sc, S.identity_rows(), W, H)["rgba"]
        bands = PAR.band_rows(H, world)
        frame = torch.zeros((H, W, 4), dtype=torch.uint8)
        rows = PAR.image_rows(H, *bands[rank])
        frame[rows] = torch.from_numpy(full[rows])          # this rank's band only
        PAR.gather_bands(frame, bands, rank, world, dist, dst=0)
        ok = True
        if rank == 0:
            ok = bool(np.array_equal(frame.numpy(), full))
        # frame-parallel: every rank reports which frames it rendered; the union is the batch
        mine = torch.zeros(16, dtype=torch.int32)
        mine[PAR.frames_of_rank(16, rank, world)] = 1
        dist.all_reduce(mine)
        ok = ok and bool((mine == 1).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_band_gather_and_frame_partition():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
