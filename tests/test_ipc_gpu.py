"""Exact row bands between PROCESSES on the GPU box: one process per band (as in production: one process per GPU),
mailboxes and the root's frame buffer shared through CUDA IPC (slpr_ipc_export / slpr_ipc_import), handles
exchanged with torch.distributed (gloo here, so that the test also runs with both processes on ONE GPU). Pixels of
the assembled frame — every band stores straight into the root's buffer — against the oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
        import vkscanlinepr_b200 as V
        from vkscanlinepr_b200 import parallel as PAR, scene as S
        from oracle import oracle_py as O
        dev = rank % torch.cuda.device_count()
        W, H = 640, 480
        sc = S.synth_scene(3000, W, H, 4.0, 60.0, seed=0x5CA71E02)  # has winding residues: independent bands differ
        rows = S.identity_rows()
        r = V.ScanlineRasterizer(dev, 0).initialize(None, W, H)
        r.loadVG(sc)
        r.setMVP(rows)
        bands = PAR.band_rows(H, world)
        r.set_band(*bands[rank])
        peers = PAR.connect_band_peers(r, dist, rank, world, root=0, frame_bytes=W * H * 4, n_frames=2)
        seq, ok = 1, True
        for frame_no in range(3):
            tries = 0
            while True:
                r.set_target(peers["frames"][seq & 1], W * 4)
                r.render_band(seq)
                if rank == 0:
                    r.band_wait_gather(seq)
                retry = PAR.finish_band_frame(r, dist)
                seq += 1
                tries += 1
                if not retry:
                    break
                assert tries < 4
            dist.barrier()
            if rank == 0:
                ref = O.render(sc, rows, W, H, keep={"rgba"})["rgba"]
                whole = V.ScanlineRasterizer(dev, 0).initialize(None, W, H)
                whole.loadVG(sc); whole.setMVP(rows); whole.render()
                assert np.array_equal(whole.readback(), ref)
                fb, _ = whole.framebuffer()
                ok = ok and whole.diff_u32(peers["frames"][(seq - 1) & 1], fb, W * H) == 0
                whole.close()
            dist.barrier()
        r.close()
        q.put((rank, ok))
    except Exception as e:  # report instead of leaving the other rank waiting
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_bands_between_processes_over_cuda_ipc(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
