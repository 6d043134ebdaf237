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
        ok = ok and _exact_band_windings(rank, world, dist, torch)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _exact_band_windings(rank, world, dist, torch):
    """The exchange step of exact row bands (csrc/bands.cuh) on oracle data: every rank keeps the fragments
    of its rows, all-gathers its per-path winding sums, and must recover the FULL frame's winding number at
    each of its fragments as local scan + correction — on a scene with a winding residue."""
    from oracle import oracle_py as O
    W = H = 192
    sc = S.synth_scene(160, W, H, 6.0, 30.0, seed=0x5CA71E01)
    ref = O.render(sc, S.identity_rows(), W, H, do_fill=False)
    assert ref["wn"][-1] != 0, "the test scene must have a winding residue"
    sk, si = ref["skey"], ref["sidx"]
    path_s, d_s, wn = ref["path"][si].astype(np.int64), ref["wind"][si].astype(np.int64), ref["wn"][:-1]
    invalid = sk.view(np.uint32) == 0xFFFEFFFE
    y = (sk.view(np.uint32) >> 16).astype(np.int64) - 0x7FFF
    cls = np.where(invalid, 1, np.where(y == 0, 2, 0))  # normal rows | outside the frame | row 0
    bands = PAR.band_rows(H, world)
    owner = np.zeros(len(sk), dtype=np.int64)           # fragments outside the frame: band 0 (any single owner works)
    for r, (y0, y1) in enumerate(bands):
        owner[~invalid & (y >= y0) & (y < y1)] = r
    mine = owner == rank
    P = sc.n_paths
    sums = np.zeros((3, P), dtype=np.int32)
    np.add.at(sums, (cls[mine], path_s[mine]), d_s[mine].astype(np.int32))
    gathered = torch.zeros((world, 3 * P), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(sums.reshape(-1)))
    corr_n, corr_z = PAR.band_corrections(gathered.numpy().reshape(world, 3, P), rank)
    local = np.cumsum(np.where(mine, d_s, 0)) - np.where(mine, d_s, 0)  # exclusive scan of this band's own deltas
    check = mine & (cls != 1)
    corr = np.where(cls == 2, corr_z[path_s], corr_n[path_s])
    ok = bool(np.array_equal((local + corr)[check].astype(np.int32), wn[check]))
    # the sparse exchange (round 2): only the paths with a non-zero sum travel, and the corrections come from a
    # sorted break-point table (k_band_merge / band_table_lookup) — same numbers as the dense tables
    nz = np.nonzero(sums.any(axis=0))[0]
    mine_entries = np.stack([nz, sums[0, nz], sums[1, nz], sums[2, nz]], axis=1) if len(nz) else np.zeros((0, 4), np.int64)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine_entries)
    table = PAR.sparse_band_table(everyone, rank)
    assert len(nz) < P // 4 and sum(len(e) for e in everyone) > 0, "the exchange must be sparse, and not empty, on this scene"
    sparse = np.array([PAR.sparse_band_lookup(table, p, c == 2) for p, c in zip(path_s[check], cls[check])])
    return ok and bool(np.array_equal(sparse, corr[check]))


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


def test_sparse_band_table_equals_dense_corrections():
    """k_band_merge's break-point table against the dense corrN / corrZ tables, random sums, 5 bands."""
    rng = np.random.default_rng(11)
    G, P = 5, 300
    dense = np.zeros((G, 3, P), np.int64)
    for r in range(G):
        hit = rng.choice(P, 12, replace=False)
        dense[r][:, hit] = rng.integers(-3, 4, (3, 12))
    entries = []
    for r in range(G):
        nz = np.nonzero(dense[r].any(axis=0))[0]
        entries.append(np.stack([nz, dense[r][0, nz], dense[r][1, nz], dense[r][2, nz]], axis=1))
    for band in range(G):
        cn, cz = PAR.band_corrections(dense, band)
        t = PAR.sparse_band_table(entries, band)
        assert np.all(np.diff(t[0]) > 0)
        for p in range(P):
            assert PAR.sparse_band_lookup(t, p, False) == cn[p] and PAR.sparse_band_lookup(t, p, True) == cz[p]
