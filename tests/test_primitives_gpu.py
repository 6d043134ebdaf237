"""Property tests of the two roofline-graded device primitives at sizes the oracle cannot reach:
the decoupled-look-back scan and the onesweep radix sort (checked against torch on the device)."""
import numpy as np
import pytest

import vkscanlinepr_b200 as V

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    r = V.ScanlineRasterizer(0, 0).initialize(None, 64, 64)
    yield r
    r.close()


# (from 2^22 elements on the TMA-pipelined kernel k_scan_tma takes over: whole tiles by bulk copy, a ragged last tile)
@pytest.mark.parametrize("n", [0, 1, 3, 4095, 4096, 4097, 8191, 8192, 8193, 100_003, (1 << 22) - 1, 1 << 22, (1 << 22) + 8192 * 7,
                               (1 << 22) + 8192 * 300 + 5, 20_000_000, 35_000_003])
def test_scan_matches_cumsum(ctx, n):
    import torch
    g = torch.Generator(device="cuda").manual_seed(n + 1)
    a = torch.randint(-3, 4, (max(n, 1),), dtype=torch.int32, device="cuda", generator=g)[:n]
    out = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.scan_i32(a.data_ptr() if n else out.data_ptr(), out.data_ptr(), n)
    ctx.synchronize()
    exp = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    exp[1:] = torch.cumsum(a.to(torch.int64), 0)
    assert torch.equal(out.to(torch.int64), exp)


def test_scan_unaligned_pointers_and_in_place(ctx):
    """Pointers that are not 16-byte aligned fall back to the ticketed kernel; in-place scanning is allowed."""
    import torch
    n = (1 << 22) + 1000
    base = torch.randint(0, 5, (n + 8,), dtype=torch.int32, device="cuda")
    out = torch.empty(n + 9, dtype=torch.int32, device="cuda")
    for off_in, off_out in ((1, 0), (0, 3), (2, 2)):
        a = base[off_in:off_in + n]
        o = out[off_out:off_out + n + 1]
        ctx.scan_i32(a.data_ptr(), o.data_ptr(), n)
        ctx.synchronize()
        exp = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        exp[1:] = torch.cumsum(a.to(torch.int64), 0)
        assert torch.equal(o.to(torch.int64), exp)
    buf = torch.zeros(n + 4, dtype=torch.int32, device="cuda")
    buf[:n] = base[:n]
    exp = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    exp[1:] = torch.cumsum(base[:n].to(torch.int64), 0)
    ctx.scan_i32(buf.data_ptr(), buf.data_ptr(), n)
    ctx.synchronize()
    assert torch.equal(buf[:n + 1].to(torch.int64), exp)


def test_scan_wraps_like_int32(ctx):
    import torch
    a = torch.full((10000,), 2**30, dtype=torch.int32, device="cuda")
    out = torch.empty(10001, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.scan_i32(a.data_ptr(), out.data_ptr(), 10000)
    ctx.synchronize()
    exp = (torch.arange(10001, dtype=torch.int64, device="cuda") * 2**30)
    exp = ((exp + 2**31) % 2**32 - 2**31).to(torch.int32)
    assert torch.equal(out, exp)


@pytest.mark.parametrize("n,bits", [(1, 8), (6143, 16), (6144, 40), (6145, 40), (1_000_003, 40), (20_000_000, 40),
                                   (3_000_000, 64), (2_000_000, 9)])
def test_sort_is_stable_and_sorted(ctx, n, bits):
    import torch
    g = torch.Generator(device="cuda").manual_seed(n)
    hi = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    lo = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    keys = (hi << 33) ^ (lo << 2) ^ (hi >> 7)
    if bits < 64:
        keys = keys & ((1 << bits) - 1)
    if n > 1000:
        keys[: n // 3] = keys[0]            # long runs of equal keys: stability must keep index order
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    k2, v2 = torch.empty_like(keys), torch.empty_like(vals)
    kin, vin = keys.clone(), vals.clone()
    torch.cuda.synchronize()
    which = ctx.sort_pairs(kin.data_ptr(), vin.data_ptr(), k2.data_ptr(), v2.data_ptr(), n, bits)
    ctx.synchronize()
    ks, vs = (k2, v2) if which else (kin, vin)
    # reference: stable sort on the unsigned key (keys < 2^63 here unless bits == 64)
    order = torch.sort(keys if bits < 64 else keys ^ (1 << 63), stable=True).indices if bits == 64 and False else None
    if bits == 64:
        # compare as unsigned: map to signed order by flipping the top bit
        skeys = keys ^ torch.tensor(-2**63, dtype=torch.int64, device="cuda")
        order = torch.sort(skeys, stable=True).indices
    else:
        order = torch.sort(keys, stable=True).indices
    assert torch.equal(vs.to(torch.int64), order)
    assert torch.equal(ks, keys[order])
