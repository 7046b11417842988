"""pcpx_extract_bands (the device-side strip cut of the multi-GPU halo exchange)
against numpy boolean masking: same rows, same (input) order, exact counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_extract_bands_matches_masking(pcpx):
    import torch

    rng = np.random.default_rng(17)
    for n in (0, 1, 4095, 4096, 4097, 1_000_003):
        xyz = rng.uniform(0, 10, (n, 3)).astype(np.float32)
        d = torch.from_numpy(xyz).cuda()
        cap = max(1, n // 5)
        lo = torch.empty((cap, 3), dtype=torch.float32, device="cuda")
        hi = torch.empty((cap, 3), dtype=torch.float32, device="cuda")
        for axis, below, above in ((0, 1.0, 9.0), (2, 0.5, 9.9), (1, float("-inf"), 9.5),
                                   (0, 1.5, float("inf"))):
            counts = torch.zeros(2, dtype=torch.int64, device="cuda")
            pcpx.extract_bands(d, axis, below, above, lo, hi, counts)
            c = xyz[:, axis]
            want_lo, want_hi = xyz[c < below], xyz[c > above]
            got = counts.tolist()
            assert got == [len(want_lo), len(want_hi)]
            assert np.array_equal(lo[: got[0]].cpu().numpy(), want_lo)
            assert np.array_equal(hi[: got[1]].cpu().numpy(), want_hi)
        # host counters work too
        hc = np.zeros(2, np.uint64)
        pcpx.extract_bands(d, 0, 1.0, 9.0, lo, hi, hc)
        assert hc.tolist() == [int((xyz[:, 0] < 1.0).sum()), int((xyz[:, 0] > 9.0).sum())]


def test_extract_bands_counts_past_capacity(pcpx):
    import torch

    xyz = np.random.default_rng(3).uniform(0, 1, (100_000, 3)).astype(np.float32)
    d = torch.from_numpy(xyz).cuda()
    lo = torch.full((10, 3), -1.0, device="cuda")
    hi = torch.full((10, 3), -1.0, device="cuda")
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    pcpx.extract_bands(d, 0, 0.5, 0.5, lo, hi, counts)
    c = xyz[:, 0]
    assert counts.tolist() == [int((c < 0.5).sum()), int((c > 0.5).sum())]  # counted, not written
    assert np.array_equal(lo.cpu().numpy(), xyz[c < 0.5][:10])
    assert np.array_equal(hi.cpu().numpy(), xyz[c > 0.5][:10])
    with pytest.raises(pcpx.PcpxError):
        pcpx.extract_bands(d, 3, 0.5, 0.5, lo, hi, counts)
