"""CPU: the fp32 summation order that libnwx's sample_pdf kernel implements (cascade_sum_emul in
csrc/sample_pdf.cu) IS the order of torch.sum on a contiguous fp32 row -- for every row length the
kernel accepts, not just the reference's 62.  If a future torch changes ATen's cascade, this test
says so before the GPU index-parity tests do."""
import numpy as np
import torch

F = np.float32


def cascade_sum_model(x: np.ndarray) -> np.ndarray:
    """numpy model of cascade_sum_emul: 8-lane vectors, 4-way ILP rows, scalar tail first, then the
    8 lane partials in order (ATen SumKernel.cpp vectorized_inner_sum / row_sum / multi_row_sum)."""
    n = x.shape[1]
    V, ilp = n >> 3, (n >> 3) >> 2
    p = [np.zeros((x.shape[0], 8), F) for _ in range(4)]
    for i in range(ilp):
        for k in range(4):
            p[k] = (p[k] + x[:, (4 * i + k) * 8:(4 * i + k) * 8 + 8]).astype(F)
    for i in range(4 * ilp, V):
        p[0] = (p[0] + x[:, i * 8:i * 8 + 8]).astype(F)
    p0 = (((p[0] + p[1]).astype(F) + p[2]).astype(F) + p[3]).astype(F)
    acc = np.zeros(x.shape[0], F)
    for k in range(V * 8, n):
        acc = (acc + x[:, k]).astype(F)
    for lane in range(8):
        acc = (acc + p0[:, lane]).astype(F)
    return acc


def test_cascade_order_matches_torch_sum_for_all_lengths():
    rng = np.random.RandomState(1)
    for n in list(range(8, 127, 3)) + [62, 126]:
        x = ((rng.rand(4000, n) ** 5).astype(F) + F(1e-5)).astype(F)
        ref = torch.sum(torch.from_numpy(x), -1).numpy()
        assert np.array_equal(cascade_sum_model(x), ref), n


def test_cdf_in_double_is_order_independent():
    """pdf entries are >= 1e-5/1.0007 and sum to ~1: every partial sum is exact in fp64, so the kernel's
    warp scan equals torch.cumsum's sequential double accumulation bit for bit."""
    rng = np.random.RandomState(2)
    w = (rng.rand(5000, 62) ** 6).astype(F) + F(1e-5)
    pdf = (w / cascade_sum_model(w)[:, None]).astype(F)
    seq = np.cumsum(pdf.astype(np.float64), -1)
    tree = pdf.astype(np.float64)
    step = 1
    while step < 62:                                   # Hillis-Steele scan, the warp-scan association
        shifted = np.zeros_like(tree); shifted[:, step:] = tree[:, :-step]
        tree = tree + shifted
        step *= 2
    assert np.array_equal(seq, tree)
    ref = torch.cumsum(torch.from_numpy(pdf), -1).numpy()
    assert np.array_equal(seq.astype(F), ref)
