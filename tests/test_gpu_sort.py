"""The engine's own stable radix sort (csrc/sort.cu) against numpy's stable argsort -- bit-exact.

It replaces the library sort behind CSR -> CSC (SMatrix::transpose, reference src/util/Smatrix.h:155-185: rows must
stay ascending inside a column, i.e. equal keys keep their input order), the per-batch CSC of the minibatch trainers
and the |score| order of AUC (src/core/Evaluation.h:56-78)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check(gpu_ctx, keys, bits):
    out, perm = gpu_ctx.sort_pairs(keys, bits)
    mask = (1 << bits) - 1 if bits < 64 else 0xFFFFFFFFFFFFFFFF
    low = keys & keys.dtype.type(mask)
    want = np.argsort(low, kind="stable")
    assert (perm == want.astype(np.uint32)).all()
    assert (out == keys[want]).all()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 4095, 4096, 4097, 50_000, 1_000_003])
@pytest.mark.parametrize("bits", [1, 7, 8, 9, 20, 28, 32])
def test_sort_u32_sizes_and_bits(gpu_ctx, n, bits):
    rng = np.random.default_rng(n * 131 + bits)
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    check(gpu_ctx, keys, bits)


@pytest.mark.parametrize("n,bits", [(1, 64), (5000, 64), (300_000, 33), (1_000_003, 40), (200_000, 64)])
def test_sort_u64(gpu_ctx, n, bits):
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, np.iinfo(np.uint64).max, n, dtype=np.uint64, endpoint=True)
    check(gpu_ctx, keys, bits)


def test_sort_stability_with_few_distinct_keys(gpu_ctx):
    # every digit bin collides massively: ranks inside a tile, across warps and across tiles must all keep input order
    rng = np.random.default_rng(7)
    for distinct in (1, 2, 3, 17, 300):
        keys = rng.integers(0, distinct, 700_001).astype(np.uint32)
        check(gpu_ctx, keys, 12)
    keys = np.zeros(100_000, np.uint32)
    check(gpu_ctx, keys, 32)


def test_sort_presorted_and_reversed(gpu_ctx):
    keys = np.arange(500_000, dtype=np.uint32)
    check(gpu_ctx, keys, 19)
    check(gpu_ctx, keys[::-1].copy(), 19)
    # (batch, column) keys as minibatch_build forms them: the batch bits are already in order
    n, F, S = 40_000, 13, 900
    rng = np.random.default_rng(3)
    col = (np.arange(F)[None, :] * S + rng.integers(0, S, (n, F))).astype(np.uint64)
    key = ((np.arange(n)[:, None] // 4096).astype(np.uint64) << np.uint64(14)) | col
    check(gpu_ctx, key.reshape(-1).astype(np.uint32), 18)


def test_sort_empty(gpu_ctx):
    out, perm = gpu_ctx.sort_pairs(np.zeros(0, np.uint32), 8)
    assert out.size == 0 and perm.size == 0
