"""BASELINE.json's FULL sizes, checked through size-independent properties (the oracle cannot run 10M rows in seconds):

  predict.FM     rows are independent: any window of the 10M-row matrix, generated separately, gets the same predictions,
                 and a sample of rows equals the oracle
  FTRL minibatch an epoch is deterministic bit for bit, every parameter stays finite, the log-likelihood rises
  ALS            exact coordinate minimisation of the squared loss: the train RMSE never increases from sweep to sweep
  CSR -> CSC     transpose of the transpose is the matrix (a checksum of checksums over 60M entries)
"""
import numpy as np
import pytest

from oracle import oracle as O
from fmwr_b200 import _lib as L
from tests.util import relerr

pytestmark = pytest.mark.gpu

F, FIELD, K = 39, 25641, 32                      # configs[1]: 10M rows x 39 nnz, 999 999 features, k = 32
N_FULL = 10_000_000


def test_predict_full_size_row_independence_and_oracle_sample(gpu_ctx, port):
    ctx = gpu_ctx
    p = F * FIELD
    d = L.Data.synth(ctx, N_FULL, [FIELD] * F, None, 1, 1, 0.1, 20240601)
    m = L.Model(ctx, L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=K), p, L.F32)
    m.init_random(0.0, 0.05, 7)
    L.predict_dev(ctx, m, d, L.LINK_LOGISTIC)
    full = L.predict_fetch(ctx, d)
    assert np.isfinite(full).all() and 0.0 < full.min() and full.max() < 1.0
    # a window generated on its own (row-sharded predict) gives the same rows
    r0, nr = 7_654_321, 200_000
    win = L.Data.synth_rows(ctx, r0, nr, [FIELD] * F, None, 1, 0, 0.1, 20240601)
    L.predict_dev(ctx, m, win, L.LINK_LOGISTIC)
    part = L.predict_fetch(ctx, win)
    assert np.array_equal(part, full[r0:r0 + nr])
    # 2000 of those rows against the oracle (Model::predict_prob, src/core/Model.h:163-180)
    rowptr, col, val, _ = win.get_csr(labels=False)
    w0, w, v = m.get()
    ns = 2000
    e1 = int(rowptr[ns])
    want = port.predict(O.make_cfg(k=K), ns, p, rowptr[:ns + 1], col[:e1], val[:e1], w0, w, v, 1)
    assert relerr(part[:ns], want) < 1e-5
    win.close(); m.close(); d.close()


def test_ftrl_epoch_full_size_is_deterministic_and_learns(gpu_ctx):
    ctx = gpu_ctx
    p = F * FIELD
    d = L.Data.synth(ctx, N_FULL, [FIELD] * F, None, 0, 1, 0.1, 20240601)
    mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=K, l1_w1=1e-3, l2_w1=1e-3, l2_v=1e-3)
    sc = L.SolverCfg(solver=L.FTRL, max_iter=N_FULL - 1, random_step=1, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0,
                     min_target=-1.0, max_target=1.0, mode=L.MODE_MINIBATCH, batch_size=65536, precision=L.F32,
                     compat=L.COMPAT_REFERENCE, step_size=-1)
    runs = []
    for _ in range(2):
        m = L.Model(ctx, mc, p, L.F32)
        m.init_random(0.0, 0.01, 20240603)
        L.predict_dev(ctx, m, d, L.LINK_LOGISTIC)
        ll0 = L.evaluate_dev(ctx, d, L.CLASSIFICATION, L.LL)
        L.train_dev(ctx, m, d, sc)
        L.predict_dev(ctx, m, d, L.LINK_LOGISTIC)
        ll1 = L.evaluate_dev(ctx, d, L.CLASSIFICATION, L.LL)
        runs.append((m.get(), ll0, ll1))
        m.close()
    (a, ll0, ll1), (b, _, _) = runs
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])      # no atomics: bit-identical epochs
    assert np.isfinite(a[1]).all() and np.isfinite(a[2]).all()
    # the reference's LL is the log-likelihood SUM (src/core/Evaluation.h:80-89): n ln(1/2) before, much higher after one epoch
    assert ll0 < -0.69 * N_FULL and ll1 > 0.8 * ll0, (ll0, ll1)
    d.close()


def test_als_full_size_rmse_never_increases(gpu_ctx):
    ctx = gpu_ctx
    fields = [138493, 26744, 2048]                 # configs[2]: 20M ratings, user / item (skewed) / context one-hot
    n, p, k = 20_000_000, sum(fields), 32
    d = L.Data.synth(ctx, n, fields, [0, 1, 0], 0, 3, 0.3, 20240601)
    m = L.Model(ctx, L.ModelCfg(task=L.REGRESSION, keep_w0=1, keep_w1=1, k=k), p, L.F32)
    m.init_random(0.0, 0.01, 3)
    sweeps = 4
    sc = L.SolverCfg(solver=L.ALS, max_iter=sweeps, random_step=1, min_target=0.5, max_target=5.0, mode=L.MODE_EXACT, precision=L.F32,
                     compat=L.COMPAT_REFERENCE, enable_v=1, step_size=1, metric=L.RMSE, convergence=0.0, seed=5)
    tr = L.TraceBuf(16)
    L.train_dev(ctx, m, d, sc, tr)
    rmse = list(tr.result()["eval_train"])
    L.predict_dev(ctx, m, d, L.LINK_CLAMP, 0.5, 5.0)
    rmse.append(L.evaluate_dev(ctx, d, L.REGRESSION, L.RMSE))
    assert len(rmse) >= sweeps and all(np.isfinite(rmse))
    assert all(b <= a * (1 + 1e-4) for a, b in zip(rmse, rmse[1:])), rmse
    assert rmse[-1] < 0.9 * rmse[0], rmse
    m.close(); d.close()


def test_transpose_full_size_involution(gpu_ctx):
    ctx = gpu_ctx
    fields = [138493, 26744, 2048]
    n = 20_000_000
    d = L.Data.synth(ctx, n, fields, [0, 1, 0], 1, 3, 0.3, 20240601)
    d.transpose()
    colptr, crow, cval = d.get_csc()
    rowptr, col, val, _ = d.get_csr()
    # checksum of checksums: per-row sums of (column id, value) recomputed from the CSC side
    assert colptr[-1] == col.size and np.all(np.diff(colptr.astype(np.int64)) >= 0)
    ccol = np.repeat(np.arange(colptr.size - 1, dtype=np.int64), np.diff(colptr.astype(np.int64)))
    rs_csc = np.bincount(crow, weights=ccol.astype(np.float64), minlength=n)
    rv_csc = np.bincount(crow, weights=cval.astype(np.float64), minlength=n)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr.astype(np.int64)))
    rs_csr = np.bincount(rows, weights=col.astype(np.float64), minlength=n)
    rv_csr = np.bincount(rows, weights=val.astype(np.float64), minlength=n)
    assert np.array_equal(rs_csc, rs_csr) and np.allclose(rv_csc, rv_csr, rtol=0, atol=1e-9)
    # rows ascend inside every column (the reference's order, src/util/Smatrix.h:155-185)
    inner = np.ones(crow.size, bool); inner[colptr[:-1][np.diff(colptr.astype(np.int64)) > 0]] = False
    assert np.all(np.diff(crow.astype(np.int64))[inner[1:]] > 0)
    d.close()
