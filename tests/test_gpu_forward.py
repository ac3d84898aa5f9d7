"""predict.FM parity: CUDA forward kernel vs the oracle (Model::predict_batch / predict_prob,
reference src/core/Model.h:106-180).  Tolerance 1e-5 relative with max(1,|ref|) denominator (fp32),
1e-12 for the fp64 instantiation."""
import numpy as np
import pytest

from oracle import oracle as O
from fmwr_b200 import _lib as L
from fmwr_b200 import synth
from tests.util import relerr, csr_to_r_lists

pytestmark = pytest.mark.gpu


def run_forward(ctx, prec, n, p, rowptr, col, val, w0, w, v, k, link=L.LINK_NONE, lo=0.0, hi=0.0, k0=1, k1=1):
    d = L.Data.from_csr32(ctx, n, p, rowptr, col, val)
    cfg = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=k0, keep_w1=k1, k=k)
    m = L.Model(ctx, cfg, p, prec)
    m.set(w0, w, v)
    L.predict_dev(ctx, m, d, link, lo, hi)
    out = L.predict_fetch(ctx, d)
    m.close(); d.close()
    return out


def test_forward_kat(gpu_ctx, port):
    # known-answer pattern from the reference's scratch test (src/test/model.cpp:12-34)
    X = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [1, 0, 0, 0], [0, 0, 1, 0], [1, 0, 0, 1], [2, 0, 0, 0],
                  [0, 0, 0, 1]], np.float32)
    rowptr = [0]; col = []; val = []
    for r in X:
        nz = np.nonzero(r)[0]; col += list(nz); val += list(r[nz]); rowptr.append(len(col))
    w = np.full(4, 0.1); v = np.full((4, 3), 0.2)
    want = np.array([.1, .1, .1, .1, .1, .32, .2, .1])
    for prec, tol in ((L.F32, 1e-6), (L.F64, 1e-14)):
        got = run_forward(gpu_ctx, prec, 8, 4, rowptr, col, val, 0.0, w, v, 3)
        assert relerr(got, want) < tol
    got = run_forward(gpu_ctx, L.F64, 8, 4, rowptr, col, val, 0.0, w, v, 3, link=L.LINK_LOGISTIC)
    assert abs(got[0] - 0.5249791875) < 1e-9 and abs(got[5] - 0.5793242521) < 1e-9
    got = run_forward(gpu_ctx, L.F64, 8, 4, rowptr, col, val, 0.0, w, v, 3, link=L.LINK_PROBIT_TABLE)
    assert abs(got[0] - 0.5398278371) < 1e-9 and abs(got[5] - 0.6255158326) < 1e-9


@pytest.mark.parametrize("k", [0, 1, 2, 3, 8, 10, 29, 30, 32, 33, 64, 100, 128, 200])
@pytest.mark.parametrize("prec", [L.F32, L.F64])
def test_forward_random_ragged(gpu_ctx, port, k, prec):
    if prec == L.F64 and k > 128 * 2:
        pytest.skip("beyond fp64 layout")
    rng = np.random.default_rng(k + 7)
    n, p = 700, 300
    rowptr, col, val = synth.random_csr(n, p, 45, seed=k, empty_rows=True)
    w = rng.normal(0, 0.3, p); v = rng.normal(0, 0.2, (p, k)); w0 = 0.25
    cfg = O.make_cfg(k=k)
    want = port.predict(cfg, n, p, rowptr, col, val, w0, w, v, 0)
    got = run_forward(gpu_ctx, prec, n, p, rowptr, col, val, w0, w, v, k)
    assert relerr(got, want) < (1e-5 if prec == L.F32 else 1e-12)


@pytest.mark.parametrize("k0,k1", [(0, 0), (0, 1), (1, 0)])
def test_forward_keep_flags(gpu_ctx, port, k0, k1):
    rng = np.random.default_rng(3)
    n, p, k = 200, 80, 8
    rowptr, col, val = synth.random_csr(n, p, 10, seed=9)
    w = rng.normal(0, 0.3, p); v = rng.normal(0, 0.2, (p, k)); w0 = -0.4
    cfg = O.make_cfg(k=k, k0=k0, k1=k1)
    want = port.predict(cfg, n, p, rowptr, col, val, w0, w, v, 0)
    got = run_forward(gpu_ctx, L.F64, n, p, rowptr, col, val, w0, w, v, k, k0=k0, k1=k1)
    assert relerr(got, want) < 1e-12


@pytest.mark.parametrize("solver,link", [(O.SGD, L.LINK_LOGISTIC), (O.ALS, L.LINK_PROBIT_TABLE)])
def test_forward_links(gpu_ctx, port, solver, link):
    rng = np.random.default_rng(11)
    n, p, k = 3000, 500, 8
    rowptr, col, val = synth.random_csr(n, p, 12, seed=2)
    w = rng.normal(0, 1.0, p); v = rng.normal(0, 0.5, (p, k)); w0 = 0.1      # wide scores: exercises the table tails
    cfg = O.make_cfg(k=k, solver=solver)
    want = port.predict(cfg, n, p, rowptr, col, val, w0, w, v, 1)
    got = run_forward(gpu_ctx, L.F64, n, p, rowptr, col, val, w0, w, v, k, link=link)
    assert relerr(got, want) < 1e-9
    got32 = run_forward(gpu_ctx, L.F32, n, p, rowptr, col, val, w0, w, v, k, link=link)
    assert relerr(got32, want) < 1e-5


def test_forward_clamp_and_r_lists(gpu_ctx, port):
    rng = np.random.default_rng(5)
    n, p, k = 500, 100, 4
    rowptr, col, val = synth.random_csr(n, p, 9, seed=4)
    w = rng.normal(0, 0.5, p); v = rng.normal(0, 0.3, (p, k)); w0 = 0.0
    cfg = O.make_cfg(k=k)
    want = np.clip(port.predict(cfg, n, p, rowptr, col, val, w0, w, v, 0), -0.3, 0.4)
    # through the one-shot host entry point with the R list layout (FMPredict body)
    rs, ci, vv = csr_to_r_lists(rowptr, col, val)
    out = np.zeros(n)
    import ctypes as C
    mc = L.ModelCfg(task=L.REGRESSION, keep_w0=1, keep_w1=1, k=k)
    L.check(L.lib().fmwr_predict(C.byref(mc), L.F64, C.c_int64(n), C.c_int64(p), C.c_int64(ci.size), L.ptr(rs), L.ptr(ci),
                                 L.ptr(vv), C.c_double(w0), L.ptr(np.ascontiguousarray(w)), L.ptr(np.ascontiguousarray(v)),
                                 L.LINK_CLAMP, C.c_double(-0.3), C.c_double(0.4), L.ptr(out)))
    assert relerr(out, want) < 1e-12


def test_forward_shape_error(gpu_ctx):
    rowptr, col, val = synth.random_csr(10, 20, 3, seed=1)
    d = L.Data.from_csr32(gpu_ctx, 10, 20, rowptr, col, val)
    m = L.Model(gpu_ctx, L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=2), 21, L.F32)
    with pytest.raises(L.FmwrError) as e:
        L.predict_dev(gpu_ctx, m, d)
    assert "features is not correct" in str(e.value)      # same message as Model::predict_batch (Model.h:113)
    with pytest.raises(L.FmwrError):
        L.Data.from_csr32(gpu_ctx, 2, 3, [0, 1, 2], [0, 5], [1.0, 1.0])   # column out of range


def test_forward_criteo_shape_full_width(gpu_ctx, port):
    # Criteo-shaped rows (39 nnz, k=32): the headline layout, small enough for the oracle
    ds = synth.make_dataset("criteo", 20000, p=39 * 2000)
    rng = np.random.default_rng(0)
    k = 32
    w = rng.normal(0, 0.1, ds["p"]); v = rng.normal(0, 0.05, (ds["p"], k))
    cfg = O.make_cfg(k=k, nthreads=8)
    want = port.predict(cfg, ds["n"], ds["p"], ds["rowptr"], ds["col"], ds["val"], 0.2, w, v, 0)
    got = run_forward(gpu_ctx, L.F32, ds["n"], ds["p"], ds["rowptr"], ds["col"], ds["val"], 0.2, w, v, k)
    assert relerr(got, want) < 1e-5


def test_synth_device_matches_host(gpu_ctx):
    fs = [1000, 777, 5000]
    d = L.Data.synth(gpu_ctx, 5000, fs, skew=[0, 1, 0], value_mode=1, label_mode=1, seed=77)
    rowptr, col, val, y = d.get_csr()
    r2, c2, v2, p = synth.fields_csr(5000, fs, [0, 1, 0], 1, 77)
    assert (rowptr == r2).all() and (col == c2).all() and (val == v2).all()
    assert set(np.unique(y)) <= {-1.0, 1.0} and 0.2 < (y > 0).mean() < 0.8
    d.close()
