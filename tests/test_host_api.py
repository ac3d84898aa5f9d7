"""Host-side mirror of the R layer (fmwr_b200/api.py): argument plumbing, defaults, coercions and error strings of
R/fm_matrix.R, fm_control.R, fm_solver_control.R, fm_track_control.R, control_tools.R, fm_train.R, fm_predict.R.
GPU-marked tests run the same calls end to end and compare with the oracle."""
import warnings

import numpy as np
import pytest

from fmwr_b200 import api as A


def test_fm_matrix_dense_and_sparse_agree():
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    X = rng.random((30, 12)) * (rng.random((30, 12)) < 0.3)
    a = A.fm_matrix(X, np.arange(30))
    b = A.fm_matrix(sp.csc_matrix(X), np.arange(30))
    for k in ("value", "col_idx", "row_size"):
        assert np.array_equal(a["features"][k], b["features"][k])
    assert a["features"]["dim"] == (30, 12) and a["features"]["size"] == int((X != 0).sum())
    assert a["features"]["col_idx"].dtype == np.int32 and a["features"]["value"].dtype == np.float64
    # ascending & unique inside each row (what SMatrix and the CSC transpose rely on)
    ptr = np.concatenate([[0], np.cumsum(a["features"]["row_size"])])
    for i in range(30):
        c = a["features"]["col_idx"][ptr[i]:ptr[i + 1]]
        assert np.all(np.diff(c) > 0)
    with pytest.raises(ValueError):
        A.fm_matrix(X, np.arange(29))


def test_control_defaults_match_r():
    mc = A.model_control()
    assert mc["task"] == "CLASSIFICATION"
    assert mc["hyper.params"] == {"keep.w0": True, "L2.w0": 0.0, "keep.w1": True, "L1.w1": 0.0, "L2.w1": 0.0, "factor.number": 2,
                                  "v.init_mean": 0.0, "v.init_stdev": 0.01, "L1.v": 0.0, "L2.v": 0.0}
    assert dict(A.SGD_solver()) == {"learn_rate": 0.01, "random_step": 1}
    assert dict(A.FTRL_solver()) == {"alpha_w": 0.1, "alpha_v": 0.1, "beta_w": 1.0, "beta_v": 1.0, "random_step": 1}
    assert dict(A.TDAP_solver()) == {"gamma": 1e-4, "alpha_w": 0.1, "alpha_v": 0.1, "random_step": 1}
    assert dict(A.ALS_solver())["w0_mean_0"] == 1.0 and A.MCMC_solver().solver == "MCMC"
    sc = A.solver_control()
    assert sc["max_iter"] == 10000 and sc["solver"].solver == "TDAP"
    tc = A.track_control()
    assert tc["step_size"] == -1 and tc["evaluate.metric"] == "LL" and tc["convergence"] == 1e-4


def test_control_assign_coercions_and_warnings():
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        mc = A.model_control(task="REGRESSION", factor_number=3.7, L2_w1=0.5, bogus=1)
    assert mc["hyper.params"]["factor.number"] == 3 and mc["hyper.params"]["L2.w1"] == 0.5
    msgs = " ".join(str(x.message) for x in w)
    assert "is not integer" in msgs and "unknown" in msgs
    with pytest.raises(TypeError):
        A.model_control(keep_w0=1)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sc = A.solver_control(max_iter=500, solver=A.ALS_solver())
    assert sc["max_iter"] == 100 and "100" in str(w[0].message)            # R/fm_solver_control.R:25-28
    with pytest.raises(ValueError):
        A.track_control(evaluate_metric="F1")


def test_fm_train_argument_errors_before_any_gpu_work():
    X = np.eye(6)
    with pytest.raises(ValueError, match="no labels"):
        A.fm_train(A.fm_matrix(X))
    with pytest.raises(ValueError, match="two levels"):
        A.fm_train(A.fm_matrix(X, [0, 1, 2, 0, 1, 2]))
    with pytest.raises(ValueError, match=r"c\(0, 1\) or c\(-1, 1\)"):
        A.fm_train(A.fm_matrix(X, [1, 2, 1, 2, 1, 2]))
    with pytest.raises(ValueError, match="out of range"):
        A.fm_train(A.fm_matrix(X, [0, 1, 0, 1, 0, 1]), normalize=np.array([0, 3]))
    with pytest.raises(ValueError, match="control list is wrong"):
        A.fm_train(A.fm_matrix(X, [0, 1, 0, 1, 0, 1]), control=[{"a": 1}])
    with pytest.raises(ValueError, match="newdata is null"):
        A.predict({}, None)


def test_options_roundtrip():
    old = A.options()
    A.options(**{"FM.mode": "minibatch", "FM.batch": 1024})
    assert A.get_option("FM.mode") == "minibatch" and A.get_option("FM.batch") == 1024
    A.options(**{"FM.mode": old["FM.mode"], "FM.batch": old["FM.batch"]})
    with pytest.raises(KeyError):
        A.options(**{"FM.nope": 1})


# ---------------------------------------------------------------------------------------------- end to end (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["SGD", "FTRL", "TDAP", "ALS"])
def test_fm_train_predict_matches_oracle_end_to_end(port, solver):
    from oracle import oracle as O
    from tests.util import relerr
    rng = np.random.default_rng(1)
    n, p, k = 400, 30, 3
    X = rng.uniform(0.5, 1.5, (n, p)) * (rng.random((n, p)) < 0.2)
    X[X.sum(1) == 0, 0] = 1.0
    y01 = (rng.random(n) < 0.5).astype(float)
    data = A.fm_matrix(X, y01)
    A.options(**{"FM.precision": "f64", "FM.seed": 5, "FM.mode": "exact", "FM.compat": "reference"})
    mk = {"SGD": A.SGD_solver, "FTRL": A.FTRL_solver, "TDAP": A.TDAP_solver, "ALS": A.ALS_solver}[solver]
    iters = 6 if solver == "ALS" else 2 * (n - 1)
    fit = A.fm_train(data, normalize=False, control=[A.model_control(factor_number=k, L2_w1=0.01), A.solver_control(max_iter=iters, solver=mk()),
                                                    A.track_control(step_size=1 if solver == "ALS" else 200)])
    # the oracle on the same inputs: labels recoded to +-1 (R/fm_train.R:112-122), V drawn factor-major from the same seed
    f = data["features"]
    rowptr = np.concatenate([[0], np.cumsum(f["row_size"])]).astype(np.uint32)
    v0 = (0.01 * np.random.default_rng(5).standard_normal((k, p))).T.copy()
    y = np.where(y01 < 1, -1.0, 1.0)
    sid = {"SGD": O.SGD, "FTRL": O.FTRL, "TDAP": O.TDAP, "ALS": O.ALS}[solver]
    cfg = O.make_cfg(solver=sid, k=k, max_iter=iters, l2_w=0.01, step_size=1 if solver == "ALS" else 200)
    rw0, rw, rv, rt = port.train(cfg, n, p, rowptr, f["col_idx"], f["value"], y, 0.0, np.zeros(p), v0, max_rec=50)
    m = fit["Model"]
    assert relerr(m["w0"], rw0) < 1e-8 and relerr(m["w"], rw) < 1e-8 and relerr(m["v"].T, rv) < 1e-8
    assert relerr(fit["Trace"]["evaluation.train"], rt["eval_train"]) < 1e-8
    assert np.array_equal(fit["Trace"]["trace"][0], rt["rec_index"])
    assert fit["Scales"]["target.range"] == (-1.0, 1.0)
    pr = A.predict(fit, A.fm_matrix(X), normalize=False)
    want = port.predict(cfg, n, p, rowptr, f["col_idx"], f["value"], rw0, rw, rv, 1)
    assert relerr(pr, want) < 1e-8
    # fm.track replays the snapshots on (here) the training data: equals the train trace
    tk = A.fm_track(fit, data, normalize=False)
    assert relerr(tk["test"], fit["Trace"]["evaluation.train"]) < 1e-8
    A.options(**{"FM.precision": "auto"})


@pytest.mark.gpu
def test_normalize_update_and_regression_clamp(port):
    from tests.util import relerr
    rng = np.random.default_rng(2)
    n, p, k = 300, 20, 2
    X = rng.uniform(0.5, 2.5, (n, p)) * (rng.random((n, p)) < 0.3)
    X[X.sum(1) == 0, 1] = 1.0
    y = rng.normal(0, 1, n)
    data = A.fm_matrix(X, y)
    A.options(**{"FM.precision": "f64", "FM.seed": 3})
    ctl = [A.model_control(task="REGRESSION", factor_number=k), A.solver_control(max_iter=n - 1, solver=A.SGD_solver(learn_rate=0.02))]
    fit = A.fm_train(data, normalize=True, control=ctl)
    f = data["features"]
    rowptr = np.concatenate([[0], np.cumsum(f["row_size"])]).astype(np.uint32)
    sval, smean, ssd = port.scales(n, p, rowptr, f["col_idx"], f["value"].astype(np.float32), np.arange(p))
    assert relerr(fit["Scales"]["mean"], smean) < 1e-12 and relerr(fit["Scales"]["std"], ssd) < 1e-12
    pr = A.predict(fit, A.fm_matrix(X), normalize=True)
    assert pr.min() >= y.min() - 1e-12 and pr.max() <= y.max() + 1e-12        # clamp to target.range (FM.cpp:204-210)
    fit2 = A.fm_update(fit, data)
    assert fit2["Scales"]["target.range"] == fit["Scales"]["target.range"]
    assert np.abs(fit2["Model"]["w"] - fit["Model"]["w"]).max() > 0           # continued from the warm start
    with pytest.raises(ValueError, match="not the same"):
        A.fm_update(fit, A.fm_matrix(X, y, feature_names=["x%d" % i for i in range(p)]))
    A.options(**{"FM.precision": "auto"})


@pytest.mark.gpu
def test_fm_update_with_kept_optimizer_state_equals_one_long_run():
    """engine extension: options(FM.keep_state=TRUE) stores the FTRL state in the FM object and fm.update
    continues from it -- train(epoch) + update(epoch) == train(2 epochs); the default drops the state like the reference"""
    rng = np.random.default_rng(4)
    n, p, k = 500, 25, 3
    X = rng.uniform(0.5, 1.5, (n, p)) * (rng.random((n, p)) < 0.25)
    X[X.sum(1) == 0, 0] = 1.0
    y01 = (rng.random(n) < 0.5).astype(float)
    data = A.fm_matrix(X, y01)
    A.options(**{"FM.precision": "f64", "FM.seed": 9, "FM.mode": "exact", "FM.keep_state": True})
    try:
        mk = lambda it: [A.model_control(factor_number=k, L1_w1=1e-3, L2_v=1e-3), A.solver_control(max_iter=it, solver=A.FTRL_solver())]
        long = A.fm_train(data, normalize=False, control=mk(2 * (n - 1)))
        half = A.fm_train(data, normalize=False, control=mk(n - 1))
        assert "State" in half and half["State"]["sw"].shape == (2, p) and half["State"]["sv"].shape == (2, p, k)
        cont = A.fm_update(half, data)
        assert np.array_equal(cont["Model"]["w"], long["Model"]["w"]) and np.array_equal(cont["Model"]["v"], long["Model"]["v"])
        A.options(**{"FM.keep_state": False})
        cold = A.fm_update(half, data)                           # the reference's fm.update: optimizer state restarts
        assert "State" not in cold and not np.array_equal(cold["Model"]["v"], long["Model"]["v"])
    finally:
        A.options(**{"FM.precision": "auto", "FM.keep_state": False})


def test_default_precision_is_the_references_unless_throughput_mode_is_asked_for():
    """a user who calls fm.train() with no options must get the reference's fp64 results (exact mode, every solver;
    the reference's default solver is TDAP); fp32 is the throughput mode's default and otherwise an explicit opt-in"""
    from fmwr_b200 import _lib as L
    old = A.options()
    try:
        A.options(**{"FM.precision": "auto", "FM.mode": "exact"})
        assert all(A._precision(s) == L.F64 for s in ("TDAP", "SGD", "FTRL", "ALS", "MCMC", None))
        A.options(**{"FM.mode": "minibatch"})
        assert A._precision("FTRL") == L.F32 and A._precision("ALS") == L.F64 and A._precision(None) == L.F64
        A.options(**{"FM.precision": "f32", "FM.mode": "exact"})
        assert A._precision("TDAP") == L.F32
    finally:
        A.options(**{"FM.precision": old["FM.precision"], "FM.mode": old["FM.mode"]})


@pytest.mark.gpu
def test_fm_matrix_is_uploaded_once_across_train_predict_update_track():
    """the device copy of an fm.matrix is parked in a slot of that object (the R glue's external pointer), so
    fm.train -> predict -> fm.update -> fm.track move the model and the labels but never X again; results equal the uncached calls"""
    rng = np.random.default_rng(6)
    n, p, k = 3000, 40, 4
    X = rng.uniform(0.5, 1.5, (n, p)) * (rng.random((n, p)) < 0.25)
    X[X.sum(1) == 0, 0] = 1.0
    y01 = (rng.random(n) < 0.5).astype(float)
    ctl = lambda: [A.model_control(factor_number=k, L2_w1=0.01), A.solver_control(max_iter=n - 1, solver=A.FTRL_solver()),
                   A.track_control(step_size=1000)]
    old = A.options()
    try:
        A.options(**{"FM.precision": "f64", "FM.seed": 2, "FM.mode": "exact", "FM.cache": False})
        d0 = A.fm_matrix(X, y01)
        f0 = A.fm_train(d0, normalize=True, control=ctl())
        p0 = A.predict(f0, d0, normalize=True)
        u0 = A.fm_update(f0, d0)
        t0 = A.fm_track(u0, d0)
        A.options(**{"FM.cache": True})
        d1 = A.fm_matrix(X, y01)
        ctx = A._ctx()
        x_bytes = 8 * d1["features"]["size"] + 4 * n                  # per entry col i32 + the value as f32 (narrowed on the host before the copy), row_size i32 per row
        model_bytes = 8 * (1 + p + p * k)
        h0 = ctx.transfer_bytes()[0]
        f1 = A.fm_train(d1, normalize=True, control=ctl())
        h1 = ctx.transfer_bytes()[0]
        assert h1 - h0 >= x_bytes                                       # the first call uploads X ...
        p1 = A.predict(f1, d1, normalize=True)
        u1 = A.fm_update(f1, d1)
        h2 = ctx.transfer_bytes()[0]
        assert h2 - h1 < x_bytes and h2 - h1 <= 4 * model_bytes + 8 * n + 4096      # ... the next ones only models (and re-coded labels)
        # a call that does NOT normalize sees the values as uploaded again
        p_raw = A.predict(f1, d1, normalize=False)
        A.options(**{"FM.cache": False})
        p_raw0 = A.predict(f0, A.fm_matrix(X), normalize=False)
        assert np.array_equal(p_raw, p_raw0)
        A.options(**{"FM.cache": True})
        t1 = A.fm_track(u1, d1)
        for a, b in ((f0, f1), (u0, u1)):
            assert np.array_equal(a["Model"]["w"], b["Model"]["w"]) and np.array_equal(a["Model"]["v"], b["Model"]["v"]) and a["Model"]["w0"] == b["Model"]["w0"]
        assert np.array_equal(p0, p1) and np.array_equal(t0["test"], t1["test"])
    finally:
        A.options(**{"FM.precision": old["FM.precision"], "FM.mode": old["FM.mode"], "FM.cache": old["FM.cache"], "FM.seed": old["FM.seed"]})
