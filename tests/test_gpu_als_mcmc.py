"""ALS / MCMC coordinate-pass parity against the oracle (reference src/solver/MCMC_ALS_Learner.h), exact
Gauss-Seidel order (nthreads = 1).  Tolerances: 1e-4 relative (north star) for fp32, 1e-8 for fp64;
MCMC with identical injected RNG streams 1e-4 / 1e-8; native RNG: posterior-mean RMSE within 0.5 %."""
import numpy as np
import pytest

from oracle import oracle as O
from fmwr_b200 import _lib as L
from fmwr_b200 import synth
from tests.util import relerr

pytestmark = pytest.mark.gpu


def gpu_als(ctx, prec, ds, y, task, solver, k, w0, w, v, sweeps, enable_v, l2_w0=0.0, streams=None, compat=L.COMPAT_REFERENCE,
            step_size=-1, metric=L.LL, seed=1, k0=1, k1=1):
    d = L.Data.from_csr32(ctx, ds["n"], ds["p"], ds["rowptr"], ds["col"], ds["val"], y)
    mc = L.ModelCfg(task=task, keep_w0=k0, keep_w1=k1, k=k, l2_w0=l2_w0)
    m = L.Model(ctx, mc, ds["p"], prec)
    m.set(w0, w, v)
    sc = L.SolverCfg(solver=solver, max_iter=sweeps, random_step=1, min_target=float(np.min(y)), max_target=float(np.max(y)),
                     mode=L.MODE_EXACT, precision=prec, compat=compat, enable_v=enable_v, step_size=step_size, metric=metric,
                     convergence=1e-4, seed=seed)
    keep = []
    if streams is not None:
        nrm, gam, rnd = streams
        nrm = np.ascontiguousarray(nrm, np.float64); gam = np.ascontiguousarray(gam, np.float64); rnd = np.ascontiguousarray(rnd, np.int32)
        keep = [nrm, gam, rnd]
        sc.normals = L.ptr(nrm); sc.n_normals = nrm.size
        sc.gammas = L.ptr(gam); sc.n_gammas = gam.size
        sc.rands = L.ptr(rnd); sc.n_rands = rnd.size
    tr = L.TraceBuf(120)
    L.train_dev(ctx, m, d, sc, tr, keep=keep)
    out = m.get()
    L.predict_dev(ctx, m, d, L.LINK_NONE)
    pred = L.predict_fetch(ctx, d)
    m.close(); d.close()
    return out, tr.result(), pred


def fields_ds(n, fields, seed, value_mode=1):
    rowptr, col, val, p = synth.fields_csr(n, fields, None, value_mode, seed)
    return dict(n=n, p=p, rowptr=rowptr, col=col, val=val)


@pytest.mark.parametrize("task", [O.CLASSIFICATION, O.REGRESSION])
@pytest.mark.parametrize("enable_v", [0, 1])
@pytest.mark.parametrize("layout", ["fields", "fields_odd", "fields_sparse", "ragged"])
def test_als_matches_oracle(gpu_ctx, port, task, enable_v, layout):
    rng = np.random.default_rng(1)
    if layout == "fields":
        ds = fields_ds(3000, [40, 25, 8], 5)            # 3 independent phases (fields): dense fused path, 4 rows per thread
    elif layout == "fields_odd":
        ds = fields_ds(3001, [40, 2500, 8], 6)          # row count not a multiple of 4 (scalar path), one phase beyond the smem table
    elif layout == "fields_sparse":
        # field-structured but some rows miss a field: phases stay few, rows are no longer implicit (row-major unfused path)
        ds = fields_ds(2000, [30, 20, 10], 7)
        keep = np.ones(ds["col"].size, bool)
        keep[np.arange(0, ds["col"].size, 7)] = False
        cnt = np.add.reduceat(keep.astype(np.int64), ds["rowptr"][:-1].astype(np.int64))
        ds = dict(n=2000, p=ds["p"], rowptr=np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32), col=ds["col"][keep], val=ds["val"][keep])
    else:
        rowptr, col, val = synth.random_csr(600, 70, 6, seed=5, empty_rows=True)     # general CSR: many small phases
        ds = dict(n=600, p=70, rowptr=rowptr, col=col, val=val)
    n, p, k = ds["n"], ds["p"], 4
    y = (np.where(rng.random(n) < 0.5, 1.0, -1.0) if task == O.CLASSIFICATION else rng.normal(0, 1, n)).astype(np.float32)
    w = rng.normal(0, 0.1, p); v = rng.normal(0, 0.1, (p, k)); w0 = 0.1
    sweeps = 6
    cfg = O.make_cfg(task=task, solver=O.ALS, k=k, max_iter=sweeps, enable_v=enable_v, l2_w0=0.1,
                     min_target=float(y.min()), max_target=float(y.max()), step_size=1, metric=O.LL if task == O.CLASSIFICATION else O.RMSE)
    # fields_odd has ~1.2 non-zeros per feature in its wide field: 1/(alpha*A) amplifies the summation-order noise
    t64 = 1e-6 if layout == "fields_odd" else 1e-8
    # fp32 handles: on fields_odd (features with one or two non-zeros: h = x q - x^2 v cancels to rounding noise and
    # 1/(alpha * sum h^2) blows it up) the engine runs the sweep in fp64 on a shadow model and hands the result back
    # (train_als.cu: precision policy), so the fp32 handle meets the north star's 1e-4 on every layout.  "Same inputs" for an
    # fp32 handle means the initial parameters as that handle stores them: on such data the reference itself turns the 1e-8
    # rounding of the initial V into O(1) differences, so the oracle is fed the fp32-rounded values.
    for prec, tol in ((L.F64, t64), (L.F32, 1e-4)):
        if prec == L.F32:
            w0, w, v = float(np.float32(w0)), w.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64)
        rw0, rw, rv, rt = port.train(cfg, n, p, ds["rowptr"], ds["col"], ds["val"], y, w0, w, v, max_rec=20)
        (gw0, gw, gv), gt, _ = gpu_als(gpu_ctx, prec, ds, y, task, L.ALS, k, w0, w, v, sweeps, enable_v, l2_w0=0.1, step_size=1,
                                       metric=L.LL if task == O.CLASSIFICATION else L.RMSE)
        assert relerr(gw0, rw0) < tol and relerr(gw, rw) < tol and relerr(gv, rv) < tol, (prec, relerr(gw0, rw0), relerr(gw, rw), relerr(gv, rv))
        assert gt["n_rec"] == rt["n_rec"] and (gt["rec_index"] == rt["rec_index"]).all()
        assert relerr(gt["eval_train"], rt["eval_train"]) < max(tol, 1e-6) * 10


def test_als_as_shipped_never_updates_v(gpu_ctx, port):
    # F1: with enable_v = 0 (the shipped update_all) V must come back untouched
    rng = np.random.default_rng(2)
    ds = fields_ds(1000, [30, 20], 7)
    y = rng.normal(0, 1, ds["n"]).astype(np.float32)
    w = np.zeros(ds["p"]); v = rng.normal(0, 0.1, (ds["p"], 3))
    (gw0, gw, gv), _, _ = gpu_als(gpu_ctx, L.F64, ds, y, L.REGRESSION, L.ALS, 3, 0.0, w, v, 3, 0)
    assert np.array_equal(gv, v) and np.any(gw != 0)


@pytest.mark.parametrize("task", [O.CLASSIFICATION, O.REGRESSION])
@pytest.mark.parametrize("enable_v", [0, 1])
def test_mcmc_injected_streams_match_oracle(gpu_ctx, port, task, enable_v):
    rng = np.random.default_rng(3)
    ds = fields_ds(800, [30, 20, 6], 9)
    n, p, k = ds["n"], ds["p"], 3
    y = (np.where(rng.random(n) < 0.5, 1.0, -1.0) if task == O.CLASSIFICATION else rng.normal(0, 1, n)).astype(np.float32)
    w = rng.normal(0, 0.1, p); v = rng.normal(0, 0.1, (p, k)); w0 = 0.1
    sweeps = 4
    normals = rng.standard_normal(20000)
    gammas = rng.gamma(30.0, 1.0, 1000)
    rands = rng.integers(0, 2**31 - 1, 400000).astype(np.int32)
    cfg = O.make_cfg(task=task, solver=O.MCMC, k=k, max_iter=sweeps, enable_v=enable_v, l2_w0=0.1,
                     min_target=float(y.min()), max_target=float(y.max()))
    # fp32 + classification runs in fp64 internally (the truncated-normal rejection loops are data-dependent): 1e-4 either way
    for prec, tol in ((L.F64, 1e-8), (L.F32, 1e-4)):
        if prec == L.F32:                      # same inputs: the initial parameters as an fp32 handle stores them
            w0, w, v = float(np.float32(w0)), w.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64)
        port.set_streams(normals, gammas, rands)
        rw0, rw, rv, _ = port.train(cfg, n, p, ds["rowptr"], ds["col"], ds["val"], y, w0, w, v)
        pos = port.stream_pos()
        port.set_streams(None, None, None)
        assert pos["overrun"] == 0
        (gw0, gw, gv), _, _ = gpu_als(gpu_ctx, prec, ds, y, task, L.MCMC, k, w0, w, v, sweeps, enable_v, l2_w0=0.1,
                                      streams=(normals, gammas, rands))
        assert relerr(gw0, rw0) < tol and relerr(gw, rw) < tol and relerr(gv, rv) < tol, (prec, relerr(gw0, rw0), relerr(gw, rw), relerr(gv, rv))


def test_mcmc_native_rng_posterior_mean(gpu_ctx, port):
    """north star: under the native RNG the posterior-mean RMSE matches within 0.5 %.  The reference returns the LAST draw
    (MCMC_ALS_Learner.h:98-127), so the posterior mean is formed outside both codes, identically: every sweep's model is
    snapshotted by the tracker (step_size = 1), predictions of the post-burn-in snapshots are averaged, then scored.
    Oracle: numpy streams through the injection hooks; engine: its own counter-based generator (seed only)."""
    rng = np.random.default_rng(4)
    ds = fields_ds(6000, [60, 40], 11, value_mode=0)
    n, p, k = ds["n"], ds["p"], 4
    score = synth.planted_scores_fast(ds["rowptr"], ds["col"], ds["val"], p, k=4, seed=12, scale=0.5)
    y = (score + 0.1 * rng.standard_normal(n)).astype(np.float32)
    w = np.zeros(p); v = rng.normal(0, 0.1, (p, k))
    sweeps, burn = 80, 30
    cfgp = O.make_cfg(task=O.REGRESSION, k=k)

    def posterior_mean_rmse(snaps):
        acc = np.zeros(n)
        for (sw0, sw, sv) in snaps:
            acc += np.clip(port.predict(cfgp, n, p, ds["rowptr"], ds["col"], ds["val"], sw0, sw, sv, 0), float(y.min()), float(y.max()))
        return float(np.sqrt(np.mean((acc / len(snaps) - y) ** 2)))

    # oracle side: one sweep per call, warm-started from the previous draw, so every draw is visible (the hyper-parameters
    # alpha / lambda / mu restart each call exactly as Learner::init does on fm.update; both sides are driven the same way)
    per = 2 + k
    ow0, ow, ov = 0.0, w.copy(), v.copy()
    osn = []
    for sidx in range(sweeps):
        g = np.empty(per)
        g[0] = rng.gamma((1 + n) / 2.0)
        g[1:] = rng.gamma((1 + p + 1) / 2.0, size=per - 1)
        port.set_streams(rng.standard_normal(4 * (p * (k + 1) + 2 * k + 4)), g, None)
        cfg = O.make_cfg(task=O.REGRESSION, solver=O.MCMC, k=k, max_iter=1, enable_v=1, min_target=float(y.min()), max_target=float(y.max()))
        ow0, ow, ov, _ = port.train(cfg, n, p, ds["rowptr"], ds["col"], ds["val"], y, ow0, ow, ov)
        assert port.stream_pos()["overrun"] == 0
        if sidx >= burn:
            osn.append((ow0, ow.copy(), ov.copy()))
    port.set_streams(None, None, None)
    # engine side: native generator, one sweep per call on a persistent fp32 handle, a different seed per sweep
    d = L.Data.from_csr32(gpu_ctx, n, p, ds["rowptr"], ds["col"], ds["val"], y)
    m = L.Model(gpu_ctx, L.ModelCfg(task=L.REGRESSION, keep_w0=1, keep_w1=1, k=k), p, L.F32)
    m.set(0.0, w, v)
    gsn = []
    for sidx in range(sweeps):
        sc = L.SolverCfg(solver=L.MCMC, max_iter=1, random_step=1, min_target=float(y.min()), max_target=float(y.max()), mode=L.MODE_EXACT,
                         precision=L.F32, compat=L.COMPAT_REFERENCE, enable_v=1, step_size=-1, seed=1000 + sidx)
        L.train_dev(gpu_ctx, m, d, sc)
        if sidx >= burn:
            gsn.append(m.get())
    m.close(); d.close()
    rmse_r, rmse_g = posterior_mean_rmse(osn), posterior_mean_rmse(gsn)
    base = float(np.std(y))
    print("MCMC posterior-mean RMSE: oracle %.5f engine %.5f (%.3f %%), sd(y) %.3f" % (rmse_r, rmse_g, 100 * (rmse_g - rmse_r) / rmse_r, base))
    assert rmse_g < 0.2 * base and rmse_r < 0.2 * base             # both chains found the planted model
    assert abs(rmse_g - rmse_r) / rmse_r < 0.005, (rmse_g, rmse_r)


def test_als_converges_on_planted_fm(gpu_ctx):
    # the survey probe: with the V block enabled ALS recovers a planted FM (RMSE near the noise floor)
    rng = np.random.default_rng(5)
    ds = fields_ds(20000, [300, 200], 13, value_mode=0)
    n, p, k = ds["n"], ds["p"], 4
    score = synth.planted_scores_fast(ds["rowptr"], ds["col"], ds["val"], p, k=4, seed=14, scale=0.7)
    y = (score + 0.1 * rng.standard_normal(n)).astype(np.float32)
    w = np.zeros(p); v = rng.normal(0, 0.1, (p, k))
    (_, _, _), _, p_v = gpu_als(gpu_ctx, L.F32, ds, y, L.REGRESSION, L.ALS, k, 0.0, w, v, 12, 1)
    (_, _, _), _, p_nov = gpu_als(gpu_ctx, L.F32, ds, y, L.REGRESSION, L.ALS, k, 0.0, w, v, 12, 0)
    rmse_v = float(np.sqrt(np.mean((p_v - y) ** 2)))
    rmse_nov = float(np.sqrt(np.mean((p_nov - y) ** 2)))
    assert rmse_v < 0.2 and rmse_v < 0.6 * rmse_nov, (rmse_v, rmse_nov)


def test_als_paths_agree_on_the_same_data(gpu_ctx):
    """the fused field-structured passes (default), the same without the warp-segmented reduction, without the row
    permutation, the row-major streaming passes and the column-parallel kernels are five implementations of one
    Gauss-Seidel sweep: on the same data they must end at the same parameters"""
    import os
    ctx = gpu_ctx
    fields = [900, 260, 24]
    n, k = 20000, 8
    rowptr, col, val, p = synth.fields_csr(n, fields, [0, 1, 0], 0, 51)
    y = synth.labels_from_scores(synth.planted_scores_fast(rowptr, col, val, p, seed=52), "regression", seed=53)
    rng = np.random.default_rng(54)
    v = rng.normal(0, 0.05, (p, k))
    sc = L.SolverCfg(solver=L.ALS, max_iter=3, random_step=1, min_target=float(y.min()), max_target=float(y.max()), mode=L.MODE_EXACT,
                     precision=L.F64, compat=L.COMPAT_REFERENCE, enable_v=1, step_size=-1, seed=1)
    mc = L.ModelCfg(task=L.REGRESSION, keep_w0=1, keep_w1=1, k=k)
    out = {}
    for tag, env in (("default", None), ("no_warpseg", "FMWR_ALS_NO_WARPSEG"), ("no_perm", "FMWR_ALS_NO_PERM"),
                     ("row_major", "FMWR_ALS_NO_DENSE"), ("column", "FMWR_ALS_COLUMN")):
        if env:
            os.environ[env] = "1"
        try:
            d = L.Data.from_csr32(ctx, n, p, rowptr, col, val, y)          # a fresh handle: the layouts are cached per handle
            m = L.Model(ctx, mc, p, L.F64)
            m.set(0.0, np.zeros(p), v)
            L.train_dev(ctx, m, d, sc)
            out[tag] = m.get()
            m.close(); d.close()
        finally:
            if env:
                os.environ.pop(env, None)
    ref = out["default"]
    assert float(np.max(np.abs(ref[2] - v))) > 1e-3
    for tag, got in out.items():
        assert relerr(got[0], ref[0]) < 1e-8 and relerr(got[1], ref[1]) < 1e-8 and relerr(got[2], ref[2]) < 1e-8, tag
