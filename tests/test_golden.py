"""Golden vectors generated from the reference's own C++ (tests/golden/make_golden.py -> fm_golden.npz).
CPU: the port oracle reproduces them.  GPU: the CUDA engine reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import relerr

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fm_golden.npz"))
N, P, K = int(G["n"]), int(G["p"]), int(G["k"])
CSR = (G["rowptr"], G["col"], G["val"])
ITERS = 2 * (N - 1) + 5
REGS = {"noreg": {}, "l1": dict(l1_w=0.01, l1_v=0.01), "l2": dict(l2_w=0.01, l2_v=0.02, l2_w0=0.01)}
ROW_SOLVERS = {"sgd": O.SGD, "ftrl": O.FTRL, "tdap": O.TDAP}
TASKS = {"cls": (O.CLASSIFICATION, "ycls"), "reg": (O.REGRESSION, "yreg")}


# ---------------------------------------------------------------------------------------------- CPU: port vs golden
def test_port_forward_golden(port):
    w0, w, v = float(G["w0"]), G["w"], G["v"]
    assert relerr(port.predict(O.make_cfg(k=K), N, P, *CSR, w0, w, v, 0), G["pred_raw"]) < 1e-14
    assert relerr(port.predict(O.make_cfg(k=K, solver=O.SGD), N, P, *CSR, w0, w, v, 1), G["pred_logistic"]) < 1e-14
    assert relerr(port.predict(O.make_cfg(k=K, solver=O.ALS), N, P, *CSR, w0, w, v, 1), G["pred_probit"]) < 1e-12


def test_port_transpose_tables_metrics_golden(port):
    tp, ti, tv = port.transpose(N, P, *CSR)
    assert (tp == G["csc_ptr"]).all() and (ti == G["csc_idx"]).all() and (tv == G["csc_val"]).all()
    xs = G["xs"]
    assert relerr([port.pnorm(x) for x in xs], G["pnorm"]) < 1e-12
    assert relerr([port.dpnorm(x) for x in xs], G["dpnorm"]) < 1e-9          # the reference table is printed to ~10 digits
    got = [port.evaluate(O.CLASSIFICATION, m, G["yh"], G["ycls"]) for m in (O.LL, O.AUC, O.ACC)]
    assert relerr(got, G["metrics_cls"]) < 1e-14
    got = [port.evaluate(O.REGRESSION, m, G["yh"], G["yreg"]) for m in (O.RMSE, O.MAE)]
    assert relerr(got, G["metrics_reg"]) < 1e-14
    sval, smean, ssd = port.scales(N, P, *CSR, np.arange(0, P, 2))
    assert (sval == G["scales_val"]).all() and relerr(smean, G["scales_mean"]) < 1e-15 and relerr(ssd, G["scales_sd"]) < 1e-15


@pytest.mark.parametrize("sname", list(ROW_SOLVERS))
@pytest.mark.parametrize("tname", list(TASKS))
@pytest.mark.parametrize("rname", list(REGS))
def test_port_row_solvers_golden(port, sname, tname, rname):
    task, ykey = TASKS[tname]
    y = G[ykey]
    cfg = O.make_cfg(task=task, solver=ROW_SOLVERS[sname], k=K, max_iter=ITERS, min_target=float(y.min()), max_target=float(y.max()),
                     **REGS[rname])
    w0, w, v, _ = port.train(cfg, N, P, *CSR, y, float(G["w0"]), G["w"], G["v"])
    key = "%s_%s_%s" % (sname, tname, rname)
    assert relerr(w0, G[key + "_w0"]) < 1e-13 and relerr(w, G[key + "_w"]) < 1e-13 and relerr(v, G[key + "_v"]) < 1e-13


@pytest.mark.parametrize("sname,solver", [("als", O.ALS), ("mcmc", O.MCMC)])
@pytest.mark.parametrize("tname", list(TASKS))
@pytest.mark.parametrize("ev", [0, 1])
def test_port_coordinate_solvers_golden(port, sname, solver, tname, ev):
    task, ykey = TASKS[tname]
    y = G[ykey]
    cfg = O.make_cfg(task=task, solver=solver, k=K, max_iter=4, enable_v=ev, l2_w0=0.1, min_target=float(y.min()), max_target=float(y.max()))
    port.set_streams(G["normals"], G["gammas"], G["rands"])
    w0, w, v, _ = port.train(cfg, N, P, *CSR, y, float(G["w0"]), G["w"], G["v"])
    port.set_streams(None, None, None)
    key = "%s_%s_v%d" % (sname, tname, ev)
    assert relerr(w0, G[key + "_w0"]) < 1e-9 and relerr(w, G[key + "_w"]) < 1e-9 and relerr(v, G[key + "_v"]) < 1e-9


# ---------------------------------------------------------------------------------------------- GPU: engine vs golden
def _ds():
    return dict(n=N, p=P, rowptr=G["rowptr"], col=G["col"], val=G["val"])


@pytest.mark.gpu
def test_gpu_forward_golden(gpu_ctx):
    from fmwr_b200 import _lib as L
    from tests.test_gpu_forward import run_forward
    w0, w, v = float(G["w0"]), G["w"], G["v"]
    for prec, tol in ((L.F32, 1e-5), (L.F64, 1e-12)):
        assert relerr(run_forward(gpu_ctx, prec, N, P, *CSR, w0, w, v, K), G["pred_raw"]) < tol
        assert relerr(run_forward(gpu_ctx, prec, N, P, *CSR, w0, w, v, K, link=L.LINK_LOGISTIC), G["pred_logistic"]) < tol
        assert relerr(run_forward(gpu_ctx, prec, N, P, *CSR, w0, w, v, K, link=L.LINK_PROBIT_TABLE), G["pred_probit"]) < max(tol, 1e-9)


@pytest.mark.gpu
def test_gpu_transpose_and_tables_golden(gpu_ctx):
    from fmwr_b200 import _lib as L
    d = L.Data.from_csr32(gpu_ctx, N, P, *CSR)
    d.transpose()
    cp, cr, cv = d.get_csc()
    assert (cp == G["csc_ptr"]).all() and (cr == G["csc_idx"]).all() and (cv == G["csc_val"]).all()
    assert relerr(gpu_ctx.link_table(0, G["xs"]), G["pnorm"]) < 1e-12
    assert relerr(gpu_ctx.link_table(1, G["xs"]), G["dpnorm"]) < 1e-9
    mean, sd = d.scales(np.arange(0, P, 2))
    assert relerr(mean, G["scales_mean"]) < 1e-12 and relerr(sd, G["scales_sd"]) < 1e-12
    assert relerr(d.get_csr()[2], G["scales_val"]) < 1e-6
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sname", list(ROW_SOLVERS))
@pytest.mark.parametrize("tname", list(TASKS))
@pytest.mark.parametrize("rname", list(REGS))
def test_gpu_row_solvers_golden(gpu_ctx, sname, tname, rname):
    from fmwr_b200 import _lib as L
    from tests.test_gpu_exact import gpu_train
    task, ykey = TASKS[tname]
    y = G[ykey]
    sid = {"sgd": L.SGD, "ftrl": L.FTRL, "tdap": L.TDAP}[sname]
    (w0, w, v), tr = gpu_train(gpu_ctx, L.F64, _ds(), y, task, sid, K, float(G["w0"]), G["w"], G["v"], ITERS, REGS[rname])
    key = "%s_%s_%s" % (sname, tname, rname)
    assert relerr(w0, G[key + "_w0"]) < 1e-9 and relerr(w, G[key + "_w"]) < 1e-9 and relerr(v, G[key + "_v"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("sname", ["als", "mcmc"])
@pytest.mark.parametrize("tname", list(TASKS))
@pytest.mark.parametrize("ev", [0, 1])
def test_gpu_coordinate_solvers_golden(gpu_ctx, sname, tname, ev):
    from fmwr_b200 import _lib as L
    from tests.test_gpu_als_mcmc import gpu_als
    task, ykey = TASKS[tname]
    y = G[ykey]
    sid = L.ALS if sname == "als" else L.MCMC
    streams = (G["normals"], G["gammas"], G["rands"]) if sname == "mcmc" else None
    (w0, w, v), _, _ = gpu_als(gpu_ctx, L.F64, _ds(), y, task, sid, K, float(G["w0"]), G["w"], G["v"], 4, ev, l2_w0=0.1, streams=streams)
    key = "%s_%s_v%d" % (sname, tname, ev)
    assert relerr(w0, G[key + "_w0"]) < 1e-8 and relerr(w, G[key + "_w"]) < 1e-8 and relerr(v, G[key + "_v"]) < 1e-8
