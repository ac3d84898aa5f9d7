"""Generates tests/golden/fm_golden.npz by running the REFERENCE's own headers (oracle/_ref, built from
/root/reference by oracle/Makefile) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The .npz holds inputs and reference outputs, so the port oracle and the CUDA engine can be checked against the
reference on boxes where /root/reference does not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from fmwr_b200 import synth  # noqa: E402


def main():
    ref = O.Oracle("ref")
    rng = np.random.default_rng(20240601)
    n, p, k = 240, 36, 4
    rowptr, col, val = synth.random_csr(n, p, 7, seed=11)
    w = rng.normal(0, 0.1, p); v = rng.normal(0, 0.1, (p, k)); w0 = 0.25
    ycls = np.where(rng.random(n) < 0.5, 1.0, -1.0).astype(np.float32)
    yreg = rng.normal(0, 1, n).astype(np.float32)
    out = dict(n=n, p=p, k=k, rowptr=rowptr, col=col, val=val, w=w, v=v, w0=w0, ycls=ycls, yreg=yreg)
    # forward
    out["pred_raw"] = ref.predict(O.make_cfg(k=k), n, p, rowptr, col, val, w0, w, v, 0)
    out["pred_logistic"] = ref.predict(O.make_cfg(k=k, solver=O.SGD), n, p, rowptr, col, val, w0, w, v, 1)
    out["pred_probit"] = ref.predict(O.make_cfg(k=k, solver=O.ALS), n, p, rowptr, col, val, w0, w, v, 1)
    # CSC
    tp, ti, tv = ref.transpose(n, p, rowptr, col, val, use_ref=1)
    out.update(csc_ptr=tp, csc_idx=ti, csc_val=tv)
    # link tables on a grid
    xs = np.linspace(-6.5, 6.5, 2601)
    out["xs"] = xs
    out["pnorm"] = np.array([ref.pnorm(x) for x in xs])
    out["dpnorm"] = np.array([ref.dpnorm(x) for x in xs])
    # metrics
    yh = rng.random(n)
    out["yh"] = yh
    out["metrics_cls"] = np.array([ref.evaluate(O.CLASSIFICATION, m, yh, ycls) for m in (O.LL, O.AUC, O.ACC)])
    out["metrics_reg"] = np.array([ref.evaluate(O.REGRESSION, m, yh, yreg) for m in (O.RMSE, O.MAE)])
    # training: every solver, both tasks, 2 epochs + ragged tail
    iters = 2 * (n - 1) + 5
    for sname, solver in (("sgd", O.SGD), ("ftrl", O.FTRL), ("tdap", O.TDAP)):
        for tname, task, y in (("cls", O.CLASSIFICATION, ycls), ("reg", O.REGRESSION, yreg)):
            for rname, regs in (("noreg", {}), ("l1", dict(l1_w=0.01, l1_v=0.01)), ("l2", dict(l2_w=0.01, l2_v=0.02, l2_w0=0.01))):
                cfg = O.make_cfg(task=task, solver=solver, k=k, max_iter=iters, min_target=float(y.min()), max_target=float(y.max()), **regs)
                a, b, c, _ = ref.train(cfg, n, p, rowptr, col, val, y, w0, w, v)
                key = "%s_%s_%s" % (sname, tname, rname)
                out[key + "_w0"] = a; out[key + "_w"] = b; out[key + "_v"] = c
    normals = rng.standard_normal(8000); gammas = rng.gamma(20.0, 1.0, 200); rands = rng.integers(0, 2**31 - 1, 60000).astype(np.int32)
    out.update(normals=normals, gammas=gammas, rands=rands)
    for sname, solver in (("als", O.ALS), ("mcmc", O.MCMC)):
        for tname, task, y in (("cls", O.CLASSIFICATION, ycls), ("reg", O.REGRESSION, yreg)):
            for ev in (0, 1):
                cfg = O.make_cfg(task=task, solver=solver, k=k, max_iter=4, enable_v=ev, l2_w0=0.1, min_target=float(y.min()), max_target=float(y.max()))
                ref.set_streams(normals, gammas, rands)
                a, b, c, _ = ref.train(cfg, n, p, rowptr, col, val, y, w0, w, v, use_ref_transpose=1)
                pos = ref.stream_pos()
                assert pos["overrun"] == 0
                key = "%s_%s_v%d" % (sname, tname, ev)
                out[key + "_w0"] = a; out[key + "_w"] = b; out[key + "_v"] = c
    ref.set_streams(None, None, None)
    # scales
    sval, smean, ssd = ref.scales(n, p, rowptr, col, val, np.arange(0, p, 2))
    out.update(scales_val=sval, scales_mean=smean, scales_sd=ssd)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
