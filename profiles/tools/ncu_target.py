"""Small fixed launch sequence for the ncu captures of round 2 (run under `ncu -k regex:... -s ... -c ...`):
configs[1] matrix resident, then  2 FTRL minibatch epochs, 1 SGD epoch, 1 TDAP epoch (plain launches, no graph replay), 3 predicts.
    python profiles/tools/ncu_target.py [rows]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["FMWR_NO_GRAPH"] = "1"

from fmwr_b200 import _lib as L  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    F, k, B = 39, 32, 65536
    field = 1_000_000 // F
    p = field * F
    ctx = L.Context(0)
    data = L.Data.synth(ctx, n, [field] * F, None, 0, 1, 0.1, 20240601)

    def cfg(solver):
        return L.SolverCfg(solver=solver, max_iter=n - 1, random_step=1, learn_rate=0.01, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0,
                           gamma=1e-4, min_target=-1.0, max_target=1.0, mode=L.MODE_MINIBATCH, batch_size=B, precision=L.F32,
                           compat=L.COMPAT_REFERENCE, step_size=-1)
    for solver, reps, l1 in ((L.FTRL, 2, 1e-3), (L.SGD, 1, 0.0), (L.TDAP, 1, 1e-3)):
        mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k, l2_w0=0.0, l1_w1=l1, l2_w1=1e-3, l1_v=0.0, l2_v=1e-3)
        m = L.Model(ctx, mc, p, L.F32)
        m.init_random(0.0, 0.01, 20240603)
        for _ in range(reps):
            L.train_dev(ctx, m, data, cfg(solver))
        if solver == L.TDAP:
            for _ in range(3):
                L.predict_dev(ctx, m, data, L.LINK_LOGISTIC)
        m.close()
    ctx.sync()
    print("ncu_target done")


if __name__ == "__main__":
    main()
