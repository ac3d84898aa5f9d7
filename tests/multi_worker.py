"""torchrun worker for tests/test_gpu_multi.py: feature-parallel FTRL/SGD minibatch on WORLD_SIZE GPUs must equal the
single-GPU minibatch run on the same data (same batches, same arithmetic; only the S_f summation is split by shard)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from fmwr_b200 import _lib as L
    from fmwr_b200 import multi, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = L.Context(local)
    ctx.comm_init(multi.exchange_unique_id(dist, L.Context, rank), rank, world)
    fields = [300] * 7 + [40]
    n, k, B = 12000, 8, 512
    rowptr, col, val, p = synth.fields_csr(n, fields, None, 1, 31)
    score = synth.planted_scores_fast(rowptr, col, val, p, seed=32)
    y = synth.labels_from_scores(score, "classification", seed=33)
    rng = np.random.default_rng(34)
    w = rng.normal(0, 0.05, p); v = rng.normal(0, 0.05, (p, k)); w0 = 0.1
    f0, f1, c0, c1 = multi.field_partition(fields, world)[rank]
    ok = True
    for exchange in ("nccl", "peer"):
      if exchange == "peer":
          # same runs again with the peer windows open: partials exchanged by our own kernels over NVLink, no NCCL call
          multi.open_peer_windows(dist, ctx, rank, world, B, k)
      for solver, prec, tol in ((L.FTRL, L.F64, 1e-9), (L.SGD, L.F64, 1e-9), (L.TDAP, L.F64, 1e-7), (L.FTRL, L.F32, 2e-4)):
          iters = 2 * (n - 1) + 77
          mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k, l1_w1=1e-3, l2_w1=1e-3, l2_v=1e-3)
          sc = L.SolverCfg(solver=solver, max_iter=iters, random_step=1, learn_rate=0.01, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0,
                           gamma=1e-4, min_target=-1.0, max_target=1.0, mode=L.MODE_MINIBATCH, batch_size=B, precision=prec,
                           compat=L.COMPAT_SKIP_ROW0, step_size=-1)
          full = L.Data.from_csr32(ctx, n, p, rowptr, col, val, y)
          part = full.slice_columns(c0, c1)
          full.close()
          m = L.Model(ctx, mc, c1 - c0, prec)
          m.set(w0, w[c0:c1], v[c0:c1])
          L.train_dev(ctx, m, part, sc)
          mine = m.get()
          m.close(); part.close()
          parts = [None] * world
          dist.all_gather_object(parts, mine)
          if rank == 0:
              gw0, gw, gv = multi.gather_model(parts)
              solo = L.Context(local)                      # no communicator: the single-GPU path
              d1 = L.Data.from_csr32(solo, n, p, rowptr, col, val, y)
              m1 = L.Model(solo, mc, p, prec)
              m1.set(w0, w, v)
              L.train_dev(solo, m1, d1, sc)
              sw0, sw, sv = m1.get()
              m1.close(); d1.close(); solo.close()
              err = max(abs(gw0 - sw0), float(np.max(np.abs(gw - sw) / np.maximum(1, np.abs(sw)))),
                        float(np.max(np.abs(gv - sv) / np.maximum(1, np.abs(sv)))))
              moved = float(np.max(np.abs(sv - v)))
              print("exchange=%s solver=%d prec=%d world=%d max rel err vs single GPU = %.3e (params moved by %.3e)" % (exchange, solver, prec, world, err, moved), flush=True)
              ok = ok and err < tol and moved > 1e-3
              for q in parts[1:]:
                  ok = ok and q[0] == parts[0][0]          # w0 is replicated bit for bit
    # k = 32 in fp32 takes the stream forward kernel, which runs the owner's reduction itself (forward_stream.cuh: partials ->
    # owners -> totals -> every rank, all in one kernel, plus the intercept's multiplier sum); same comparison, fp32 tolerance
    k2 = 32
    v2 = rng.normal(0, 0.05, (p, k2))
    for solver in (L.FTRL, L.SGD):
        iters = 2 * (n - 1) + 77
        mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k2, l1_w1=1e-3, l2_w1=1e-3, l2_v=1e-3)
        sc = L.SolverCfg(solver=solver, max_iter=iters, random_step=1, learn_rate=0.01, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0,
                         gamma=1e-4, min_target=-1.0, max_target=1.0, mode=L.MODE_MINIBATCH, batch_size=B, precision=L.F32,
                         compat=L.COMPAT_SKIP_ROW0, step_size=-1)
        full = L.Data.from_csr32(ctx, n, p, rowptr, col, val, y)
        part = full.slice_columns(c0, c1)
        full.close()
        m = L.Model(ctx, mc, c1 - c0, L.F32)
        m.set(w0, w[c0:c1], v2[c0:c1])
        L.train_dev(ctx, m, part, sc)
        mine = m.get()
        m.close(); part.close()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            gw0, gw, gv = multi.gather_model(parts)
            solo = L.Context(local)
            d1 = L.Data.from_csr32(solo, n, p, rowptr, col, val, y)
            m1 = L.Model(solo, mc, p, L.F32)
            m1.set(w0, w, v2)
            L.train_dev(solo, m1, d1, sc)
            sw0, sw, sv = m1.get()
            m1.close(); d1.close(); solo.close()
            err = max(abs(gw0 - sw0), float(np.max(np.abs(gw - sw) / np.maximum(1, np.abs(sw)))),
                      float(np.max(np.abs(gv - sv) / np.maximum(1, np.abs(sv)))))
            moved = float(np.max(np.abs(sv - v2)))
            print("fused exchange (k=32, fp32) solver=%d world=%d max rel err vs single GPU = %.3e (params moved by %.3e)" % (solver, world, err, moved), flush=True)
            ok = ok and err < 2e-4 and moved > 1e-3
    # tracker on the sharded model (step_size > 0): every rank scores its slice's partials through one all-reduce per chunk and
    # records the same train metric the single-GPU run records (reference: SGD_Learner.h:140-166)
    iters = 2 * (n - 1)
    mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k, l1_w1=1e-3, l2_w1=1e-3, l2_v=1e-3)
    sc = L.SolverCfg(solver=L.FTRL, max_iter=iters, random_step=1, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0, min_target=-1.0, max_target=1.0,
                     mode=L.MODE_MINIBATCH, batch_size=B, precision=L.F64, compat=L.COMPAT_SKIP_ROW0, step_size=5000, metric=L.LL, convergence=0.0)
    full = L.Data.from_csr32(ctx, n, p, rowptr, col, val, y)
    part = full.slice_columns(c0, c1)
    full.close()
    m = L.Model(ctx, mc, c1 - c0, L.F64)
    m.set(w0, w[c0:c1], v[c0:c1])
    trb = L.TraceBuf(20)
    L.train_dev(ctx, m, part, sc, trb)
    mine_tr = trb.result()
    m.close(); part.close()
    trs = [None] * world
    dist.all_gather_object(trs, (mine_tr["eval_train"], mine_tr["rec_index"]))
    if rank == 0:
        solo = L.Context(local)
        d1 = L.Data.from_csr32(solo, n, p, rowptr, col, val, y)
        m1 = L.Model(solo, mc, p, L.F64)
        m1.set(w0, w, v)
        tr1 = L.TraceBuf(20)
        L.train_dev(solo, m1, d1, sc, tr1)
        r1 = tr1.result()
        m1.close(); d1.close(); solo.close()
        same_idx = all(np.array_equal(t[1], r1["rec_index"]) for t in trs)
        terr = max(float(np.max(np.abs(t[0] - r1["eval_train"]) / np.maximum(1, np.abs(r1["eval_train"])))) for t in trs) if same_idx and len(r1["eval_train"]) else 1.0
        print("sharded tracker: %d records, max rel err of the train metric vs single GPU = %.3e" % (len(r1["eval_train"]), terr), flush=True)
        ok = ok and same_idx and len(r1["eval_train"]) >= 3 and terr < 1e-9
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    ctx.comm_destroy(); ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_OK" if flag[0] else "MULTI_FAIL", flush=True)
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
