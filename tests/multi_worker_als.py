"""torchrun worker for tests/test_gpu_multi.py: ALS / MCMC on ROW-sharded data (WORLD_SIZE GPUs, one NCCL all-reduce of
the per-feature statistics per coordinate step) must give every rank the parameters of the single-GPU run on the whole
data.  MCMC runs with injected normal / gamma streams, so its draws are the same in both runs."""
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
    fields = [700, 260, 24]
    n, k = 24000, 8
    rowptr, col, val, p = synth.fields_csr(n, fields, [0, 1, 0], 0, 41)
    score = synth.planted_scores_fast(rowptr, col, val, p, seed=42)
    y = synth.labels_from_scores(score, "regression", seed=43)
    rng = np.random.default_rng(44)
    w = np.zeros(p); v = rng.normal(0, 0.05, (p, k)); w0 = 0.0
    r0, r1 = multi.row_partition(n, world)[rank]
    sweeps = 3
    normals = np.ascontiguousarray(rng.normal(size=sweeps * (2 + (k + 1) * (p + 1)) + 64), np.float64)
    gammas = np.ascontiguousarray(rng.gamma(5.0, size=sweeps * (2 + k) + 64), np.float64)
    ok = True
    for exchange in ("nccl", "peer"):
      if exchange == "peer":
          # the same runs with the peer windows open: the per-step statistics are exchanged by our own kernels over NVLink
          multi.open_peer_windows(dist, ctx, rank, world, 4096, k)
      for solver, prec, tol in ((L.ALS, L.F64, 1e-8), (L.ALS, L.F32, 2e-3), (L.MCMC, L.F64, 1e-8)):
          mc = L.ModelCfg(task=L.REGRESSION, keep_w0=1, keep_w1=1, k=k)
          kw = dict(solver=solver, max_iter=sweeps, random_step=1, min_target=float(y.min()), max_target=float(y.max()), mode=L.MODE_EXACT,
                    precision=prec, compat=L.COMPAT_REFERENCE, enable_v=1, step_size=-1, seed=7)
          sc = L.SolverCfg(**kw)
          if solver == L.MCMC:      # injected standard normals / unit gammas (no rand(): regression draws no truncated normals)
              sc.normals = L.ptr(normals); sc.n_normals = normals.size
              sc.gammas = L.ptr(gammas); sc.n_gammas = gammas.size
          e0, e1 = int(rowptr[r0]), int(rowptr[r1])
          shard = L.Data.from_csr32(ctx, r1 - r0, p, (rowptr[r0:r1 + 1] - rowptr[r0]).astype(np.uint32), col[e0:e1], val[e0:e1], y[r0:r1])
          m = L.Model(ctx, mc, p, prec)
          m.set(w0, w, v)
          L.train_dev(ctx, m, shard, sc)
          mine = m.get()
          m.close(); shard.close()
          parts = [None] * world
          dist.all_gather_object(parts, mine)
          if rank == 0:
              solo = L.Context(local)
              d1 = L.Data.from_csr32(solo, n, p, rowptr, col, val, y)
              m1 = L.Model(solo, mc, p, prec)
              m1.set(w0, w, v)
              L.train_dev(solo, m1, d1, sc)
              sw0, sw, sv = m1.get()
              m1.close(); d1.close(); solo.close()
              gw0, gw, gv = parts[0]
              err = max(abs(gw0 - sw0), float(np.max(np.abs(gw - sw) / np.maximum(1, np.abs(sw)))),
                        float(np.max(np.abs(gv - sv) / np.maximum(1, np.abs(sv)))))
              moved = float(np.max(np.abs(sv - v)))
              same = all(q[0] == parts[0][0] and np.array_equal(q[1], parts[0][1]) and np.array_equal(q[2], parts[0][2]) for q in parts[1:])
              print("exchange=%s row-sharded solver=%d prec=%d world=%d max rel err vs single GPU = %.3e (moved %.3e) replicas identical=%s" % (
                  exchange, solver, prec, world, err, moved, same), flush=True)
              ok = ok and err < tol and moved > 1e-3 and same
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    ctx.comm_destroy(); ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_ALS_OK" if flag[0] else "MULTI_ALS_FAIL", flush=True)
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
