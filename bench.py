#!/usr/bin/env python
"""bench.py -- headline benchmark of the FM hot path (BASELINE.json configs[1]).

A "step" is one FTRL-Proximal training epoch (minibatch throughput mode) over the Criteo-shaped
synthetic matrix (10M rows x 39 nnz, 1M features, k=32, binary logloss, L1+L2); `value` is samples/s with
the data resident in HBM.  The same run also times predict.FM over the same rows (`predict`), the
end-to-end call through the C ABI with HOST buffers (`e2e`), the reference's CPU path on a bounded sample
(`cpu_baseline`) and reports the dominant kernel's achieved HBM bandwidth (`roofline`).

    python bench.py --gpus N --steps K --warmup W            # engine arm
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU implementation

For N > 1 launch with torchrun (one rank per GPU); see DESIGN.md section "Multi-GPU".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# one metric string for both arms (the reference arm runs the reference's batch = 1 CPU loop; `predict` only on the engine arm)
METRIC = "samples/sec per epoch (fm.train FTRL.solver, L1+L2 binary logloss, configs[1]); predict rows/sec in `predict`"

ALG_BYTES = {
    # fp32 params/state, u32 ids, f32 x, every gather at full width
    "predict": lambda m, k: 8 + m * (12 + 4 * k),
    "sgd": lambda m, k: 8 + m * (8 + 2 * (4 + 4 * k)),
    "ftrl": lambda m, k: 8 + m * (8 + 6 * (4 + 4 * k)),
    "tdap": lambda m, k: 8 + m * (8 + 10 * (4 + 4 * k)),
}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(nm)
        if sm:
            top = sorted(sm)[len(sm) // 2:]          # upper half ~ samples under load
            out.update(sm_mhz=float(np.median(top)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def ncu_traffic(kernel, units_per_launch):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic_r01.json), scaled to this run's launch size"""
    path = os.path.join(ROOT, "profiles", "traffic_r01.json")
    try:
        t = json.load(open(path))[kernel]
        return int(t["dram_bytes"] * units_per_launch / t["rows_per_launch"])
    except Exception:
        return None


def parse_profile(txt):
    agg = {}
    for line in txt.strip().split("\n"):
        if not line:
            continue
        tag, n, ms = line.split("\t")
        name = tag.strip("()").split("<")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += int(n); a[1] += float(ms)
    return agg


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------ engine arm
def run_engine(args):
    from fmwr_b200 import _lib as L
    rank, world, local = dist_env()
    if world > 1:
        return run_engine_multi(args, rank, world, local)
    lib = L.lib()
    ctx = L.Context(local)
    peak, peak_src = measured_peak()
    n, F, k, B = args.rows, 39, args.k, args.batch
    field = args.features // F
    p = field * F
    m_nnz = F

    t0 = time.time()
    data = L.Data.synth(ctx, n, [field] * F, None, 0, 1, 0.1, 20240601)
    mcfg = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k, l2_w0=0.0, l1_w1=1e-3, l2_w1=1e-3, l1_v=0.0, l2_v=1e-3)
    model = L.Model(ctx, mcfg, p, L.F32)
    model.init_random(0.0, 0.01, 20240603)
    ctx.sync()
    t_gen = time.time() - t0

    def scfg(max_iter):
        return L.SolverCfg(solver=L.FTRL, max_iter=max_iter, random_step=1, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0,
                           min_target=-1.0, max_target=1.0, mode=L.MODE_MINIBATCH, batch_size=B, precision=L.F32,
                           compat=L.COMPAT_REFERENCE, step_size=-1)

    epoch = n - 1                       # one reference epoch == n-1 sample updates
    sc = scfg(epoch)
    # ---- FTRL epoch, data resident ------------------------------------------------------------------
    for _ in range(args.warmup):
        L.train_dev(ctx, model, data, sc)
    ctx.sync()
    clocks = ClockSampler(local)
    clocks.start()
    # timed region A: K steps, no per-kernel events -> `value`
    l0 = ctx.launches()
    ctx.timer_start()
    for _ in range(args.steps):
        L.train_dev(ctx, model, data, sc)
    ms_train = ctx.timer_stop_ms()
    l1 = ctx.launches()
    train_sps = epoch * args.steps / (ms_train * 1e-3)
    # timed region B: the same K steps with a CUDA-event pair around every kernel launch (events pre-created by one
    # untimed profiled step) -> per-kernel durations for the roofline
    buf = C.create_string_buffer(1 << 16)
    L.check(lib.fmwr_profile_enable(ctx.h, 1))
    L.train_dev(ctx, model, data, sc)
    L.check(lib.fmwr_profile_enable(ctx.h, 1))          # returns the events to the pool, clears the records
    ctx.timer_start()
    for _ in range(args.steps):
        L.train_dev(ctx, model, data, sc)
    ms_train_prof = ctx.timer_stop_ms()
    L.check(lib.fmwr_profile_read(ctx.h, buf, C.c_int64(len(buf))))
    prof_train = parse_profile(buf.value.decode())
    L.check(lib.fmwr_profile_enable(ctx.h, 0))

    # ---- predict.FM, data resident -------------------------------------------------------------------
    for _ in range(args.warmup):
        L.predict_dev(ctx, model, data, L.LINK_LOGISTIC)
    ctx.timer_start()
    for _ in range(args.steps):
        L.predict_dev(ctx, model, data, L.LINK_LOGISTIC)
    ms_pred = ctx.timer_stop_ms()
    L.check(lib.fmwr_profile_enable(ctx.h, 1))
    for _ in range(args.steps):
        L.predict_dev(ctx, model, data, L.LINK_LOGISTIC)
    L.check(lib.fmwr_profile_read(ctx.h, buf, C.c_int64(len(buf))))
    prof_pred = parse_profile(buf.value.decode())
    L.check(lib.fmwr_profile_enable(ctx.h, 0))
    clk = clocks.stop()
    pred_rps = n * args.steps / (ms_pred * 1e-3)

    # ---- roofline of the dominant kernel (the coordinate-update kernel K2) ---------------------------------
    b_fwd = ALG_BYTES["predict"](m_nnz, k)
    b_ftrl = ALG_BYTES["ftrl"](m_nnz, k)
    roof = None
    kn = "mb_update_kernel"
    if kn in prof_train and prof_train[kn][1] > 0:
        launches, ms = prof_train[kn]
        units = epoch * args.steps                      # samples whose coordinates the launches updated
        ach = units * (b_ftrl - b_fwd) / (ms * 1e-3) / 1e9
        roof = {"kernel": kn, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": ncu_traffic(kn, min(B, n)) if (p, k) == (999999, 32) else None,
                "alg_bytes_per_launch": int((b_ftrl - b_fwd) * units / launches),
                "traffic_note": "ncu dram bytes per launch (profiles/r01_ncu_full.csv); below the algorithmic bytes because a batch "
                                "touches each coordinate ~2.6x and updates it once, and V re-reads hit L2",
                "peak_source": peak_src, "launches": launches, "avg_launch_ms": round(ms / launches, 5),
                "alg_bytes_per_sample": b_ftrl - b_fwd,
                "share_of_step": round(ms / ms_train_prof, 4), "ms_per_step_with_events": round(ms_train_prof / args.steps, 3)}
    fwd_roof = None
    if "forward_kernel" in prof_pred and prof_pred["forward_kernel"][1] > 0:
        launches, ms = prof_pred["forward_kernel"]
        ach = n * args.steps * b_fwd / (ms * 1e-3) / 1e9
        fwd_roof = {"kernel": "forward_kernel", "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": ncu_traffic("forward_kernel", n) if (p, k) == (999999, 32) else None,
                    "alg_bytes_per_row": b_fwd, "avg_launch_ms": round(ms / launches, 4)}
    k1 = prof_train.get("mb_forward_kernel")
    kernels = {name: {"launches": v[0], "ms": round(v[1], 3)} for name, v in prof_train.items()}

    # ---- the other solvers of the north_star (secondary lines) ------------------------------------------------------
    solvers = None
    if not args.no_solvers:
        solvers = run_other_solvers(L, lib, ctx, data, args, peak, n, F, p, k, B)

    # ---- end to end through the C ABI with host buffers -------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(L, lib, ctx, data, mcfg, scfg(epoch), n, p, k, args)

    # ---- CPU baseline (the reference's own C++ where available, else the C port) on a bounded sample ----------------
    cpu = None
    if not args.no_cpu:
        cpu = cpu_baseline_ftrl(L, data, n, p, k, F, seconds=args.cpu_seconds)

    out = {
        "metric": METRIC,
        "value": round(train_sps, 1), "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_train / args.steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: Criteo-shaped %d rows x %d nnz, %d features, k=%d, FTRL L1+L2 binary logloss" % (n, F, p, k),
                   "rows": n, "nnz_per_row": F, "features": p, "k": k, "batch_size": B, "mode": "minibatch",
                   "l2_flush": "inputs (%.1f GB CSR + %.1f GB params/state) larger than L2" % (n * F * 8 / 1e9, p * (k + 1) * 12 / 1e9),
                   "epoch": "n-1 sample updates (reference scan skips row 0)"},
        "roofline": roof,
        "step_roofline": {"alg_bytes_per_sample": b_ftrl, "achieved": round(train_sps * b_ftrl / 1e9, 1), "unit": "GB/s",
                          "frac": round(train_sps * b_ftrl / 1e9 / peak, 4)},
        "kernels": kernels,
        "predict": {"value": round(pred_rps, 1), "unit": "rows/s", "ms_per_step": round(ms_pred / args.steps, 3),
                    "roofline": fwd_roof},
        "solvers": solvers,
        "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(l1 - l0),
        "clocks": clk, "setup_s": round(t_gen, 2),
    }
    print(json.dumps(out), flush=True)
    data.close(); model.close(); ctx.close()


def run_other_solvers(L, lib, ctx, data, args, peak, n, F, p, k, B):
    """the north_star's other training paths, same timing rules (3 warm-ups, CUDA events, data resident): SGD and TDAP
    minibatch epochs on the configs[1] matrix already on the device; one ALS sweep on configs[2] and one MCMC sweep on
    configs[3] (MovieLens-shaped 20M ratings, 3 one-hot fields).  Fractions are against SURVEY 8(d)'s algorithmic bytes."""
    out = {}
    steps = max(1, min(args.steps, 3))

    def timed(fn):
        for _ in range(3):
            fn()
        ctx.sync(); ctx.timer_start()
        for _ in range(steps):
            fn()
        return ctx.timer_stop_ms() / steps

    for name, solver in (("sgd", L.SGD), ("tdap", L.TDAP)):
        mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k, l2_w0=0.0, l1_w1=0.0 if name == "sgd" else 1e-3, l2_w1=1e-3,
                        l1_v=0.0, l2_v=1e-3)
        m = L.Model(ctx, mc, p, L.F32)
        m.init_random(0.0, 0.01, 20240603)
        sc = L.SolverCfg(solver=solver, max_iter=n - 1, random_step=1, learn_rate=0.01, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0,
                         gamma=1e-4, min_target=-1.0, max_target=1.0, mode=L.MODE_MINIBATCH, batch_size=B, precision=L.F32,
                         compat=L.COMPAT_REFERENCE, step_size=-1)
        ms = timed(lambda: L.train_dev(ctx, m, data, sc))
        b = ALG_BYTES[name](F, k)
        sps = (n - 1) / (ms * 1e-3)
        out[name] = {"workload": "configs[1] matrix, %s minibatch epoch (batch %d)" % (name.upper(), B), "value": round(sps, 1), "unit": "samples/s",
                     "ms_per_step": round(ms, 3), "alg_bytes_per_sample": b, "achieved_gbs": round(sps * b / 1e9, 1), "frac": round(sps * b / 1e9 / peak, 4)}
        m.close()
    # exact mode (batch = 1, the reference's own sample order and arithmetic): one persistent CTA, latency-bound by
    # construction (every sample reads w0 written by the previous one), so it is reported in samples/s on a bounded
    # sample of the same matrix, next to the reference's single-thread rate in cpu_baseline -- no HBM fraction.
    it = int(min(n - 1, 200_000))
    ex = {}
    for name, solver in (("sgd", L.SGD), ("ftrl", L.FTRL), ("tdap", L.TDAP)):
        mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k, l2_w0=0.0, l1_w1=1e-3, l2_w1=1e-3, l1_v=0.0, l2_v=1e-3)
        m = L.Model(ctx, mc, p, L.F32)
        m.init_random(0.0, 0.01, 20240603)
        sc = L.SolverCfg(solver=solver, max_iter=it, random_step=1, learn_rate=0.01, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0,
                         gamma=1e-4, min_target=-1.0, max_target=1.0, mode=L.MODE_EXACT, batch_size=1, precision=L.F32,
                         compat=L.COMPAT_REFERENCE, step_size=-1)
        L.train_dev(ctx, m, data, sc)
        ctx.sync(); ctx.timer_start()
        L.train_dev(ctx, m, data, sc)
        ms = ctx.timer_stop_ms()
        ex[name] = {"value": round(it / (ms * 1e-3), 1), "unit": "samples/s", "us_per_sample": round(ms * 1e3 / it, 3)}
        m.close()
    out["exact"] = {"workload": "configs[1] matrix, batch=1 reference-order updates, first %d samples, fp32" % it, **ex}
    if args.rows < 1_000_000:
        return out                                  # reduced smoke runs: skip the 20M-rating sweeps
    na = 20_000_000
    fields = [138493, 26744, 2048]
    pa, Na = sum(fields), 3 * na
    da = L.Data.synth(ctx, na, fields, [0, 1, 0], 0, 3, 0.3, 20240601)
    for name, solver, ka in (("als", L.ALS, 32), ("mcmc", L.MCMC, 64)):
        mc = L.ModelCfg(task=L.REGRESSION, keep_w0=1, keep_w1=1, k=ka)
        m = L.Model(ctx, mc, pa, L.F32)
        m.init_random(0.0, 0.01, 20240603)
        sc = L.SolverCfg(solver=solver, max_iter=1, random_step=1, min_target=0.5, max_target=5.0, mode=L.MODE_EXACT, precision=L.F32,
                         compat=L.COMPAT_REFERENCE, enable_v=1, step_size=-1, seed=5)
        ms = timed(lambda: L.train_dev(ctx, m, da, sc))
        bfwd = 8 + 3 * (12 + 4 * ka)
        sweep_bytes = na * bfwd + 16 * na + 16 * Na + ka * (32 * Na + 4 * na + 8 * pa) + (8 * na if name == "mcmc" else 0)
        out[name] = {"workload": "configs[%d]: %d ratings x 3 one-hot fields (138493 / 26744 skewed / 2048), k=%d, one %s sweep (w and V blocks)" % (
                         2 if name == "als" else 3, na, ka, name.upper()),
                     "value": round(na / (ms * 1e-3), 1), "unit": "ratings/s", "ms_per_step": round(ms, 3), "alg_bytes_per_sweep": sweep_bytes,
                     "achieved_gbs": round(sweep_bytes / (ms * 1e-3) / 1e9, 1), "frac": round(sweep_bytes / (ms * 1e-3) / 1e9 / peak, 4)}
        m.close()
    da.close()
    return out


def run_e2e(L, lib, ctx, data, mcfg, sc, n, p, k, args):
    """fmwr_train one-shot: host fm.matrix lists in, host (w0, w, V) out, every step"""
    rowptr, col, val, y = data.get_csr()
    row_size = np.diff(rowptr.astype(np.int64)).astype(np.int32)
    col_i = col.view(np.int32)
    val64 = val.astype(np.float64)
    y64 = y.astype(np.float64)
    del rowptr, val
    w0 = C.c_double(0.0)
    w = np.zeros(p)
    rng = np.random.default_rng(20240603)
    v = rng.normal(0.0, 0.01, (p, k))
    pinned = []
    for a in (row_size, col_i, val64, y64, w, v):
        if lib.fmwr_host_pin(L.ptr(a), C.c_int64(a.nbytes)) == 0:
            pinned.append(a)
    h2d = row_size.nbytes + col_i.nbytes + val64.nbytes + y64.nbytes + w.nbytes + v.nbytes + 8
    d2h = w.nbytes + v.nbytes + 8
    steps = max(1, min(args.steps, args.e2e_steps))

    def one():
        L.check(lib.fmwr_train(C.byref(mcfg), C.byref(sc), C.c_int64(n), C.c_int64(p), C.c_int64(col_i.size), L.ptr(row_size),
                               L.ptr(col_i), L.ptr(val64), L.ptr(y64), C.byref(w0), L.ptr(w), L.ptr(v), None))
    one()                                   # warm-up (creates the default context)
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    # predict one-shot as well
    out = np.zeros(n)
    lib.fmwr_host_pin(L.ptr(out), C.c_int64(out.nbytes))

    def pone():
        L.check(lib.fmwr_predict(C.byref(mcfg), L.F32, C.c_int64(n), C.c_int64(p), C.c_int64(col_i.size), L.ptr(row_size), L.ptr(col_i),
                                 L.ptr(val64), w0, L.ptr(w), L.ptr(v), L.LINK_LOGISTIC, C.c_double(0), C.c_double(0), L.ptr(out)))
    pone()
    t1 = time.perf_counter()
    for _ in range(steps):
        pone()
    dtp = time.perf_counter() - t1
    for a in pinned:
        lib.fmwr_host_unpin(L.ptr(a))
    lib.fmwr_host_unpin(L.ptr(out))
    return {"value": round((n - 1) * steps / dt, 1), "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "ms_per_step": round(dt / steps * 1e3, 2), "call": "fmwr_train (host fm.matrix lists -> host model)",
            "predict": {"value": round(n * steps / dtp, 1), "unit": "rows/s", "ms_per_step": round(dtp / steps * 1e3, 2),
                        "h2d_bytes_per_step": int(row_size.nbytes + col_i.nbytes + val64.nbytes + w.nbytes + v.nbytes),
                        "d2h_bytes_per_step": int(out.nbytes), "call": "fmwr_predict"}}


# ------------------------------------------------------------------------------------------------ CPU legs
def host_prefix(L, data, rows):
    """first `rows` rows of the device-resident synthetic matrix, on the host"""
    rowptr, col, val, y = data.get_csr()
    e = int(rowptr[rows])
    return rowptr[:rows + 1].copy(), col[:e].copy(), val[:e].copy(), y[:rows].copy()


def oracle_for_baseline():
    from oracle import oracle as O
    if O.available("ref"):
        return O, O.Oracle("ref"), "reference"
    if not O.available("port"):
        O.build("port")
    return O, O.Oracle("port"), "port"


def time_ftrl_cpu(O, orc, rowptr, col, val, y, p, k, rows):
    rng = np.random.default_rng(20240603)
    w = np.zeros(p)
    v = rng.normal(0, 0.01, (p, k))
    cfg = O.make_cfg(solver=O.FTRL, k=k, max_iter=rows - 1, l1_w=1e-3, l2_w=1e-3, l2_v=1e-3, nthreads=1)
    t0 = time.perf_counter()
    orc.train(cfg, rows, p, rowptr[:rows + 1], col[:rowptr[rows]], val[:rowptr[rows]], y[:rows], 0.0, w, v)
    return (rows - 1) / (time.perf_counter() - t0)


def cpu_baseline_ftrl(L, data, n, p, k, F, seconds=15.0):
    O, orc, kind = oracle_for_baseline()
    probe = min(n, 4000)
    cap = min(n, 400_000)
    rowptr, col, val, y = host_prefix(L, data, cap)
    sps = time_ftrl_cpu(O, orc, rowptr, col, val, y, p, k, probe)
    rows = int(max(probe, min(cap, sps * seconds)))
    sps = time_ftrl_cpu(O, orc, rowptr, col, val, y, p, k, rows)
    # predict_batch with every host thread (the reference's only OpenMP row loop, src/core/Model.h:123,138)
    nt = orc.num_threads()
    rng = np.random.default_rng(1)
    w = rng.normal(0, 0.1, p); v = rng.normal(0, 0.01, (p, k))
    cfgp = O.make_cfg(k=k, nthreads=nt)
    t0 = time.perf_counter()
    orc.predict(cfgp, cap, p, rowptr, col, val, 0.0, w, v, 1)
    rps = cap / (time.perf_counter() - t0)
    return {"value": round(sps, 1), "unit": "samples/s", "cores": 1, "kind": kind,
            "sample": "FTRL epoch on the first %d rows of the same synthetic matrix (same p, k, nnz); 1 thread -- the reference has no parallel "
                      "sample loop (todo_list.md:7)" % rows,
            "predict": {"value": round(rps, 1), "unit": "rows/s", "cores": nt, "sample": "predict_prob on the first %d rows, %d OpenMP threads" % (cap, nt)}}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the FTRL epoch, bounded sample per step"""
    rank, world, local = dist_env()
    if rank != 0:
        return
    from fmwr_b200 import synth
    O, orc, kind = oracle_for_baseline()
    F, k = 39, args.k
    field = args.features // F
    p = field * F
    cap = min(args.rows, 200_000)
    rowptr, col, val, _ = synth.fields_csr(cap, [field] * F, None, 0, 20240601)
    score = synth.planted_scores_fast(rowptr, col, val, p, seed=20240602)
    y = synth.labels_from_scores(score, "classification", seed=20240604)
    sps = time_ftrl_cpu(O, orc, rowptr, col, val, y, p, k, min(cap, 4000))
    rows = int(max(4000, min(cap, sps * args.ref_step_seconds)))
    for _ in range(args.warmup):
        time_ftrl_cpu(O, orc, rowptr, col, val, y, p, k, min(rows, 4000))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_ftrl_cpu(O, orc, rowptr, col, val, y, p, k, rows)
    dt = time.perf_counter() - t0
    value = (rows - 1) * args.steps / dt
    sample = "each step = one FTRL pass over the first %d rows of the configs[1] matrix (p=%d, k=%d, 39 nnz), 1 thread" % (rows, p, k)
    out = {"impl": "reference", "metric": METRIC,
           "value": round(value, 1), "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "configs[1]: Criteo-shaped %d rows x 39 nnz, %d features, k=%d, FTRL L1+L2 binary logloss" % (args.rows, p, k),
                      "rows": args.rows, "nnz_per_row": 39, "features": p, "k": k},
           "cpu_baseline": {"value": round(value, 1), "unit": "samples/s", "cores": 1, "kind": kind, "sample": sample},
           "e2e": {"value": round(value, 1), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_engine_multi(args, rank, world, local):
    """N > 1 (torchrun, one rank per GPU): FTRL minibatch epoch FEATURE-parallel with one NCCL all-reduce per batch;
    predict.FM row-sharded with no collective.  Strong scaling: the configs[1] problem is fixed, N grows."""
    import torch
    import torch.distributed as dist
    from fmwr_b200 import _lib as L
    from fmwr_b200 import multi
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)       # control plane only (id exchange, barriers, max-reduce)
    lib = L.lib()
    ctx = L.Context(local)
    ctx.comm_init(multi.exchange_unique_id(dist, L.Context, rank), rank, world)
    peak, peak_src = measured_peak()
    n, F, k, B = args.rows, 39, args.k, args.batch
    use_peer = os.environ.get("FMWR_BENCH_NCCL", "0") != "1"
    if use_peer:
        multi.open_peer_windows(dist, ctx, rank, world, B, k)      # per-batch exchange inside our kernels (NVLink), no NCCL call
    field = args.features // F
    p = field * F
    f0, f1, c0, c1 = multi.field_partition([field] * F, world)[rank]

    full = L.Data.synth(ctx, n, [field] * F, None, 0, 1, 0.1, 20240601)
    data = full.slice_columns(c0, c1)
    full.close()
    mcfg = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=k, l2_w0=0.0, l1_w1=1e-3, l2_w1=1e-3, l1_v=0.0, l2_v=1e-3)
    model = L.Model(ctx, mcfg, c1 - c0, L.F32)
    model.init_random(0.0, 0.01, 20240603 + c0)
    epoch = n - 1
    sc = L.SolverCfg(solver=L.FTRL, max_iter=epoch, random_step=1, alpha_w=0.1, alpha_v=0.1, beta_w=1.0, beta_v=1.0, min_target=-1.0,
                     max_target=1.0, mode=L.MODE_MINIBATCH, batch_size=B, precision=L.F32, compat=L.COMPAT_REFERENCE, step_size=-1)

    def timed(fn, steps):
        ctx.sync(); dist.barrier()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop_ms()
        ctx.sync(); dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                      # device time, max over ranks
        return float(t[0])

    for _ in range(args.warmup):
        L.train_dev(ctx, model, data, sc)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ctx.launches()
    ms_train = timed(lambda: L.train_dev(ctx, model, data, sc), args.steps)
    l1 = ctx.launches()
    train_sps = epoch * args.steps / (ms_train * 1e-3)
    # per-kernel events on rank 0's stream (second pass)
    buf = C.create_string_buffer(1 << 16)
    L.check(lib.fmwr_profile_enable(ctx.h, 1))
    L.train_dev(ctx, model, data, sc)
    L.check(lib.fmwr_profile_enable(ctx.h, 1))
    ms_prof = timed(lambda: L.train_dev(ctx, model, data, sc), args.steps)
    L.check(lib.fmwr_profile_read(ctx.h, buf, C.c_int64(len(buf))))
    prof = parse_profile(buf.value.decode())
    L.check(lib.fmwr_profile_enable(ctx.h, 0))
    data.close(); model.close()

    # predict.FM: rows sharded, full model replicated
    r0, r1 = multi.row_partition(n, world)[rank]
    pdata = L.Data.synth_rows(ctx, r0, r1 - r0, [field] * F, None, 0, 0, 0.1, 20240601)
    pmodel = L.Model(ctx, mcfg, p, L.F32)
    pmodel.init_random(0.0, 0.01, 20240603)
    for _ in range(args.warmup):
        L.predict_dev(ctx, pmodel, pdata, L.LINK_LOGISTIC)
    ms_pred = timed(lambda: L.predict_dev(ctx, pmodel, pdata, L.LINK_LOGISTIC), args.steps)
    pred_rps = n * args.steps / (ms_pred * 1e-3)
    pdata.close(); pmodel.close()

    # ALS sweep, rows sharded: configs[2] split into `world` row ranges, one NCCL all-reduce of the field's
    # statistics per coordinate step; every rank ends the sweep with the same model
    als = None
    if not args.no_solvers and args.rows >= 1_000_000:
        na, fields, ka = 20_000_000, [138493, 26744, 2048], 32
        a0, a1 = multi.row_partition(na, world)[rank]
        ad = L.Data.synth_rows(ctx, a0, a1 - a0, fields, [0, 1, 0], 0, 3, 0.3, 20240601)
        am = L.Model(ctx, L.ModelCfg(task=L.REGRESSION, keep_w0=1, keep_w1=1, k=ka), sum(fields), L.F32)
        am.init_random(0.0, 0.01, 20240603)
        asc = L.SolverCfg(solver=L.ALS, max_iter=1, random_step=1, min_target=0.5, max_target=5.0, mode=L.MODE_EXACT, precision=L.F32,
                          compat=L.COMPAT_REFERENCE, enable_v=1, step_size=-1, seed=5)
        for _ in range(args.warmup):
            L.train_dev(ctx, am, ad, asc)
        ms_als = timed(lambda: L.train_dev(ctx, am, ad, asc), args.steps) / args.steps
        pa_, Na_ = sum(fields), 3 * na
        sweep_bytes = na * (8 + 3 * (12 + 4 * ka)) + 16 * na + 16 * Na_ + ka * (32 * Na_ + 4 * na + 8 * pa_)
        als = {"workload": "configs[2]: 20000000 ratings x 3 one-hot fields, k=32, one ALS sweep, rows sharded over %d GPUs "
                           "(one NCCL all-reduce of 2 x |field| f32 per coordinate step, 99 per sweep)" % world,
               "value": round(na / (ms_als * 1e-3), 1), "unit": "ratings/s", "ms_per_step": round(ms_als, 3),
               "frac_of_n_gpus_peak": round(sweep_bytes / (ms_als * 1e-3) / 1e9 / (peak * world), 4)}
        ad.close(); am.close()
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        b_fwd, b_ftrl = ALG_BYTES["predict"](F, k), ALG_BYTES["ftrl"](F, k)
        roof = None
        if "mb_update_kernel" in prof and prof["mb_update_kernel"][1] > 0:
            launches, ms = prof["mb_update_kernel"]
            share = (f1 - f0) / F                                      # rank 0's share of every sample's coordinates
            ach = epoch * args.steps * (b_ftrl - b_fwd) * share / (ms * 1e-3) / 1e9
            roof = {"kernel": "mb_update_kernel", "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                    "traffic": None, "peak_source": peak_src, "launches": launches, "avg_launch_ms": round(ms / launches, 5),
                    "note": "rank 0 only: %d of %d fields" % (f1 - f0, F), "share_of_step": round(ms / ms_prof, 4)}
        out = {
            "metric": METRIC,
            "value": round(train_sps, 1), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_train / args.steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: Criteo-shaped %d rows x %d nnz, %d features, k=%d, FTRL L1+L2 binary logloss" % (n, F, p, k),
                       "rows": n, "nnz_per_row": F, "features": p, "k": k, "batch_size": B, "mode": "minibatch",
                       "parallelism": "feature-parallel x%d (fields split %s), per minibatch one exchange of rows x (k+4) f32 partials %s; predict row-sharded" % (
                           world, [b - a for a, b, _, _ in multi.field_partition([field] * F, world)],
                           "inside the forward/exchange/update kernels (peer-memory stores over NVLink + in-kernel flags)" if use_peer else "by NCCL all-reduce"),
                       "l2_flush": "inputs larger than L2"},
            "roofline": roof,
            "step_roofline": {"alg_bytes_per_sample": b_ftrl, "achieved": round(train_sps * b_ftrl / 1e9, 1), "unit": "GB/s",
                              "frac_of_n_gpus_peak": round(train_sps * b_ftrl / 1e9 / (peak * world), 4)},
            "kernels": {name: {"launches": v[0], "ms": round(v[1], 3)} for name, v in prof.items()},
            "predict": {"value": round(pred_rps, 1), "unit": "rows/s", "ms_per_step": round(ms_pred / args.steps, 3),
                        "roofline": {"bound": "hbm", "achieved": round(pred_rps * b_fwd / 1e9, 1), "unit": "GB/s",
                                     "frac_of_n_gpus_peak": round(pred_rps * b_fwd / 1e9 / (peak * world), 4)}},
            "solvers": {"als": als} if als else None,
            "cpu_baseline": None,
            "e2e": None,
            "gpu_launches": int(l1 - l0),
            "clocks": clk,
        }
        print(json.dumps(out), flush=True)
    ctx.comm_destroy()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--features", type=int, default=1_000_000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solvers", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-step-seconds", type=float, default=4.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = max(args.warmup, 3) if os.environ.get("FMWR_BENCH_STRICT", "1") == "1" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
