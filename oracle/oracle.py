"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the two checker libraries.

  Oracle("port")  -> oracle/libfm_oracle.so      plain-C restatement (oracle/fm_oracle.c)
  Oracle("ref")   -> oracle/_ref/libfmwr_ref.so  reference headers compiled unmodified (oracle/ref_driver.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The engine (fmwr_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

CLASSIFICATION, REGRESSION = 10, 20
MCMC, ALS, SGD, FTRL, TDAP = 100, 200, 300, 500, 600
LL, AUC, ACC, RMSE, MSE, MAE = 0, 111, 222, 333, 444, 555


class Cfg(C.Structure):
    """mirror of fmwr_oracle_cfg (oracle/oracle_abi.h)"""
    _fields_ = [('task', C.c_int), ('solver', C.c_int), ('k0', C.c_int), ('k1', C.c_int), ('k', C.c_int),
                ('l2_w0', C.c_double), ('l1_w', C.c_double), ('l2_w', C.c_double), ('l1_v', C.c_double),
                ('l2_v', C.c_double),
                ('max_iter', C.c_int), ('random_step', C.c_int), ('nthreads', C.c_int),
                ('learn_rate', C.c_double),
                ('alpha_w', C.c_double), ('alpha_v', C.c_double), ('beta_w', C.c_double), ('beta_v', C.c_double),
                ('gamma', C.c_double),
                ('enable_v', C.c_int), ('min_target', C.c_double), ('max_target', C.c_double),
                ('step_size', C.c_int), ('metric', C.c_int), ('convergence', C.c_double)]


class TraceInfo(C.Structure):
    _fields_ = [('n_rec', C.c_int), ('convergent', C.c_int), ('iters_done', C.c_int)]


def make_cfg(**kw):
    d = dict(task=CLASSIFICATION, solver=SGD, k0=1, k1=1, k=2, l2_w0=0.0, l1_w=0.0, l2_w=0.0, l1_v=0.0, l2_v=0.0,
             max_iter=1, random_step=1, nthreads=1, learn_rate=0.01, alpha_w=0.1, alpha_v=0.1, beta_w=1.0,
             beta_v=1.0, gamma=1e-4, enable_v=0, min_target=-1.0, max_target=1.0, step_size=-1, metric=LL,
             convergence=1e-4)
    d.update(kw)
    return Cfg(**d)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build(kind="port"):
    """(re)build a checker library with oracle/Makefile; returns True when the .so exists afterwards"""
    target = "restatement" if kind == "port" else "ref"
    subprocess.run(["make", "-s", "-C", HERE, target], check=False, stdout=subprocess.DEVNULL)
    return os.path.exists(lib_path(kind))


def lib_path(kind):
    return os.path.join(HERE, "libfm_oracle.so") if kind == "port" else os.path.join(HERE, "_ref", "libfmwr_ref.so")


def available(kind):
    return os.path.exists(lib_path(kind))


class Oracle:
    def __init__(self, kind="port"):
        assert kind in ("port", "ref")
        self.kind = kind
        path = lib_path(kind)
        if not os.path.exists(path):
            build(kind)
        self.lib = C.CDLL(path)
        self.pfx = "fmwr_orc_" if kind == "port" else "fmwr_ref_"
        for name in ("pnorm", "dpnorm", "trnorm_left", "trnorm_right", "evaluate"):
            getattr(self.lib, self.pfx + name).restype = C.c_double
        getattr(self.lib, self.pfx + "last_error").restype = C.c_char_p
        self._keep = []

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._f("last_error")().decode())

    # -- streams ---------------------------------------------------------------
    def set_streams(self, normals=None, gammas=None, rands=None):
        n = np.ascontiguousarray(normals, np.float64) if normals is not None else None
        g = np.ascontiguousarray(gammas, np.float64) if gammas is not None else None
        r = np.ascontiguousarray(rands, np.int32) if rands is not None else None
        self._keep = [n, g, r]
        self._f("set_streams")(_p(n), C.c_long(0 if n is None else n.size), _p(g), C.c_long(0 if g is None else g.size),
                               _p(r), C.c_long(0 if r is None else r.size))

    def stream_pos(self):
        out = (C.c_long * 4)()
        self._f("stream_pos")(out)
        return dict(normals=out[0], gammas=out[1], rands=out[2], overrun=out[3])

    # -- forward -----------------------------------------------------------------
    @staticmethod
    def _csr(rowptr, col, val):
        return (np.ascontiguousarray(rowptr, np.uint32), np.ascontiguousarray(col, np.uint32),
                np.ascontiguousarray(val, np.float32))

    def predict(self, cfg, n, p, rowptr, col, val, w0, w, v, link=0):
        rowptr, col, val = self._csr(rowptr, col, val)
        w = np.ascontiguousarray(w, np.float64)
        v = np.ascontiguousarray(v, np.float64)
        out = np.zeros(n, np.float64)
        self._check(self._f("predict")(C.byref(cfg), n, p, col.size, _p(rowptr), _p(col), _p(val), C.c_double(w0),
                                       _p(w), _p(v), int(link), _p(out)))
        return out

    def predict_rows(self, cfg, n, p, rowptr, col, val, w0, w, v):
        rowptr, col, val = self._csr(rowptr, col, val)
        w = np.ascontiguousarray(w, np.float64)
        v = np.ascontiguousarray(v, np.float64)
        out = np.zeros(n, np.float64)
        sums = np.zeros((n, max(cfg.k, 1)), np.float64)
        self._check(self._f("predict_rows")(C.byref(cfg), n, p, col.size, _p(rowptr), _p(col), _p(val),
                                            C.c_double(w0), _p(w), _p(v), _p(out), _p(sums)))
        return out, sums[:, :cfg.k]

    def transpose(self, n, p, rowptr, col, val, use_ref=1):
        rowptr, col, val = self._csr(rowptr, col, val)
        tp = np.zeros(p + 1, np.uint32)
        ti = np.zeros(col.size, np.uint32)
        tv = np.zeros(col.size, np.float32)
        self._check(self._f("transpose")(n, p, col.size, _p(rowptr), _p(col), _p(val), int(use_ref), _p(tp), _p(ti),
                                         _p(tv)))
        return tp, ti, tv

    # -- training ----------------------------------------------------------------
    def train(self, cfg, n, p, rowptr, col, val, y, w0, w, v, use_ref_transpose=0, max_rec=0):
        """returns (w0, w, v[p][k], trace dict)"""
        rowptr, col, val = self._csr(rowptr, col, val)
        y = np.ascontiguousarray(y, np.float32)
        w = np.array(w, np.float64, copy=True)
        v = np.array(v, np.float64, copy=True).reshape(p, max(cfg.k, 0)) if cfg.k > 0 else np.zeros((p, 0))
        v = np.ascontiguousarray(v)
        w0c = C.c_double(w0)
        ev = np.zeros(max(max_rec, 1), np.float64)
        ri = np.zeros(max(max_rec, 1), np.int32)
        info = TraceInfo()
        self._check(self._f("train")(C.byref(cfg), n, p, col.size, _p(rowptr), _p(col), _p(val), _p(y),
                                     int(use_ref_transpose), C.byref(w0c), _p(w), _p(v),
                                     int(max_rec), _p(ev), _p(ri), C.byref(info)))
        nr = min(info.n_rec, max_rec)
        return w0c.value, w, v, dict(n_rec=info.n_rec, convergent=bool(info.convergent), iters_done=info.iters_done,
                                     eval_train=ev[:nr].copy(), rec_index=ri[:nr].copy())

    # -- scalars -----------------------------------------------------------------
    def pnorm(self, x):
        return self._f("pnorm")(C.c_double(x))

    def dpnorm(self, x):
        return self._f("dpnorm")(C.c_double(x))

    def trnorm_left(self, left, mean=0.0, sd=1.0):
        return self._f("trnorm_left")(C.c_double(left), C.c_double(mean), C.c_double(sd))

    def trnorm_right(self, right, mean=0.0, sd=1.0):
        return self._f("trnorm_right")(C.c_double(right), C.c_double(mean), C.c_double(sd))

    def random_select(self, n):
        f = self._f("random_select")
        f.restype = C.c_uint
        return f(int(n))

    def evaluate(self, task, metric, y_hat, y_true):
        y_hat = np.ascontiguousarray(y_hat, np.float64)
        y_true = np.ascontiguousarray(y_true, np.float32)
        return self._f("evaluate")(int(task), int(metric), y_hat.size, _p(y_hat), _p(y_true))

    def scales(self, n, p, rowptr, col, val, norm_cols):
        rowptr, col, val = self._csr(rowptr, col, val)
        val = val.copy()
        nc = np.ascontiguousarray(norm_cols, np.int32)
        mean = np.zeros(p)
        sd = np.zeros(p)
        self._check(self._f("scales")(n, p, col.size, _p(rowptr), _p(col), _p(val), _p(nc), nc.size, _p(mean), _p(sd)))
        return val, mean, sd

    def num_threads(self):
        return self._f("num_threads")()
