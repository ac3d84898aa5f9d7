"""The drop-in boundary on a box without a GPU: the C-ABI library loads, exports every symbol the header declares,
fails loudly instead of falling back to a CPU path, and the product never touches the oracle."""
import ctypes as C
import os
import re

import pytest

from fmwr_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.lib()
    names = L.exported_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.fmwr_version() >= 100


def test_struct_mirrors_match_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "fmwr_b200.h")).read()

    def fields(struct):
        body = dict((nm, b) for b, nm in re.findall(r"typedef struct \{([^}]*)\} (\w+);", hdr))[struct]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
            for nm in names.split(","):
                out.append(nm.replace("*", "").strip())
        return out

    assert fields("fmwr_model_cfg") == [f[0] for f in L.ModelCfg._fields_]
    assert fields("fmwr_solver_cfg") == [f[0] for f in L.SolverCfg._fields_]
    assert fields("fmwr_trace") == [f[0] for f in L.Trace._fields_]


@pytest.mark.skipif(have_gpu(), reason="needs a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(L.FmwrError) as e:
        L.Context(0)
    assert "no CPU fallback" in str(e.value)
    # the one-shot entry points fail the same way
    mc = L.ModelCfg(task=L.CLASSIFICATION, keep_w0=1, keep_w1=1, k=2)
    rc = L.lib().fmwr_predict(C.byref(mc), L.F32, C.c_int64(0), C.c_int64(0), C.c_int64(0), None, None, None, C.c_double(0), None, None,
                              0, C.c_double(0), C.c_double(0), None)
    assert rc != 0 and b"no CPU fallback" in L.lib().fmwr_last_error()


def test_null_arguments_are_errors_not_crashes():
    lib = L.lib()
    assert lib.fmwr_ctx_create(0, None) != 0
    assert lib.fmwr_ctx_sync(None) != 0
    assert lib.fmwr_data_transpose(None) != 0
    assert lib.fmwr_model_get(None, None, None, None) != 0
    assert lib.fmwr_predict_dev(None, None, None, 0, C.c_double(0), C.c_double(0)) != 0
    assert lib.fmwr_train_dev(None, None, None, None, None) != 0
    assert len(lib.fmwr_last_error()) > 0


def test_product_never_uses_the_oracle_or_torch():
    pkg = os.path.join(ROOT, "fmwr_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py", (dirpath, f)
                assert "libfm_oracle" not in txt and "/root/reference" not in txt, (dirpath, f)
                # torch appears only as torch.distributed plumbing for the communicator-id exchange (multi.py)
                assert "import torch" not in txt or f == "multi.py", (dirpath, f)


def test_enum_values_are_the_references():
    # src/util/Macros.h:10-30
    assert (L.CLASSIFICATION, L.REGRESSION) == (10, 20)
    assert (L.MCMC, L.ALS, L.SGD, L.FTRL, L.TDAP) == (100, 200, 300, 500, 600)
    assert (L.LL, L.AUC, L.ACC, L.RMSE, L.MSE, L.MAE) == (0, 111, 222, 333, 444, 555)


def test_peer_window_size_covers_its_regions():
    # fmwr_comm_peer_bytes (host-only arithmetic): control block + world x ceil(B/world) partial rows + B total rows, any precision
    for B, k, world in ((65536, 32, 2), (65536, 32, 8), (512, 8, 2), (1000, 100, 3), (1, 0, 2)):
        got = L.Context.comm_peer_bytes(B, k, world)
        kp = max(4, 1 << (max(k, 1) - 1).bit_length())              # padded factor count never exceeds the next power of two (>= 4)
        stride = kp + 4
        rpo = -(-B // world)
        need = 4096 + 8 * (world * rpo * stride + B * stride) + 2 * 256
        assert got >= need, (B, k, world, got, need)
    assert L.Context.comm_peer_bytes(0, 8, 2) == 0 and L.Context.comm_peer_bytes(16, 8, 0) == 0
    # a window cannot be allocated without a communicator
    assert L.lib().fmwr_comm_peer_alloc(None, C.c_int64(1 << 20), None) != 0
