import numpy as np


def relerr(a, b):
    """max |a-b| / max(1, |b|): the north-star tolerance form"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def csr_to_r_lists(rowptr, col, val):
    """CSR -> the fm.matrix lists (value f64, col_idx i32, row_size i32)"""
    rowptr = np.asarray(rowptr, np.int64)
    return (np.diff(rowptr).astype(np.int32), np.asarray(col, np.int32), np.asarray(val, np.float64))
