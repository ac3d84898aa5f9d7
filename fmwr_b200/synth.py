"""Host (numpy) twin of the device synthetic-data generator (fmwr_b200/csrc/data.cu: synth_fill).

Integer-only hashing so the ids and values are bit-identical on CPU and GPU:
    h(seed, entry)  = splitmix64(seed ^ entry),  entry = row * F + field
    id(row, field)  = offset_f + (h mod size_f)                  (uniform)
                    = offset_f + ((u^3 >> 64) * size_f >> 32)    (power-law skew, u = h >> 32)
    x               = 1.0  or  0.5 + (splitmix64(h ^ C) >> 40) / 2^24
Shapes follow SURVEY.md section 8(d): Criteo-shaped (39 fields), C1 (10 x 1000), MovieLens-shaped (3 fields).
Labels here come from a planted FM evaluated in numpy fp64 (tests upload them; the device generator
draws its own labels for benchmarks).
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _ids(h, size, skew):
    size = np.uint64(size)
    if not skew:
        return h % size
    u = h >> np.uint64(32)
    u2 = (u * u) >> np.uint64(32)
    u3 = (u2 * u) >> np.uint64(32)
    return (u3 * size) >> np.uint64(32)


def fields_csr(n, field_size, skew=None, value_mode=0, seed=20240601):
    """returns rowptr u32[n+1], col u32[n*F], val f32[n*F], p"""
    field_size = [int(s) for s in field_size]
    F = len(field_size)
    skew = list(skew) if skew is not None else [0] * F
    offs = np.concatenate([[0], np.cumsum(field_size)]).astype(np.uint64)
    entry = np.arange(n * F, dtype=np.uint64)
    h = splitmix64(np.uint64(seed) ^ entry).reshape(n, F)
    col = np.empty((n, F), np.uint64)
    for f in range(F):
        col[:, f] = offs[f] + _ids(h[:, f], field_size[f], skew[f])
    if value_mode:
        h2 = splitmix64(h ^ np.uint64(0xD1B54A32D192ED03))
        val = (np.float32(0.5) + (h2 >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
    else:
        val = np.ones((n, F), np.float32)
    rowptr = (np.arange(n + 1, dtype=np.uint64) * np.uint64(F)).astype(np.uint32)
    return rowptr, col.reshape(-1).astype(np.uint32), val.reshape(-1), int(offs[-1])


def planted_scores(rowptr, col, val, p, k=8, seed=20240602, scale=0.1):
    """fp64 FM score of a planted model (w*, V* ~ N(0, scale^2))"""
    rng = np.random.default_rng(seed)
    w = rng.normal(0, scale, p)
    V = rng.normal(0, scale, (p, k))
    n = rowptr.size - 1
    out = np.zeros(n)
    for i in range(n):
        c = col[rowptr[i]:rowptr[i + 1]]
        x = val[rowptr[i]:rowptr[i + 1]].astype(np.float64)
        vx = V[c] * x[:, None]
        s = vx.sum(0)
        out[i] = (w[c] * x).sum() + 0.5 * (s * s - (vx * vx).sum(0)).sum()
    return out


def planted_scores_fast(rowptr, col, val, p, k=8, seed=20240602, scale=0.1):
    """vectorised variant for constant-width rows"""
    n = rowptr.size - 1
    F = col.size // max(n, 1)
    assert n * F == col.size
    rng = np.random.default_rng(seed)
    w = rng.normal(0, scale, p)
    V = rng.normal(0, scale, (p, k))
    c = col.reshape(n, F)
    x = val.reshape(n, F).astype(np.float64)
    vx = V[c] * x[:, :, None]
    s = vx.sum(1)
    return (w[c] * x).sum(1) + 0.5 * (s * s - (vx * vx).sum(1)).sum(1)


def labels_from_scores(score, mode, noise=0.1, seed=20240604):
    rng = np.random.default_rng(seed)
    if mode == "classification":
        u = rng.random(score.size)
        return np.where(u < 1.0 / (1.0 + np.exp(-score)), 1.0, -1.0).astype(np.float32)
    if mode == "regression":
        return (score + noise * rng.standard_normal(score.size)).astype(np.float32)
    if mode == "rating":
        return np.clip(3.5 + score + noise * rng.standard_normal(score.size), 0.5, 5.0).astype(np.float32)
    raise ValueError(mode)


# ---- named shapes -----------------------------------------------------------
def criteo_fields(p=1_000_000, F=39):
    return [p // F] * F


def c1_fields():
    return [1000] * 10


def movielens_fields():
    return [138_493, 26_744, 2_048]


def make_dataset(shape, n, seed=20240601, value_mode=None, task=None, p=None):
    """small host datasets for the parity tests"""
    if shape == "criteo":
        fs, sk, vm, tk = criteo_fields(p or 1_000_000), None, 0, "classification"
    elif shape == "c1":
        fs, sk, vm, tk = c1_fields(), None, 1, "regression"
    elif shape == "movielens":
        fs, sk, vm, tk = movielens_fields(), [0, 1, 0], 0, "rating"
    else:
        raise ValueError(shape)
    if value_mode is not None:
        vm = value_mode
    if task is not None:
        tk = task
    rowptr, col, val, pp = fields_csr(n, fs, sk, vm, seed)
    score = planted_scores_fast(rowptr, col, val, pp, seed=seed + 1)
    y = labels_from_scores(score, tk, noise=0.3 if tk == "rating" else 0.1, seed=seed + 3)
    return dict(n=n, p=pp, rowptr=rowptr, col=col, val=val, y=y, fields=fs)


def random_csr(n, p, max_nnz, seed=0, empty_rows=False, real_values=True):
    """ragged rows with ascending unique columns (general CSR for edge-case tests)"""
    rng = np.random.default_rng(seed)
    rowptr = [0]
    cols, vals = [], []
    for i in range(n):
        m = int(rng.integers(0 if empty_rows else 1, max_nnz + 1))
        m = min(m, p)
        c = np.sort(rng.choice(p, size=m, replace=False))
        cols.append(c)
        vals.append(rng.uniform(-1.5, 1.5, m) if real_values else np.ones(m))
        rowptr.append(rowptr[-1] + m)
    col = np.concatenate(cols).astype(np.uint32) if cols else np.zeros(0, np.uint32)
    val = np.concatenate(vals).astype(np.float32) if vals else np.zeros(0, np.float32)
    return np.array(rowptr, np.uint32), col, val
