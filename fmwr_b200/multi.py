"""Multi-GPU plans (one process per GPU).  Pure host logic: how the hot path is partitioned.

  predict.FM                 rows are sharded, parameters replicated, no collective
  SGD / FTRL / TDAP minibatch FEATURE-parallel: rank g owns a contiguous, field-aligned range of features
                             (its parameters and optimizer state) and the column slice of X for ALL rows;
                             per minibatch the per-row partials (S_f, linear term - 1/2 sum Q) are summed with one
                             NCCL all-reduce of rows x (kp + 4) reals, every rank derives the same multipliers
                             and updates only its own coordinates (csrc/train_minibatch.cu).
  exact (batch = 1) mode      inherently serial: single GPU
  ALS / MCMC                  replicas only in this round
"""
import numpy as np


def row_partition(n, world):
    """contiguous row ranges, sizes differing by at most one"""
    base, rem = divmod(int(n), int(world))
    out, r = [], 0
    for g in range(world):
        m = base + (1 if g < rem else 0)
        out.append((r, r + m))
        r += m
    return out


def field_partition(field_size, world):
    """contiguous field ranges balanced by field COUNT (every row has one non-zero per field, so work ~ fields);
    returns [(field_begin, field_end, col_begin, col_end)] per rank.  Ranks beyond the field count get empty slices."""
    F = len(field_size)
    offs = np.concatenate([[0], np.cumsum(np.asarray(field_size, np.int64))])
    out = []
    for f0, f1 in row_partition(F, world):
        out.append((f0, f1, int(offs[f0]), int(offs[f1])))
    return out


def gather_model(parts):
    """reassemble (w0, w, V) from the per-rank slices [(w0, w_slice, v_slice)] (w0 is replicated)"""
    w0 = parts[0][0]
    w = np.concatenate([p[1] for p in parts])
    v = np.concatenate([p[2] for p in parts], axis=0)
    return w0, w, v


def exchange_unique_id(dist, ctx_cls, rank):
    """rank 0 creates the NCCL id, everyone receives it through the host's process group (gloo or nccl)"""
    import torch
    if rank == 0:
        raw = ctx_cls.comm_unique_id()
        t = torch.tensor(list(raw), dtype=torch.uint8)
    else:
        t = torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(t, src=0)
    return bytes(t.tolist())


def open_peer_windows(dist, ctx, rank, world, batch_size, k):
    """map every rank's exchange window into every other rank (CUDA IPC over NVLink): afterwards the feature-parallel
    minibatch path exchanges its per-row partials inside its own kernels instead of calling NCCL (csrc/train_minibatch.cu)"""
    import torch
    nbytes = ctx.comm_peer_bytes(batch_size, k, world)
    mine = ctx.comm_peer_alloc(nbytes)
    t = torch.tensor(list(mine), dtype=torch.uint8)
    allh = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(allh, t)
    ctx.comm_peer_open([bytes(h.tolist()) for h in allh])
    dist.barrier()          # every window is zeroed and mapped before anyone stores into it
