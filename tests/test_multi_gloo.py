"""Host-side multi-GPU logic on CPU: partition plans and the communicator-id exchange over a world_size-2 gloo group."""
import os
import subprocess
import sys

import numpy as np

from fmwr_b200 import multi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition_covers_everything():
    for n, w in ((10, 3), (7, 8), (10_000_000, 8), (0, 2), (39, 8)):
        parts = multi.row_partition(n, w)
        assert len(parts) == w and parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_field_partition_is_field_aligned():
    fs = [25641] * 39
    parts = multi.field_partition(fs, 8)
    assert [b - a for a, b, _, _ in parts] == [5, 5, 5, 5, 5, 5, 5, 4]
    assert parts[0][2] == 0 and parts[-1][3] == sum(fs)
    for f0, f1, c0, c1 in parts:
        assert c0 == sum(fs[:f0]) and c1 == sum(fs[:f1])
    parts = multi.field_partition([3, 5], 4)              # more ranks than fields: empty slices at the end
    assert [(a, b) for a, b, _, _ in parts] == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_gather_model_roundtrip():
    rng = np.random.default_rng(0)
    w = rng.normal(size=20); v = rng.normal(size=(20, 3))
    cuts = [(0, 7), (7, 12), (12, 20)]
    parts = [(0.5, w[a:b], v[a:b]) for a, b in cuts]
    w0, w2, v2 = multi.gather_model(parts)
    assert w0 == 0.5 and np.array_equal(w2, w) and np.array_equal(v2, v)


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from fmwr_b200 import multi
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
class FakeCtx:
    @staticmethod
    def comm_unique_id():
        return bytes((7 * i + 3) %% 256 for i in range(128))
got = multi.exchange_unique_id(dist, FakeCtx, rank)
assert got == FakeCtx.comm_unique_id(), got
# max-over-ranks timing reduction used by bench.py
t = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert float(t[0]) == 10.0 + world - 1
parts = [None] * world
dist.all_gather_object(parts, (0.25, [rank], [[rank]]))
assert [p[1][0] for p in parts] == list(range(world))
dist.barrier(); dist.destroy_process_group()
print("GLOO_OK", rank)
'''


def test_id_exchange_over_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.count("GLOO_OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]
