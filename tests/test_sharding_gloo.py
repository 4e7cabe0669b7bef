"""Multi-GPU host logic on CPU: environments are block-partitioned across ranks with no data-path collective; the only
exchange is the validation gather of sfh_info records + checksums (NCCL on the GPU box, gloo here, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from stonefish_b200 import INFO_DTYPE
from stonefish_b200.sharding import checksum, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _fake_infos(lo, hi):
    """Deterministic stand-in for per-environment results (function of the environment id only)."""
    a = np.zeros(hi - lo, dtype=INFO_DTYPE)
    ids = np.arange(lo, hi)
    a["Fdq"] = np.stack([np.sin(ids), np.cos(ids), ids * 1e-3], axis=1)
    a["Swet"] = ids + 0.5
    a["n_faces"] = 50000
    a["class_count"][:, 7] = 50000
    return a


def test_checksum_is_order_independent_and_sensitive():
    a = _fake_infos(0, 100)
    assert checksum(a) == checksum(a[::-1].copy()) == checksum(np.concatenate([a[50:], a[:50]]))
    b = a.copy()
    b["Fdq"][17, 1] = np.nextafter(b["Fdq"][17, 1], 1.0)
    assert checksum(a) != checksum(b)
    assert checksum(a[:0]) == 0


def _worker(rank, world, port, n_env, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from stonefish_b200.sharding import gather_infos

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n_env, rank, world)
    local = _fake_infos(lo, hi)
    allrec, sums = gather_infos(local, device="cpu")
    dist.barrier()
    if rank == 0:
        q.put((allrec.tobytes(), sums))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_env", [10, 11, 1])
def test_gather_over_gloo_world_size_2(n_env):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_env, q)) for r in range(2)]
    for p in procs:
        p.start()
    raw, sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.frombuffer(raw, dtype=INFO_DTYPE)
    want = _fake_infos(0, n_env)
    assert got.tobytes() == want.tobytes()  # rank order == environment order
    parts = [shard_range(n_env, r, 2) for r in range(2)]
    assert sums == [checksum(want[lo:hi]) & 0x7FFFFFFFFFFFFFFF for lo, hi in parts]
