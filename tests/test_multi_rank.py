"""world_size-2 gloo test of the multi-GPU host logic: frames shard across ranks with no data-path collective;
the only collective is the counter/timing reduction.  CPU only: each rank checks its shard of the golden corpus with the oracle."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import _pkg
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = _pkg.load().sharding if hasattr(_pkg.load(), "sharding") else None
    from importlib import import_module
    sh = import_module("zstd_rs_b200.sharding")
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    names = sorted(man["corpus"])
    sizes = [man["corpus"][n]["compressed_size"] for n in names]
    lo, hi = sh.shard_frames(sizes, world, rank)
    dec_bytes = 0
    for n in names[lo:hi]:
        data = open(os.path.join(ROOT, "tests", "golden", "decodecorpus", n), "rb").read()
        out, _ = O.decode_frame(data)
        assert hashlib.sha256(out).hexdigest() == man["corpus"][n]["sha256"]
        dec_bytes += len(out)
    tot = sh.gather_counters({"frames": hi - lo, "C": sum(sizes[lo:hi]), "D": dec_bytes, "max_ms": 10.0 + rank}, world, dist)
    q.put((rank, lo, hi, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_counter_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    (r0, lo0, hi0, t0), (r1, lo1, hi1, t1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == len(man["corpus"])          # disjoint, contiguous, complete
    assert t0 == t1                                                         # every rank sees the same totals
    assert t0["frames"] == 101 and t0["D"] == sum(m["size"] for m in man["corpus"].values())
    assert t0["C"] == sum(m["compressed_size"] for m in man["corpus"].values()) and t0["max_ms"] == 11.0
    c0 = sum(sorted(man["corpus"].items())[i][1]["compressed_size"] for i in range(lo0, hi0))
    assert 0.3 < c0 / t0["C"] < 0.7                                         # balanced by compressed bytes


def test_shard_frames_properties():
    sys.path.insert(0, ROOT)
    import _pkg
    _pkg.load()
    from importlib import import_module
    sh = import_module("zstd_rs_b200.sharding")
    rng = np.random.Generator(np.random.PCG64(5))
    for n in (0, 1, 7, 1000):
        sizes = rng.integers(1, 10000, n)
        for w in (1, 2, 3, 8):
            cuts = [sh.shard_frames(sizes, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            if n >= 100:
                per = [int(sizes[a:b].sum()) for a, b in cuts]
                assert max(per) - min(per) <= 2 * int(sizes.max())
