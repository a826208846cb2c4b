"""Multi-GPU host logic on CPU: world_size-2 (and 3, ragged) gloo process groups exercise the shard arithmetic and the
path's single collective (all-gather of per-instance results).  The per-rank solve is replaced by the oracle here
(no GPU); on the box the same functions run under torchrun with NCCL (bench.py --gpus N)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding = importlib.import_module("cddp-cpp_b200.sharding")
    problems = importlib.import_module("cddp-cpp_b200.problems")
    import oracle_binding as ob
    cfg = problems.make_config("unicycle", batch=batch, horizon=30)
    lo, hi = sharding.shard_bounds(batch, rank, world)
    sh = sharding.shard_arrays({k: cfg[k] for k in ("x0", "xref", "X0", "U0", "ref_traj")}, rank, world)
    assert sh["x0"].shape[0] == hi - lo
    P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**dict(cfg["options"], max_iterations=6))
    if hi > lo:
        r = ob.solve_batch(P, oo, sh["x0"], sh["xref"], sh["X0"], sh["U0"])
        cost, it, st = r["cost"], r["iterations"], r["status"]
    else:
        cost, it, st = np.zeros(0), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)
    gc, gi, gs = sharding.gather_results(cost, it, st, batch)
    q.put((rank, gc, gi, gs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 10), (3, 7), (2, 1)])
def test_shard_and_gather_gloo(world, batch):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_binding as ob
    problems = importlib.import_module("cddp-cpp_b200.problems")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = problems.make_config("unicycle", batch=batch, horizon=30)
    ref = ob.solve_batch(ob.OracleProblem(cfg["spec"]), ob.make_options(**dict(cfg["options"], max_iterations=6)), cfg["x0"], cfg["xref"],
                         cfg["X0"], cfg["U0"])
    for rank, gc, gi, gs in got:  # every rank holds the whole batch's results, in instance order
        np.testing.assert_array_equal(gc, ref["cost"])
        np.testing.assert_array_equal(gi, ref["iterations"])
        np.testing.assert_array_equal(gs, ref["status"])


def test_shard_bounds_cover_the_batch_exactly():
    sharding = importlib.import_module("cddp-cpp_b200.sharding")
    for B in (1, 7, 8, 4096, 8191):
        for G in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) == -(-B // G)
