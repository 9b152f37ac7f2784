"""N > 1 path on CPU (not gpu): one process per theta-slice, wired exactly like the GPU job
(axisem_b200.dist.connect_ranks over a gloo group), running the oracle's one-process-per-
rank mode (POSIX shared-memory "MPI").  world_size 2 and 4; rendezvous on 127.0.0.1."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, src, anel, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from axisem_b200.dist import connect_ranks, neighbours
    from oracle import oracle
    from tests.util import make_problem
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = make_problem(src, anel=anel, ntheta=8, nr=12, niter=n, t_0=8.0, rank=rank, nranks=world)
        assert neighbours(prob) == [r for r in (rank - 1, rank + 1) if 0 <= r < world]
        loop = oracle.make_loop(prob)
        connect_ranks(loop, rank, world)
        loop.run(n // 2)
        loop.run(n - n // 2)                # a second call continues the same exchange sequence
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), seis=loop.seismograms(), idx=prob.rec_index)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,src,anel", [(2, "mtr", True), (2, "explosion", False), (4, "mtp", False)])
def test_one_process_per_slice_matches_single_rank(world, src, anel, tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle
    from tests.util import make_problem, rel_l2
    n = 40
    oracle.build()
    mp.spawn(_worker, args=(world, _free_port(), src, anel, n, str(tmp_path)), nprocs=world, join=True)
    one = oracle.make_loop(make_problem(src, anel=anel, ntheta=8, nr=12, niter=n, t_0=8.0))
    one.run(n)
    ref = one.seismograms()
    got = np.zeros_like(ref)
    seen = np.zeros(ref.shape[1], bool)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        got[:, z["idx"]] = z["seis"]
        seen[z["idx"]] = True
    assert seen.all() and np.abs(ref).max() > 0
    assert rel_l2(got, ref) < 1e-5
