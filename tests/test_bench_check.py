"""bench.py's correctness check (`check` in its JSON line) compares the N-rank product run with
tests/golden/bench_check_mtr_cg4.npz, the 1-rank oracle run of tests/golden/make_bench_check.py.
Here, without a GPU: the golden is what the oracle produces, and an N-slice run of the same case
stays within the summation-order noise of it — so a `check` above 1e-5 on the GPU means the
device path, not the decomposition."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_golden_prefix_is_reproduced_by_a_four_slice_oracle_run():
    import bench
    from axisem_b200.capi import TimeLoop, connect_local, run_group
    from oracle import oracle
    from tests.util import rel_l2
    world = 4
    lib = oracle.load()
    probs, z = [], None
    for r in range(world):
        p, z = bench.check_problem(r, world)
        probs.append(p)
    assert abs(probs[0].deltat - float(z["deltat"])) < 1e-15
    nsteps = 1200                                   # a prefix of the 9600-step golden run
    loops = [TimeLoop(lib, p) for p in probs]
    connect_local(lib, loops)
    run_group(lib, loops, nsteps)
    ref = z["seismograms"]
    ns = nsteps // int(z["seis_it"]) + 1
    got = np.zeros((ns,) + ref.shape[1:], np.float32)
    seen = 0
    for p, L in zip(probs, loops):
        if p.num_rec:
            got[:, p.rec_index, :] = L.seismograms()
            seen += p.num_rec
    assert seen == ref.shape[1]
    assert np.abs(ref[:ns]).max() > 0
    assert rel_l2(got, ref[:ns]) <= 1e-5, rel_l2(got, ref[:ns])
