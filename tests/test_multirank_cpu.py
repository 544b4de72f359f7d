"""world_size-2 gloo test (CPU) of the N>1 host logic: chunks are dealt round-robin to
ranks (rank r owns chunks c with (c-1) % n_ranks == r), Philox streams depend only on
the global packet id, and one all-reduce(sum) of the packed tallies reproduces the
single-rank result.  The packet physics here is the CPU oracle (test infrastructure);
what is under test is the partition + reduction contract shared with the CUDA path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KW = dict(letape_th=0, lmono=1, lambda_in=5, p_lambda_in=5, n_photons2=10 ** 9, n_phot_lim=20.0)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import small_problems
    from oracle.binding import Oracle
    P = small_problems()["cyl2D"]()
    t = Oracle(P).run(n_threads=1, rank=rank, n_ranks=world, **KW)
    packed = torch.from_numpy(np.concatenate([t.sed.ravel(order="F"), t.n_phot_sed.ravel(order="F"), t.n_phot_envoyes, t.stats]))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, packed.numpy())
    dist.destroy_process_group()


def test_two_ranks_sum_to_one_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import small_problems
    from oracle.binding import Oracle
    out = str(tmp_path / "packed.npy")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    got = np.load(out)
    P = small_problems()["cyl2D"]()
    t = Oracle(P).run(n_threads=1, **KW)
    ref = np.concatenate([t.sed.ravel(order="F"), t.n_phot_sed.ravel(order="F"), t.n_phot_envoyes, t.stats])
    assert got[-12] == ref[-12] == 128 * 20         # packets (stats[0] of 12)
    assert np.allclose(got, ref, rtol=1e-12, atol=0)
