"""CPU, world_size 2 over gloo: chains shard over ranks with no data-path collective and the per-chain
traces are gathered on rank 0; the result must equal the same chains run in one process."""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from tests.chain_utils import load_chain, signature


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _data():
    from oracle import numerics as onum
    from phyclone_b200.data.base import DataPoint

    vals = load_chain("syn6_boot")["values"]
    return [DataPoint(i, vals[i], outlier_marginal_prob=onum.outlier_marginal(vals[i])) for i in range(len(vals))]


def _run(rank, world, num_chains):
    from phyclone_b200.run import run_chains
    from tests.fake_store import FakeTileStore

    return run_chains(_data(), ["s"], num_chains=num_chains, seed=11, burnin=1, num_iters=3, num_particles=6,
                      rank=rank, world_size=world, store_factory=FakeTileStore)


def _worker(rank, world, port, num_chains, out_path):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _run(rank, world, num_chains)
    if rank == 0:
        n = len(_data())
        summary = {c: [(e["alpha"], e["log_p_one"], signature(e["tree"], n)[1].tolist()) for e in r["trace"]]
                   for c, r in res.items()}
        np.save(out_path, np.array([summary], dtype=object), allow_pickle=True)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_chain_assignment():
    from phyclone_b200.run import chain_generators, chains_of_rank

    assert chains_of_rank(8, 0, 2) == [0, 2, 4, 6] and chains_of_rank(8, 1, 2) == [1, 3, 5, 7]
    assert sorted(sum((chains_of_rank(5, r, 4) for r in range(4)), [])) == list(range(5))
    # one chain uses the main Generator, several chains its spawned children (reference run.py:87-113)
    assert chain_generators(3, 1)[0].random() == np.random.default_rng(3).random()
    a = [g.random() for g in chain_generators(3, 4)]
    b = [g.random() for g in np.random.default_rng(3).spawn(4)]
    assert a == b


@pytest.mark.timeout(600)
def test_two_ranks_match_single_process(tmp_path):
    num_chains = 3
    single = _run(0, 1, num_chains)
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), num_chains, out), nprocs=2, join=True)
    gathered = np.load(out, allow_pickle=True)[0]
    assert sorted(gathered) == list(range(num_chains))
    n = len(_data())
    for c in range(num_chains):
        ref = [(e["alpha"], e["log_p_one"], signature(e["tree"], n)[1].tolist()) for e in single[c]["trace"]]
        assert gathered[c] == ref, "chain %d differs between 1 and 2 ranks" % c
