"""CPU, world_size = 2 over gloo: the multi-GPU path's host logic -- contiguous cost-balanced molecule ranges, noise keyed by
global molecule id, one gather of results -- reproduces the single-process result bit for bit (CPU oracle stands in for the
per-rank CUDA sampler; the GPU version of the same property is tests/test_gpu_parity.py::test_results_do_not_depend_...)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowmol_b200 import sharding as SH
from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig
from oracle import flowmol_oracle as O

N_ATOMS = [5, 9, 3, 14, 4, 7, 6]
T, SEED = 5, 77


def _sample(cfg, sd, n_atoms, x0, mol_id_offset):
    bt = O.make_batch(n_atoms)
    with torch.no_grad():
        return O.integrate(O.OracleModel(cfg, sd), bt, x0, torch.full((bt.N,), 6), torch.full((bt.N,), 6), torch.full((bt.U,), 4),
                           T, seed=SEED, mol_id_offset=mol_id_offset)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = ModelConfig.named("dev", 6)
    sd = WT.init_state_dict(cfg, 2)
    n = np.array(N_ATOMS)
    x0 = torch.randn(int(n.sum()), 3, generator=torch.Generator().manual_seed(0))
    ranges = SH.partition(n, world)
    lo, hi = ranges[rank]
    noff = np.concatenate([[0], np.cumsum(n)])
    out = _sample(cfg, sd, n[lo:hi], x0[noff[lo]:noff[hi]], lo)
    res = SH.gather_results(out, n, ranges, rank, world)
    if rank == 0:
        q.put({k: v.numpy() for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_contiguous_balanced_and_complete():
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        n = rng.integers(3, 120, size=64)
        r = SH.partition(n, world)
        assert r[0][0] == 0 and r[-1][1] == len(n) and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert all(hi > lo for lo, hi in r)
        cost = [int((n[lo:hi] * (n[lo:hi] - 1)).sum()) for lo, hi in r]
        assert max(cost) <= 1.35 * (sum(cost) / world)
    assert SH.partition([5, 5], 2) == [(0, 1), (1, 2)]


def test_two_rank_gloo_run_equals_single_process_run():
    cfg = ModelConfig.named("dev", 6)
    sd = WT.init_state_dict(cfg, 2)
    n = np.array(N_ATOMS)
    x0 = torch.randn(int(n.sum()), 3, generator=torch.Generator().manual_seed(0))
    want = _sample(cfg, sd, n, x0, 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got["x"], want["x"].numpy())
    for k in "ace":
        assert np.array_equal(got[k], want[k].numpy().astype(np.uint8))
