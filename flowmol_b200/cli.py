"""Sampling command line: the caller of the hot path whose wall clock the reference's tooling records (test.py:16-133,215-262).

    python -m flowmol_b200.cli --model_dir <dir> | --checkpoint <ckpt> | --config flowmol3 [--dataset geom]
        [--n_mols 100] [--n_atoms_per_mol N] [--n_timesteps 250] [--max_batch_size 128] [--max_batch_edges E]
        [--xt_traj] [--ep_traj] [--stochasticity S] [--hc_thresh H] [--seed SEED] [--output_file out.sdf]

Same arguments and meaning as the reference's test.py for the sampling part (metrics, pickles of rdkit objects and the
baseline-comparison format are out of scope: SURVEY.md section 2).  Differences that come with the B200 design:

* batches are planned by COST, not only by count: molecules are sorted by size and packed under `--max_batch_size` molecules AND
  `--max_batch_edges` directed edges (the kernels' work and workspace scale with sum n(n-1)), so a batch of large molecules
  does not blow the workspace and small molecules are not sampled in half-empty launches;
* under torchrun every rank samples its cost-balanced share of every batch (flowmol_b200/sharding.py: no collective inside a
  timestep), rank 0 gathers and writes;
* the SDF is written from the decoded arrays directly (V2000 mol blocks), so the tool works without rdkit; when rdkit is
  importable the molecules also carry `.rdkit_mol` exactly as in the reference.
"""
import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="FlowMol sampling on B200 (flowmol_b200)")
    p.add_argument('--model_dir', type=Path, default=None, help='model directory (uses <model_dir>/checkpoints/last.ckpt)')
    p.add_argument('--checkpoint', type=Path, default=None, help='Lightning checkpoint file')
    p.add_argument('--config', type=str, default=None, help="random-init weights of a named reference config ('flowmol3', 'dev')")
    p.add_argument('--dataset', type=str, default='geom', choices=['geom', 'qm9'], help='size histogram / atom map for --config')
    p.add_argument('--output_file', type=Path, default=None)
    p.add_argument('--n_mols', type=int, default=100)
    p.add_argument('--n_atoms_per_mol', type=int, default=None)
    p.add_argument('--n_timesteps', type=int, default=250)
    p.add_argument('--xt_traj', action='store_true')
    p.add_argument('--ep_traj', action='store_true')
    p.add_argument('--max_batch_size', type=int, default=128)
    p.add_argument('--max_batch_edges', type=int, default=4_000_000, help='directed edges per batch and GPU (workspace bound)')
    p.add_argument('--stochasticity', type=float, default=None)
    p.add_argument('--hc_thresh', type=float, default=None)
    p.add_argument('--seed', type=int, default=None)
    args = p.parse_args(argv)
    if sum(x is not None for x in (args.model_dir, args.checkpoint, args.config)) != 1:
        p.error('exactly one of --model_dir, --checkpoint, --config is required')      # test.py:47-51
    return args


def plan_batches(n_atoms, max_batch_size, max_batch_edges):
    """Size-aware packing: -> list of index arrays into `n_atoms`.  Molecules are taken in order of decreasing size (similar
    sizes share a batch, so the per-step cost of a batch is not dominated by one outlier) and a batch is closed when it holds
    `max_batch_size` molecules or adding the next one would exceed `max_batch_edges` directed edges.  Every molecule lands in
    exactly one batch; a single molecule larger than the edge budget still gets its own batch."""
    n = np.asarray(n_atoms, dtype=np.int64)
    order = np.argsort(-n, kind="stable")
    batches, cur, edges = [], [], 0
    for i in order:
        e = int(n[i] * (n[i] - 1))
        if cur and (len(cur) >= max_batch_size or edges + e > max_batch_edges):
            batches.append(np.asarray(cur, dtype=np.int64))
            cur, edges = [], 0
        cur.append(int(i))
        edges += e
    if cur:
        batches.append(np.asarray(cur, dtype=np.int64))
    return batches


_BOND_ORDER_SDF = {1: 1, 2: 2, 3: 3, 4: 4}       # 4 = aromatic in a V2000 bond block


def mol_block(mol, name=""):
    """V2000 mol block of a SampledMolecule (decoded arrays only; charges in an 'M  CHG' property line)."""
    pos = mol.positions.numpy()
    lines = [name, "  flowmol_b200", "", f"{len(mol.atom_types):3d}{len(mol.bond_types):3d}  0  0  0  0  0  0  0  0999 V2000"]
    for sym, (x, y, z) in zip(mol.atom_types, pos):
        lines.append(f"{x:10.4f}{y:10.4f}{z:10.4f} {sym:<3s} 0  0  0  0  0  0  0  0  0  0  0  0")
    for s, d, b in zip(mol.bond_src_idxs.tolist(), mol.bond_dst_idxs.tolist(), mol.bond_types.tolist()):
        lines.append(f"{s + 1:3d}{d + 1:3d}{_BOND_ORDER_SDF[int(b)]:3d}  0")
    charged = [(i + 1, int(c)) for i, c in enumerate(mol.atom_charges.tolist()) if c != 0]
    for k in range(0, len(charged), 8):
        chunk = charged[k:k + 8]
        lines.append(f"M  CHG{len(chunk):3d}" + "".join(f"{i:4d}{c:4d}" for i, c in chunk))
    lines.append("M  END")
    return "\n".join(lines)


def write_sdf(path, mols):
    with open(path, "w") as f:
        for i, m in enumerate(mols):
            f.write(mol_block(m, f"mol_{i}") + "\n$$$$\n")


def main(argv=None):
    args = parse_args(argv)
    from . import api
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(device))
    # test.py:70-71 seeds everything only when --seed is given and draws fresh randomness otherwise.  Here every rank must draw the
    # same sizes and the same noise, so an unseeded run picks ONE random base seed on rank 0, broadcasts it and prints it.
    base_seed = args.seed
    if base_seed is None:
        base_seed = int.from_bytes(os.urandom(7), "little")
        if world > 1:
            import torch.distributed as dist
            box = [base_seed]
            dist.broadcast_object_list(box, src=0)
            base_seed = int(box[0])
        if rank == 0:
            print(f"no --seed given: using base seed {base_seed} (pass --seed {base_seed} to reproduce this run)")
    torch.manual_seed(base_seed)
    if args.config is not None:
        model = api.FlowMolB200.from_config(args.config, dataset=args.dataset, seed=0, device=device)
        out_dir = Path(".")
    else:
        ckpt = args.checkpoint if args.checkpoint is not None else args.model_dir / "checkpoints" / "last.ckpt"
        model = api.FlowMolB200.from_checkpoint(ckpt, device=device)
        out_dir = ckpt.parent.parent / "samples"
    model = model.cuda(local).eval()
    # all sizes up front (every rank draws the same list: same torch seed), then cost-aware batches, each sharded over the ranks
    if args.n_atoms_per_mol is None:
        g = torch.Generator().manual_seed(base_seed)
        n_atoms = model.n_atoms_map[torch.multinomial(model.n_atoms_dist.probs.float(), args.n_mols, replacement=True, generator=g)]
    else:
        n_atoms = torch.full((args.n_mols,), args.n_atoms_per_mol, dtype=torch.long)
    n_atoms = n_atoms.numpy()
    batches = plan_batches(n_atoms, args.max_batch_size * world, args.max_batch_edges * world)
    from . import sharding as SH
    molecules = [None] * args.n_mols
    # noise that does not depend on the GPU count: one generator (same state on every rank) draws, per batch, the Philox seed of
    # the CTMC noise and the prior positions of the WHOLE batch; a rank integrates its slice with mol_id_offset = its first molecule
    ngen = torch.Generator().manual_seed(base_seed + 0x5EED)
    torch.cuda.synchronize()
    start = time.time()
    for bi, idx in enumerate(batches):
        nb = n_atoms[idx]
        lo, hi = SH.partition(nb, world)[rank] if world > 1 else (0, len(nb))
        seed_b = int(torch.randint(0, 2 ** 62, (1,), generator=ngen).item())
        x0 = api.FlowMolB200.centered_normal(nb, ngen)
        noff = np.concatenate([[0], np.cumsum(nb)])
        mine = []
        if hi > lo:
            prior = model.prior_from_x0(nb[lo:hi], x0[noff[lo]:noff[hi]])
            mine = model.sample(torch.from_numpy(nb[lo:hi]), n_timesteps=args.n_timesteps, device=device, prior=prior,
                                stochasticity=args.stochasticity, high_confidence_threshold=args.hc_thresh,
                                xt_traj=args.xt_traj, ep_traj=args.ep_traj, seed=seed_b, mol_id_offset=lo)
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world if rank == 0 else None
            dist.gather_object([(int(idx[lo + k]), m) for k, m in enumerate(mine)], gathered, dst=0)
            if rank == 0:
                for part in gathered:
                    for gi, m in part:
                        molecules[gi] = m
        else:
            for k, m in enumerate(mine):
                molecules[int(idx[k])] = m
    torch.cuda.synchronize()
    sampling_time = time.time() - start
    if rank == 0:
        print(f"sampled {args.n_mols} molecules in {len(batches)} batches on {world} GPU(s): {sampling_time:.2f} s "
              f"({args.n_mols / sampling_time:.2f} molecules/s, {args.n_timesteps} timesteps)")
        out = args.output_file or (out_dir / "sampled_mols.sdf")
        out.parent.mkdir(parents=True, exist_ok=True)
        if not (args.xt_traj or args.ep_traj):
            write_sdf(out, molecules)
            print(f"wrote {out}")
        else:                                                                # test.py:224-257: one file per molecule trajectory
            for i, m in enumerate(molecules):
                if args.xt_traj:
                    write_sdf(out.parent / f"{out.stem}_{i}_xt{out.suffix}", m.traj_mols)
                if args.ep_traj:
                    write_sdf(out.parent / f"{out.stem}_{i}_ep{out.suffix}", m.ep_traj_mols)
            print(f"wrote {len(molecules)} trajectories to {out.parent}")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
