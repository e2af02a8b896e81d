"""Batched complete molecular graphs: the reference's graph contract, host side.

The reference builds one DGL graph per molecule with `build_edge_idxs(n)` (all n(n-1) directed pairs: upper triangle
row-major, then the same list with src/dst swapped, flowmol/data_processing/utils.py:4-17), batches them with
`dgl.batch` (flowmol/models/flowmol.py:509-523) and derives `upper_edge_mask`, `node_batch_idx`, `edge_batch_idx`
(utils.py:19-46).  Because the edge list is a closed form of the atom counts, the CUDA library never sees it: it takes
`n_atoms[B]` and works in its own dst-major internal order.  This module provides

  * `MolGraphBatch` -- a light container with the slice of the DGLGraph API the sampling path touches
    (`ndata`, `edata`, `batch_size`, `batch_num_nodes()`, `batch_num_edges()`, `num_nodes()`, `num_edges()`, `edges()`,
    `device`, `to()`), so code written against the reference's `integrate(g, ...)` works with or without DGL installed;
  * index maps between the reference edge order and the library's internal padded order (used by tests and by
    anything that wants to look at per-edge hidden state).
"""
import numpy as np
import torch

TM = 64  # tile rows of the CUDA kernels (csrc/common.cuh)


def reference_edges(n):
    """(src, dst) int64 arrays of one molecule in the reference order (utils.py:4-17)."""
    iu = np.triu_indices(n, k=1)
    return np.concatenate([iu[0], iu[1]]), np.concatenate([iu[1], iu[0]])


def internal_edge_pos(i, j, n):
    """Local slot of directed edge src i -> dst j in the library's dst-major order (csrc/common.cuh:edge_pos)."""
    return j * (n - 1) + np.where(i < j, i, i - 1)


def batch_sizes(n_atoms):
    n = np.asarray(n_atoms, dtype=np.int64)
    return dict(B=len(n), N=int(n.sum()), E=int((n * (n - 1)).sum()), U=int((n * (n - 1) // 2).sum()),
                EP=int((((n * (n - 1)) + TM - 1) // TM * TM).sum()))


def ref_edge_to_internal(n_atoms):
    """int64 [E]: for every directed edge in the reference's batched order, its slot in the internal padded layout."""
    out, ebase = [], 0
    for n in np.asarray(n_atoms, dtype=np.int64):
        src, dst = reference_edges(int(n))
        out.append(ebase + internal_edge_pos(src, dst, int(n)))
        ebase += (n * (n - 1) + TM - 1) // TM * TM
    return np.concatenate(out)


class MolGraphBatch:
    """Batched complete graphs with reference-ordered node / edge data dictionaries."""

    def __init__(self, n_atoms, device="cpu"):
        self.n_atoms = torch.as_tensor(n_atoms, dtype=torch.int64).reshape(-1).cpu()
        if (self.n_atoms < 2).any():
            raise ValueError("every molecule needs at least 2 atoms")
        self._device = torch.device(device)
        self.ndata = {}
        self.edata = {}
        self._edges = None

    # -- DGLGraph API subset --------------------------------------------------------------------------------------
    @property
    def device(self):
        return self._device

    @property
    def batch_size(self):
        return int(self.n_atoms.shape[0])

    def batch_num_nodes(self):
        return self.n_atoms.to(self._device)

    def batch_num_edges(self):
        return (self.n_atoms * (self.n_atoms - 1)).to(self._device)

    def num_nodes(self):
        return int(self.n_atoms.sum())

    def num_edges(self):
        return int((self.n_atoms * (self.n_atoms - 1)).sum())

    def edges(self):
        if self._edges is None:
            srcs, dsts, off = [], [], 0
            for n in self.n_atoms.tolist():
                s, d = reference_edges(n)
                srcs.append(s + off)
                dsts.append(d + off)
                off += n
            self._edges = (torch.from_numpy(np.concatenate(srcs)).to(self._device),
                           torch.from_numpy(np.concatenate(dsts)).to(self._device))
        return self._edges

    def to(self, device):
        g = MolGraphBatch(self.n_atoms, device)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device) for k, v in self.edata.items()}
        return g

    # -- helpers mirroring flowmol/data_processing/utils.py:19-46 -------------------------------------------------------
    def upper_edge_mask(self):
        u = self.n_atoms * (self.n_atoms - 1) // 2
        pattern = torch.tensor([1, 0]).repeat(self.batch_size)
        return pattern.repeat_interleave(u.repeat_interleave(2)).bool().to(self._device)

    def node_batch_idx(self):
        return torch.arange(self.batch_size).repeat_interleave(self.n_atoms).to(self._device)

    def edge_batch_idx(self):
        return torch.arange(self.batch_size).repeat_interleave(self.n_atoms * (self.n_atoms - 1)).to(self._device)

    def unbatch(self):
        """Per-molecule views (dgl.unbatch analogue) as plain dicts of CPU tensors."""
        outs, no, eo = [], 0, 0
        for n in self.n_atoms.tolist():
            e = n * (n - 1)
            outs.append({"n_atoms": n,
                         "ndata": {k: v[no:no + n] for k, v in self.ndata.items()},
                         "edata": {k: v[eo:eo + e] for k, v in self.edata.items()}})
            no += n
            eo += e
        return outs


def n_atoms_of(g):
    """Atom counts of a DGLGraph / MolGraphBatch after checking it honours the complete-graph contract."""
    n = torch.as_tensor(g.batch_num_nodes()).detach().cpu().to(torch.int64)
    e = torch.as_tensor(g.batch_num_edges()).detach().cpu().to(torch.int64)
    if not torch.equal(e, n * (n - 1)):
        raise ValueError("flowmol_b200 needs complete molecular graphs with n(n-1) directed edges per molecule "
                         "(flowmol/data_processing/utils.py:4-17)")
    if not isinstance(g, MolGraphBatch):
        # a foreign graph object: verify the reference edge order on the first and the last molecule
        src, dst = g.edges()
        src, dst = src.detach().cpu(), dst.detach().cpu()
        for b in {0, len(n) - 1}:
            nb = int(n[b])
            no, eo = int(n[:b].sum()), int(e[:b].sum())
            s, d = reference_edges(nb)
            if not (np.array_equal(src[eo:eo + nb * (nb - 1)].numpy() - no, s)
                    and np.array_equal(dst[eo:eo + nb * (nb - 1)].numpy() - no, d)):
                raise ValueError("edge order differs from build_edge_idxs (upper triangle first, then its mirror)")
    return n
