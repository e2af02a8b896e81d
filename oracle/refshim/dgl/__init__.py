"""Minimal stand-in for the DGL graph container (TEST INFRASTRUCTURE ONLY).

The reference (Dunni3/FlowMol) needs `dgl==2.0.0`, which is not installed in
this image.  To execute the reference's own modules *verbatim* (see
oracle/ref_loader.py) this package supplies only the container API the
sampling path touches:

  dgl.graph, dgl.batch, dgl.unbatch, dgl.readout_nodes, DGLGraph.{ndata, edata,
  local_scope, apply_edges, update_all, edges, num_nodes, num_edges, batch_size,
  batch_num_nodes, batch_num_edges, device, to}

The four arithmetic ops DGL would do natively are restated with documented
DGL semantics: u_sub_v = x[src] - x[dst]; copy_e+sum = index_add over dst;
copy_e+mean; readout mean = per-graph mean.  Nothing in the product path
(flowmol_b200/) imports this.
"""
import contextlib
import torch
from . import function  # noqa: F401


class _Frame(dict):
    pass


class _EdgeBatch:
    def __init__(self, g):
        s, d = g._src, g._dst
        self.src = {k: v[s] for k, v in g.ndata.items()}
        self.dst = {k: v[d] for k, v in g.ndata.items()}
        self.data = g.edata


class DGLGraph:
    def __init__(self, src, dst, num_nodes, batch_num_nodes=None, batch_num_edges=None):
        self._src = src.long()
        self._dst = dst.long()
        self._n = int(num_nodes)
        self.ndata = _Frame()
        self.edata = _Frame()
        dev = self._src.device
        self._bnn = batch_num_nodes if batch_num_nodes is not None else torch.tensor([self._n], device=dev)
        self._bne = batch_num_edges if batch_num_edges is not None else torch.tensor([self._src.shape[0]], device=dev)

    # --- structure -------------------------------------------------------
    @property
    def device(self):
        return self._src.device

    @property
    def batch_size(self):
        return int(self._bnn.shape[0])

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.shape[0])

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    def edges(self):
        return self._src, self._dst

    def to(self, device):
        g = DGLGraph(self._src.to(device), self._dst.to(device), self._n,
                     self._bnn.to(device), self._bne.to(device))
        for k, v in self.ndata.items():
            g.ndata[k] = v.to(device)
        for k, v in self.edata.items():
            g.edata[k] = v.to(device)
        return g

    def remove_nodes(self, nids):
        """DGLGraph.remove_nodes: drops the nodes and their incident edges, relabels the remaining nodes in order, keeps the
        feature rows of what remains (molecule_builder.py:231 uses it on one unbatched molecule)."""
        nids = torch.as_tensor(nids).long()
        keep = torch.ones(self._n, dtype=torch.bool, device=self._src.device)
        keep[nids] = False
        new_id = torch.cumsum(keep, 0) - 1
        ek = keep[self._src] & keep[self._dst]
        self._src, self._dst = new_id[self._src[ek]], new_id[self._dst[ek]]
        self.ndata = _Frame({k: v[keep] for k, v in self.ndata.items()})
        self.edata = _Frame({k: v[ek] for k, v in self.edata.items()})
        self._n = int(keep.sum())
        self._bnn = torch.tensor([self._n], device=self._src.device)
        self._bne = torch.tensor([self._src.shape[0]], device=self._src.device)

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = _Frame(self.ndata), _Frame(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    # --- message passing ---------------------------------------------------
    def apply_edges(self, func):
        if isinstance(func, function._USubV):
            self.edata[func.out] = self.ndata[func.lhs][self._src] - self.ndata[func.rhs][self._dst]
        else:
            out = func(_EdgeBatch(self))
            for k, v in out.items():
                self.edata[k] = v

    def update_all(self, msg, red):
        assert isinstance(msg, function._CopyE)
        m = self.edata[msg.e]
        out = torch.zeros((self._n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        out.index_add_(0, self._dst, m)
        if red.op == 'mean':
            deg = torch.zeros(self._n, dtype=m.dtype, device=m.device)
            deg.index_add_(0, self._dst, torch.ones_like(self._dst, dtype=m.dtype))
            out = out / deg.clamp(min=1).view((-1,) + (1,) * (m.dim() - 1))
        self.ndata[red.out] = out


def graph(data, num_nodes=None, device=None):
    u, v = data
    u = torch.as_tensor(u)
    v = torch.as_tensor(v)
    if device is not None:
        u, v = u.to(device), v.to(device)
    return DGLGraph(u, v, int(num_nodes))


def batch(graphs):
    srcs, dsts, off = [], [], 0
    for g in graphs:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        off += g._n
    dev = graphs[0].device
    bnn = torch.tensor([g._n for g in graphs], device=dev)
    bne = torch.tensor([g.num_edges() for g in graphs], device=dev)
    out = DGLGraph(torch.cat(srcs), torch.cat(dsts), off, bnn, bne)
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs])
    for k in graphs[0].edata:
        out.edata[k] = torch.cat([g.edata[k] for g in graphs])
    return out


def unbatch(g):
    outs, no, eo = [], 0, 0
    for n, e in zip(g._bnn.tolist(), g._bne.tolist()):
        gi = DGLGraph(g._src[eo:eo + e] - no, g._dst[eo:eo + e] - no, n)
        for k, v in g.ndata.items():
            gi.ndata[k] = v[no:no + n]
        for k, v in g.edata.items():
            gi.edata[k] = v[eo:eo + e]
        outs.append(gi)
        no += n
        eo += e
    return outs


def readout_nodes(g, feat, op='sum'):
    x = g.ndata[feat]
    B = g.batch_size
    idx = torch.arange(B, device=x.device).repeat_interleave(g._bnn)
    out = torch.zeros((B,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    out.index_add_(0, idx, x)
    if op == 'mean':
        out = out / g._bnn.to(x.dtype).view((-1,) + (1,) * (x.dim() - 1))
    return out
