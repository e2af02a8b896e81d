"""Stand-in for pytorch-scatter==2.1.2 `segment_csr` (sum) used at flowmol/utils/ctmc_utils.py:15,18 (test infrastructure only)."""
import torch


def segment_csr(src, indptr, reduce='sum'):
    assert reduce == 'sum'
    c = torch.zeros(src.shape[0] + 1, dtype=src.dtype, device=src.device)
    c[1:] = src.cumsum(0)
    return c[indptr[1:]] - c[indptr[:-1]]
