"""Counter-based RNG shared by the oracle and the CUDA path (TEST INFRASTRUCTURE side).

Philox4x32-10 (Salmon et al., SC'11; the same generator cuRAND/torch-CUDA use), restated in numpy.
The reference draws its per-step noise from the global torch RNG
(flowmol/models/ctmc_vector_field.py:428,445,450; flowmol/utils/ctmc_utils.py:34), which is not
reproducible across devices; "identical noise seeds" is therefore defined here as:

    counter = (item_local_index, global_molecule_id, step_index, modality)   modality: 0=a 1=c 2=e
    key     = (seed & 0xffffffff, seed >> 32)
    r       = philox4x32_10(counter, key)
    u_cat, u_unmask, u_remask = (r[0] >> 8, r[1] >> 8, r[2] >> 8) * 2**-24      (fp32-exact, in [0,1))

item_local_index is the atom index inside its molecule (a, c) or the upper-triangle edge index inside its
molecule in the reference's edge order (flowmol/data_processing/utils.py:4-17).  The CUDA kernel
(flowmol_b200/csrc/ctmc.cuh) evaluates the same function, so results do not depend on batch composition
or on how molecules are sharded across GPUs.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32 arrays / ints. Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def uniforms(item, mol, step, modality, seed):
    """Three fp32 uniforms in [0,1) per item: (categorical, unmask, re-mask)."""
    seed = int(seed)
    r = philox4x32_10(item, mol, step, modality, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    scale = np.float32(2.0 ** -24)
    return tuple((x >> np.uint32(8)).astype(np.float32) * scale for x in r[:3])
