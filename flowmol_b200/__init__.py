"""flowmol_b200 -- B200-native sampling hot path of FlowMol (CTMC flow-matching integrator + GVP vector field).

Public surface mirrors the reference's (flowmol/__init__.py:30, flowmol/models/flowmol.py:473-589):

    import flowmol_b200 as flowmol
    model = flowmol.load_pretrained('flowmol3').cuda().eval()
    mols  = model.sample_random_sizes(n_molecules=10, n_timesteps=250)
"""
from .config import ModelConfig  # noqa: F401


def __getattr__(name):
    # heavy members are imported lazily so that `import flowmol_b200` works on a CPU-only box
    if name in ("load_pretrained", "FlowMolB200", "pretrained_model_names"):
        from . import api
        return getattr(api, name)
    if name in ("CTMCVectorFieldB200",):
        from . import vector_field
        return getattr(vector_field, name)
    raise AttributeError(name)
