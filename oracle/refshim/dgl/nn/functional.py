"""Import-only stand-in: both shipped configs have `attention: False` (configs/flowmol3.yml, configs/dev.yml)."""


def edge_softmax(*a, **k):
    raise NotImplementedError("edge_softmax is not on the sampling path (attention: False)")
