"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by flowmol_b200/).

CPU restatement of the reference's sampling hot path
(FlowMol.sample -> CTMCVectorField.integrate -> step -> EndpointVectorField.forward)
plus the tooling that pins it to the reference's own code executed verbatim.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package, and only as the checker / CPU baseline.
"""
