"""axisem_b200 — B200-native AxiSEM SOLVER time loop."""
__version__ = "0.1.0"
