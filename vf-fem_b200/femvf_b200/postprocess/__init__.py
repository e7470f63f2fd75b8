"""Post-processing of simulation results (mirror of ``femvf.postprocess``)."""
from . import base, solid  # noqa: F401
