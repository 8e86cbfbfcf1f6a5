"""jues.jl_b200 -- placeholder, filled in below."""
from . import synth  # noqa: F401
