"""Host-side stand-in for the Fortran host (mesh, model, pre-computed terms, source)."""
from .spectral import SpectralBasis
from .mesh import MeshSpec, prem_mesh_spec, build_rank, surface_receivers
from .model import prem_layers, homogeneous_layers
from .precomp import AttenuationModel
from .source import SourceParams
from .problem import Problem, build_problem, stable_timestep
