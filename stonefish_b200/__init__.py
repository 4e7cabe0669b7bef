"""stonefish_b200 — B200 (sm_100a) implementation of Stonefish's per-triangle hydrodynamic force integration.

The product is the C-ABI shared library `libsfh.so` (include/sfh.h, sources in stonefish_b200/csrc). This
package binds it with ctypes and mirrors the reference's host interface for that path (hydro.py); scenes.py
and meshes.py hold the synthetic BASELINE inputs. Importing the package does not require a GPU; creating a
`HydroContext` does, and fails loudly without one — there is no CPU fallback.
"""
from . import _abi
from .hydro import (BODY_STATE_DTYPE, INFO_DTYPE, Atmosphere, BodyFluidPosition, GeometryApproxType, HydroContext,
                    HydrodynamicsSettings, Jet, Mesh, Ocean, Pipe, SfhError, Uniform, VelocityField, aero_proxy, compound_part_states,
                    make_transform, new_body_states, quat_to_basis, sum_compound, sum_compound_aero, transform_mul, wave_params)

__all__ = ["_abi", "BODY_STATE_DTYPE", "INFO_DTYPE", "Atmosphere", "BodyFluidPosition", "GeometryApproxType", "HydroContext",
           "HydrodynamicsSettings", "Jet", "Mesh", "Ocean", "Pipe", "SfhError", "Uniform", "VelocityField", "aero_proxy",
           "compound_part_states", "make_transform", "new_body_states", "quat_to_basis", "sum_compound", "sum_compound_aero",
           "transform_mul", "wave_params"]
