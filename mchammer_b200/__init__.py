"""mchammer_b200 -- MCHammer's Monte-Carlo step loop on NVIDIA B200 (sm_100a).

The public names are the reference's (src/mchammer/__init__.py:19-36); ``Optimizer`` and
``Collapser`` run their loops in hand-written CUDA kernels through libmchammer_b200.so and
gain batched entry points (``get_results_batch``).  Importing the package never needs a GPU;
calling a driver without CUDA or without the built library raises.
"""

from .collapser import CollapseBatchResult, Collapser
from .elements import get_radius
from .batch import BatchPlan, BatchResult, BatchTrajectory, shard_bounds
from .optimizer import Optimizer
from .primitives import (
    get_atom_distance,
    get_bond_vector,
    rotate_molecule_by_angle,
    rotation_matrix_arbitrary_axis,
    test_move,
    translate_atoms_along_vector,
    translate_molecule_along_vector,
)
from .records import MCStepResult, Result, StepResult
from .structure import Atom, Bond, Molecule

__all__ = [
    "Atom",
    "Bond",
    "Molecule",
    "get_radius",
    "StepResult",
    "MCStepResult",
    "Result",
    "get_atom_distance",
    "Optimizer",
    "Collapser",
    "get_bond_vector",
    "translate_atoms_along_vector",
    "translate_molecule_along_vector",
    "test_move",
    "rotate_molecule_by_angle",
    "rotation_matrix_arbitrary_axis",
    # additions of this framework
    "BatchPlan",
    "BatchResult",
    "BatchTrajectory",
    "CollapseBatchResult",
    "shard_bounds",
]
