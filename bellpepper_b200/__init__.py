"""bellpepper_b200 -- B200-native R1CS evaluation engine behind bellpepper's ConstraintSystem interface.

Only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of include/bp_r1cs.h), `ffi` (ctypes
binding), `cs` (host-side mirror of the reference interface).  Importing the package does not load the
native library; the first use does, and fails loudly if it is missing (there is no CPU path).
"""

from . import fields  # noqa: F401
from .cs import (  # noqa: F401
    AUX,
    INPUT,
    ONE,
    AssignmentMissing,
    DivisionByZero,
    Index,
    LinearCombination,
    Namespace,
    NativeError,
    SizedWitness,
    SynthesisError,
    TestConstraintSystem,
    Unsatisfiable,
    Variable,
    WitnessCS,
    compute_path,
)

__version__ = "0.1.0"
