"""Circuit front-end (circom replacement): builder DSL, template library, main components."""
from .builder import CircuitBuilder, CompiledCircuit, LC, FR  # noqa: F401
from .library import CIRCUIT_NAMES, build_circuit  # noqa: F401
