"""zkfl_b200 -- B200-native Groth16/BN254 proving backend for the ZK-FL reference circuits.

The package directory name contains hyphens (it mirrors the reference repository name), so it
is imported through the `zkfl_b200` shim at the repository root.
"""
__version__ = "0.1.0"
