"""Seeded weights/inputs for parity tests — shared with the product's synthetic-data module (TEST INFRASTRUCTURE)."""
from gcd_b200.synthetic import seeded_inputs, seeded_state, seeded_tensor  # noqa: F401
