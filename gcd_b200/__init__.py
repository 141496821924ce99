"""gcd_b200 — B200-native (sm_100a) denoising hot path for Generative Camera Dolly (basilevh/gcd)."""
__version__ = "0.1.0"
