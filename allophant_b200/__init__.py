"""allophant_b200 — B200-native acoustic-model forward/loss path of kgnlp/allophant."""
__version__ = "0.1.0"
