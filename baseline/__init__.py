"""GPU-side baseline of the reference's own PyTorch path (bench.py only; not product, not oracle)."""
