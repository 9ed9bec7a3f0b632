"""Dataset readers with the reference class names, plus the synthetic dataset used by tests and benches."""
