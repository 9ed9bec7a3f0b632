"""CPU oracle: test infrastructure only (tests/, smoke(), bench.py cpu_baseline / --impl reference)."""
