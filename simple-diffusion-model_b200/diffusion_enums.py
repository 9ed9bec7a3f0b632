"""Enumerations of the reference's public surface (diffusion_enums.py:5-14): member names and integer values are part of the
drop-in contract (the CLIs compare `member.name.lower()` against `--diff_alg`, the export JSON stores `member.name`)."""
import enum

# Which reverse process a generator / trainer plot uses: ancestral DDPM or the deterministic DDIM skip schedule.
DiffusionAlg = enum.Enum("DiffusionAlg", [("DDPM", 0), ("DDIM", 1)], module=__name__)

# Forward-process schedule: linear beta table (NoiseDegradation) or closed-form cosine alpha-bar (CosineNoiseDegradation).
NoiseScheduler = enum.Enum("NoiseScheduler", [("LINEAR", 0), ("COSINE", 1)], module=__name__)
