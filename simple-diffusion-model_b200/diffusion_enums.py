"""Enumerations of the reference's public surface (diffusion_enums.py:5-14)."""
from enum import Enum


class DiffusionAlg(Enum):
    DDPM = 0
    DDIM = 1


class NoiseScheduler(Enum):
    LINEAR = 0
    COSINE = 1
