"""Entry point with the reference's function name, signature and CLI (reference generate_sr_images_diffusion.py); body in b200/generator.py."""
from b200.generator import generate_sr_images_diffusion  # noqa: F401

if __name__ == "__main__":
    generate_sr_images_diffusion()
