"""Entry point with the reference's CLI (`-c/--config-path`, `--device`) and JSON config schema (train_noise_cold_diffusion.py:24-469).
The step loop runs on the sm_100a kernels through b200/trainer.py; launch under torchrun for data parallelism."""
from b200.trainer import run_training


def main(raw_args=None):
    return run_training("cold", raw_args)


if __name__ == "__main__":
    main()
