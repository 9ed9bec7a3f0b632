"""Entry point with the reference's CLI (`-c/--config-path`, `--device`) and JSON config schema (train_doodle_diffusion.py:26-465).
The step loop runs on the sm_100a kernels through b200/trainer.py; launch under torchrun for data parallelism."""
from b200.trainer import run_training


def main(raw_args=None):
    return run_training("doodle", raw_args)


if __name__ == "__main__":
    main()
