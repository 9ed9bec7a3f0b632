"""Host-side helpers with the reference's names and semantics (utils/utils.py): console progress bar, sample-grid
plotting and checkpoint files ({"model", "optimizer"} dicts saved with torch.save, interchangeable with the reference)."""
import os

import torch


def printProgressBar(iteration, total, prefix='', suffix='', decimals=1, length=100, fill='█', printEnd="\r", log=print):
    """One-line console bar redrawn in place; `log` is the callable the trainers pass so that rank != 0 stays silent
    (reference utils/utils.py:8-36, same keyword names because the train scripts call it by keyword)."""
    frac = iteration / float(total)
    done = int(length * iteration // total)
    log(f"\r{prefix} |{fill * done}{'-' * (length - done)}| {100 * frac:.{decimals}f}% {suffix}", end=printEnd)
    if iteration == total:
        log()


def plot_sampled_images(sampled_imgs, file_name, dest_path=None, log=print):
    """BGR -> RGB, 5-per-row grid normalised from [-1, 1], written as <dest>/plots/<file_name>.jpg (reference
    utils/utils.py:39-65).  Samples that live on the GPU never travel as fp32: channel swap, make_grid and save_image's
    quantisation are one kernel (b200.image_io.image_grid_u8) and only the uint8 picture is copied back for JPEG encoding --
    the bytes handed to the encoder are identical to torchvision's."""
    base = os.path.dirname(os.path.abspath(__file__)) if dest_path is None else dest_path
    out_dir = os.path.join(base, "plots")
    os.makedirs(out_dir, exist_ok=True)
    try:
        path = os.path.join(out_dir, str(file_name) + ".jpg")
        if sampled_imgs.is_cuda and sampled_imgs.shape[1] in (1, 3):
            from PIL import Image
            from b200.image_io import image_grid_u8
            picture = image_grid_u8(sampled_imgs, nrow=5, padding=2, value_range=(-1, 1), swap_rb=True).cpu().numpy()
            Image.fromarray(picture[:, :, 0] if picture.shape[2] == 1 else picture).save(path)
        else:
            import torchvision
            grid = torchvision.utils.make_grid(sampled_imgs[:, [2, 1, 0]], nrow=5, normalize=True, value_range=(-1, 1))
            torchvision.utils.save_image(grid, path)
        log(f"Saving generated image: {path}")
    except Exception as e:  # same forgiving behaviour as the reference
        log(f"An error occured while plotting reconstructed image: {e}")


def save_model(model_net, file_name, dest_path, checkpoint=False, steps=0, log=print):
    """torch.save of `model_net` (a state dict or a {"model", "optimizer", ...} bundle) to
    <dest>/checkpoint/<name>_<steps>.pt or <dest>/models/<name>_<steps>.pt; True on success, never raises
    (reference utils/utils.py:67-82).  The trainers call it on rank 0 only."""
    try:
        folder = os.path.join(dest_path, "checkpoint" if checkpoint else "models")
        os.makedirs(folder, exist_ok=True)
        torch.save(model_net, os.path.join(folder, f"{file_name}_{str(steps)}.pt"))
        return True
    except Exception as e:
        log(f"Exception occured while saving model: {e}.")
        return False


def load_checkpoint(checkpoint_path, log=print):
    """Returns (ok, object); maps to CPU; never raises (reference utils/utils.py:85-95)."""
    if not os.path.exists(checkpoint_path):
        log("Checkpoint does not exist.")
        return False, None
    log(f"Loading checkpoint: {checkpoint_path}")
    try:
        return True, torch.load(checkpoint_path, map_location=torch.device("cpu"), weights_only=False)
    except Exception:
        return False, None
