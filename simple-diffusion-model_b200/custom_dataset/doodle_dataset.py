"""Image + conditioning-image pairs with the reference's interface (custom_dataset/doodle_dataset.py:14-65)."""
import random

from torch.utils.data import Dataset

from ._tables import load_tables
from .img_dataset import read_image


class DoodleImgDataset(Dataset):
    def __init__(self, dataset_path=None, shuffle_seed=None, raw_uint8=False):
        self.raw_uint8 = raw_uint8
        rows, self.all_labels = load_tables(dataset_path)
        (random.Random(shuffle_seed) if shuffle_seed is not None else random).shuffle(rows)
        key = self.all_labels[0]
        self.dataset = [(r["filename"], r[key]) for r in rows]

    def get_labels(self):
        return self.all_labels

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, index):
        img_path, label_path = self.dataset[index]
        return read_image(img_path, self.raw_uint8), read_image(label_path, self.raw_uint8)
