"""Labelled image dataset with the reference's interface (custom_dataset/conditional_img_dataset.py:14-64)."""
import random

import torch
from torch.utils.data import Dataset

from ._tables import load_tables
from .img_dataset import read_image


class ConditionalImgDataset(Dataset):
    def __init__(self, dataset_path=None, shuffle_seed=None, raw_uint8=False):
        self.raw_uint8 = raw_uint8
        rows, self.all_labels = load_tables(dataset_path)
        # the reference shuffles once in case the table is sorted; data-parallel ranks pass a common seed so that
        # they all index the same row order
        (random.Random(shuffle_seed) if shuffle_seed is not None else random).shuffle(rows)
        self.dataset = [(r["filename"], [float(r[name]) for name in self.all_labels]) for r in rows]

    def get_labels(self):
        return self.all_labels

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, index):
        path, labels = self.dataset[index]
        return read_image(path, self.raw_uint8), torch.Tensor(labels)
