"""Labelled image dataset with the reference's interface (custom_dataset/conditional_img_dataset.py:14-64)."""
import random

import torch
from torch.utils.data import Dataset

from ._tables import load_tables
from .img_dataset import read_image


class ConditionalImgDataset(Dataset):
    def __init__(self, dataset_path=None):
        rows, self.all_labels = load_tables(dataset_path)
        random.shuffle(rows)                      # the reference shuffles once in case the table is sorted
        self.dataset = [(r["filename"], [float(r[name]) for name in self.all_labels]) for r in rows]

    def get_labels(self):
        return self.all_labels

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, index):
        path, labels = self.dataset[index]
        return read_image(path), torch.Tensor(labels)
