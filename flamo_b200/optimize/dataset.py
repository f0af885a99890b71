"""Datasets with the interface of flamo.optimize.dataset (reference dataset.py:9-174): one
(input, target) pair expanded to `expand` identical items, plus the colourless-reverb variant."""
from typing import Optional

import torch
import torch.utils.data as data


class Dataset(torch.utils.data.Dataset):
    def __init__(self, input: torch.Tensor = torch.randn(1, 1), target: torch.Tensor = torch.randn(1, 1),
                 expand: int = 1, device=torch.get_default_device(), dtype: Optional[torch.dtype] = None):
        dtype = dtype if dtype is not None else input.dtype
        self.expand, self.device = expand, device
        self.input = input.to(device).to(dtype).expand(expand, *input.shape[1:])
        self.target = target.to(device).to(dtype).expand(expand, *target.shape[1:])

    def __len__(self):
        return len(self.target)

    def __getitem__(self, index):
        return self.input[index], self.target[index]


class DatasetColorless(Dataset):
    """Input: unit impulse at index 0 of dim 1; target: flat magnitude (reference dataset.py:54-85)."""

    def __init__(self, input_shape: tuple, target_shape: tuple, expand: int = 1000,
                 device=torch.get_default_device(), dtype: torch.dtype = torch.float32):
        x = torch.zeros(input_shape, device=device, dtype=dtype)
        x[:, 0, :] = 1
        super().__init__(input=x, target=torch.ones(target_shape, device=device, dtype=dtype), expand=expand,
                         device=device, dtype=dtype)


def get_dataloader(dataset, batch_size: int = 2000, shuffle: bool = True):
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, drop_last=True)


def split_dataset(dataset, split: float, device=torch.get_default_device()):
    n_train = int(len(dataset) * split)
    return data.random_split(dataset, [n_train, len(dataset) - n_train], generator=torch.Generator(device=device))


def load_dataset(dataset, batch_size: int = 2000, split: float = 0.8, shuffle: bool = True,
                 device=torch.get_default_device()):
    train_set, valid_set = split_dataset(dataset, split, device)
    return get_dataloader(train_set, batch_size, shuffle), get_dataloader(valid_set, batch_size, shuffle)
