"""Synthetic stand-in for the reference's missing `datasets` package (test_diml_cvt.py:44,78):
select(name, opt, path)['testing'] is a map-style dataset with `.avail_classes` whose items are
(label, image[3, 224, 224], index).  Images of one class share a prototype, so retrieval metrics
are non-trivial.  Size: $VITRERANK_SHIM_N images (default 256), $VITRERANK_SHIM_CLASSES classes."""
import os

import torch


class SyntheticImages(torch.utils.data.Dataset):
    def __init__(self, n, classes, seed=0, size=224):
        g = torch.Generator().manual_seed(seed)
        self.labels = (torch.arange(n) * classes // n).tolist()
        self.avail_classes = sorted(set(self.labels))
        self.protos = torch.randn(classes, 3, 14, 14, generator=g)
        self.seed, self.size, self.n = seed, size, n
        self.image_list = [f"synthetic_{i:06d}.png" for i in range(n)]

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        idx = int(idx)
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx)
        low = self.protos[self.labels[idx]] + 0.7 * torch.randn(3, 14, 14, generator=g)
        img = torch.nn.functional.interpolate(low[None], size=(self.size, self.size), mode='nearest')[0]
        return self.labels[idx], img, idx


def select(dataset, opt, data_path):
    n = int(os.environ.get("VITRERANK_SHIM_N", "256"))
    classes = int(os.environ.get("VITRERANK_SHIM_CLASSES", "8"))
    return {'testing': SyntheticImages(n, classes, seed=getattr(opt, 'seed', 0))}
