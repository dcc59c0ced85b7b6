"""Seeded synthetic galleries with the shapes of the reference's test sets.

The reference embeds a real dataset (evaluation/eval_cvt_diml.py:247-305) into three
fp32 banks; datasets and checkpoints are not available, so the benchmark and the
parity tests build banks of the same shape and value range from a seed
(SURVEY.md section 8d):

  patches  [N, C, R]  per-patch L2-normalised over C   (eval_cvt_diml.py:304)
  centers  [N, C]     L2-normalised                    (eval_cvt_diml.py:305)
  rollout  [N, R]     strictly positive, mean 1/R      (eval_cvt_diml.py:256)
  labels   [N] int64  class-structured

Everything is generated on the CPU with a seeded torch.Generator so the oracle
and the CUDA path see bit-identical inputs on any machine.
"""
from __future__ import annotations

import dataclasses

import torch

# name -> (N, C, R, classes)
SHAPES = {
    "cub200": (5924, 128, 49, 100),
    "cars196": (8131, 128, 49, 98),
    "sop": (60502, 128, 49, 11316),
    "sop_vitb16": (60502, 768, 196, 11316),
}


@dataclasses.dataclass
class Gallery:
    patches: torch.Tensor   # [N, C, R] fp32
    centers: torch.Tensor   # [N, C] fp32
    rollout: torch.Tensor   # [N, R] fp32
    labels: torch.Tensor    # [N] int64

    @property
    def shape(self):
        n, c, r = self.patches.shape
        return n, c, r

    def to(self, device):
        return Gallery(self.patches.to(device), self.centers.to(device),
                       self.rollout.to(device), self.labels.to(device))

    def pin(self):
        return Gallery(self.patches.pin_memory(), self.centers.pin_memory(),
                       self.rollout.pin_memory(), self.labels.pin_memory())


def make_labels(n: int, classes: int, gen: torch.Generator) -> torch.Tensor:
    """Contiguous class blocks with uneven sizes (every class has >= 2 images)."""
    classes = max(1, min(classes, n // 2))
    w = torch.rand(classes, generator=gen) + 0.5
    sizes = torch.floor(w / w.sum() * (n - 2 * classes)).long() + 2
    rest = n - int(sizes.sum())
    sizes[:rest] += 1
    return torch.repeat_interleave(torch.arange(classes), sizes)[:n].contiguous()


def make_gallery(n: int, c: int = 128, r: int = 49, classes: int = 100, seed: int = 0,
                 sigma: float = 0.6, structured: bool = True, chunk: int = 4096) -> Gallery:
    """Class-structured banks: patch = normalize(prototype[class] + sigma * noise).

    sigma ~0.3-1.0 spreads positive-pair patch similarities over ~0.2-0.9, which gives
    data-dependent Sinkhorn iteration counts like real features; structured=False gives
    iid Gaussian patches (the low-iteration regime, n* ~ 13).
    """
    gen = torch.Generator().manual_seed(seed)
    labels = make_labels(n, classes, gen)
    ncls = int(labels.max()) + 1
    patches = torch.empty(n, c, r, dtype=torch.float32)
    centers = torch.empty(n, c, dtype=torch.float32)
    proto = torch.randn(ncls, c, r, generator=gen) if structured else None
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        x = torch.randn(hi - lo, c, r, generator=gen)
        if structured:
            x = proto[labels[lo:hi]] + sigma * x
        x = torch.nn.functional.normalize(x, p=2, dim=1)
        patches[lo:hi] = x
        g = x.mean(dim=2) + 0.3 / (c ** 0.5) * torch.randn(hi - lo, c, generator=gen)
        centers[lo:hi] = torch.nn.functional.normalize(g, p=2, dim=1)
    rollout = torch.softmax(torch.randn(n, r, generator=gen), dim=-1).contiguous()
    return Gallery(patches, centers, rollout, labels)


def make_named(name: str, seed: int = 0, n: int | None = None, **kw) -> Gallery:
    n0, c, r, classes = SHAPES[name]
    if n is not None and n != n0:
        classes = max(2, int(round(classes * n / n0)))
        n0 = n
    return make_gallery(n0, c, r, classes, seed=seed, **kw)
