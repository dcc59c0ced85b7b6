"""On-disk gallery bank: the feature cache the reference has but keeps switched off
(evaluation/eval_diml.py:80-85 loads `feat.pt` behind `and False`, :151-153 would torch.save the
feature bank / centre bank / labels), as a versioned, mmap-able binary file instead of a pickle.

Layout (little endian):
    0    8 bytes   magic  b"VRBANK\\x00\\x01"            (last byte = format version 1)
    8    int64[6]  n, c, r, has_rollout, has_labels, reserved
    56   8 bytes   64-bit checksum of the payload (per array, in file order: an FNV-style fold over blocks of
                   2^20 little-endian u64 words, h = ((h ^ xor(block)) * prime) ^ (sum(block) * prime))
    64   payload, every array 64-byte aligned, in this order:
         labels  int64 [n]          (when has_labels)
         centers fp32  [n, c]       L2-normalised               (eval_cvt_diml.py:305)
         rollout fp32  [n, r]       (when has_rollout)          (eval_cvt_diml.py:256)
         patches fp32  [n, c, r]    per-patch L2-normalised     (eval_cvt_diml.py:304), r contiguous
`load(..., mmap=True)` returns CPU tensors that alias the file (no copy; pin or upload from them directly);
a wrong magic / version / checksum / size raises BankFileError.
"""
from __future__ import annotations

import os

import numpy as np
import torch

MAGIC = b"VRBANK\x00\x01"
HEADER = 64
ALIGN = 64


class BankFileError(RuntimeError):
    pass


def _digest(chunks) -> int:
    """64-bit checksum of the padded arrays, one after the other (see the module docstring)."""
    h = np.uint64(0xcbf29ce484222325)
    prime = np.uint64(0x100000001b3)
    with np.errstate(over="ignore"):
        for a in chunks:
            w = np.frombuffer(a, dtype=np.uint64)
            # vectorised, order-sensitive across blocks, and cheap enough for GB-sized banks (a byte-serial FNV would
            # take minutes in Python)
            for lo in range(0, w.size, 1 << 20):
                blk = w[lo:lo + (1 << 20)]
                h = (h ^ np.bitwise_xor.reduce(blk)) * prime
                h ^= (blk.sum(dtype=np.uint64) * prime)
    return int(h)


def _pad(nbytes: int) -> int:
    return (nbytes + ALIGN - 1) // ALIGN * ALIGN


def _layout(n, c, r, has_rollout, has_labels):
    off = HEADER
    out = {}
    for name, on, nbytes in (("labels", has_labels, n * 8), ("centers", True, n * c * 4),
                             ("rollout", has_rollout, n * r * 4), ("patches", True, n * c * r * 4)):
        if on:
            out[name] = (off, nbytes)
            off += _pad(nbytes)
    return out, off


def save(path, patches, centers, rollout=None, labels=None):
    """Write the banks (CPU or CUDA tensors) to `path` atomically (temp file + rename)."""
    patches = patches.detach().to("cpu", torch.float32).contiguous()
    centers = centers.detach().to("cpu", torch.float32).contiguous()
    n, c, r = patches.shape
    if centers.shape != (n, c):
        raise BankFileError(f"centers {tuple(centers.shape)} do not match patches {tuple(patches.shape)}")
    arrays = {"patches": patches.numpy(), "centers": centers.numpy()}
    if rollout is not None:
        arrays["rollout"] = rollout.detach().to("cpu", torch.float32).contiguous().numpy().reshape(n, r)
    if labels is not None:
        arrays["labels"] = labels.detach().to("cpu", torch.int64).contiguous().numpy().reshape(n)
    lay, total = _layout(n, c, r, rollout is not None, labels is not None)
    order = [k for k in ("labels", "centers", "rollout", "patches") if k in lay]
    padded = []
    for k in order:
        raw = arrays[k].tobytes()
        padded.append(raw + b"\0" * (_pad(len(raw)) - len(raw)))
    digest = _digest(padded)
    head = MAGIC + np.array([n, c, r, int(rollout is not None), int(labels is not None), 0], dtype="<i8").tobytes() + \
        np.array([digest], dtype="<u8").tobytes()
    assert len(head) == HEADER
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(head)
        for b in padded:
            f.write(b)
    os.replace(tmp, path)
    return total


def load(path, mmap=True, verify=True):
    """-> (patches [n, c, r], centers [n, c], rollout [n, r] or None, labels [n] or None) as CPU tensors."""
    size = os.path.getsize(path)
    if size < HEADER:
        raise BankFileError(f"{path}: too short for a bank file")
    with open(path, "rb") as f:
        head = f.read(HEADER)
    if head[:7] != MAGIC[:7]:
        raise BankFileError(f"{path}: not a vitrerank bank file")
    if head[7] != MAGIC[7]:
        raise BankFileError(f"{path}: bank file version {head[7]}, this build reads version {MAGIC[7]}")
    n, c, r, has_roll, has_lab, _ = np.frombuffer(head[8:56], dtype="<i8").tolist()
    digest = int(np.frombuffer(head[56:64], dtype="<u8")[0])
    if n <= 0 or c <= 0 or r <= 0:
        raise BankFileError(f"{path}: bad shape [{n}, {c}, {r}]")
    lay, total = _layout(n, c, r, bool(has_roll), bool(has_lab))
    if size != total:
        raise BankFileError(f"{path}: {size} bytes, header says {total}")
    buf = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    order = [k for k in ("labels", "centers", "rollout", "patches") if k in lay]
    if verify and _digest([buf[lay[k][0]:lay[k][0] + _pad(lay[k][1])] for k in order]) != digest:
        raise BankFileError(f"{path}: checksum mismatch (truncated or corrupted bank file)")

    def view(name, dtype, shape):
        if name not in lay:
            return None
        off, nbytes = lay[name]
        a = np.frombuffer(buf, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize, offset=off).reshape(shape)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")   # read-only memmap -> tensor: the banks are never written through it
            return torch.from_numpy(a)
    return (view("patches", np.float32, (n, c, r)), view("centers", np.float32, (n, c)),
            view("rollout", np.float32, (n, r)), view("labels", np.int64, (n,)))
