"""Latent / feature data path of the train step.

Restates the pieces of the reference that sit between the files on disk and ``SILoss``:

  * ``CustomDataset``      /root/reference/image/dataset.py:18-85   the on-disk layout produced by the reference's
                           preprocessing: ``<data_dir>/images/**`` (PNG / .npy pixels for the frozen target encoders),
                           ``<data_dir>/vae-sd/**.npy`` (SD-VAE moments, [1, 8, S, S] or [8, S, S] fp32) and
                           ``<data_dir>/vae-sd/dataset.json`` ({"labels": [[fname, label], ...]}), optionally
                           ``<data_dir>/<text_embeds_dir>/**.npy`` (caption embeddings); same sorting, same label lookup,
                           same 4-tuple per item.
  * ``sample_posterior``   /root/reference/image/train.py:84-91     the posterior draw, here ONE kernel
                           (``reed_sample_posterior``) instead of chunk views + four elementwise kernels; bit-identical
                           for the same device RNG state (the noise is drawn with the same ``torch.randn_like`` call).
  * ``LatentBatchLoader``  /root/reference/image/train.py:263-271,333-336  what ``DataLoader(pin_memory=True)`` +
                           ``.to(device)`` + ``x.squeeze(dim=1)`` do, as a double-buffered pipeline: batches are
                           collated straight into pinned staging buffers and copied host->device on a side stream
                           while the previous step computes.

Decoding the raw pixels is only needed by the frozen target encoders, which are outside this path (the bench and
the tests feed target features directly): ``CustomDataset(..., load_images=False)`` skips it and returns an empty
uint8 tensor in the image slot.
"""
from __future__ import annotations

import json
import os
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import ops

_IMAGE_EXT = {".png", ".jpg", ".jpeg", ".bmp", ".webp", ".tif", ".tiff", ".gif", ".ppm"}


def _walk(root: str) -> List[str]:
    return sorted(os.path.relpath(os.path.join(base, f), start=root) for base, _dirs, files in os.walk(root) for f in files)


class CustomDataset(Dataset):
    """Drop-in for dataset.py:18-85.  Items: (image, moments, label, text_embeds or zeros_like(moments))."""

    def __init__(self, data_dir, text_embeds_dir=None, load_images=True):
        self.images_dir = os.path.join(data_dir, "images")
        self.features_dir = os.path.join(data_dir, "vae-sd")
        self.load_images = load_images
        supported = _IMAGE_EXT | {".npy"}
        self.image_fnames = [f for f in _walk(self.images_dir) if self._file_ext(f) in supported]
        self.feature_fnames = [f for f in _walk(self.features_dir) if self._file_ext(f) in supported]
        with open(os.path.join(self.features_dir, "dataset.json"), "rb") as f:
            labels = dict(json.load(f)["labels"])
        labels = np.array([labels[fname.replace("\\", "/")] for fname in self.feature_fnames])
        self.labels = labels.astype({1: np.int64, 2: np.float32}[labels.ndim])
        self.text_embeds_dir = text_embeds_dir
        if text_embeds_dir is not None:
            self.full_text_embeds_dir = os.path.join(data_dir, text_embeds_dir)
            assert os.path.exists(self.full_text_embeds_dir), f"Text embeds dir {self.full_text_embeds_dir} does not exist"

    @staticmethod
    def _file_ext(fname):
        return os.path.splitext(fname)[1].lower()

    def __len__(self):
        assert len(self.image_fnames) == len(self.feature_fnames), \
            "Number of feature files and label files should be same"
        return len(self.feature_fnames)

    def _image(self, fname):
        if not self.load_images:
            return torch.empty(0, dtype=torch.uint8)
        path = os.path.join(self.images_dir, fname)
        if self._file_ext(fname) == ".npy":
            image = np.load(path)
            return torch.from_numpy(image.reshape(-1, *image.shape[-2:]))
        import PIL.Image
        with PIL.Image.open(path) as im:
            image = np.array(im)
        return torch.from_numpy(image.reshape(*image.shape[:2], -1).transpose(2, 0, 1).copy())

    def __getitem__(self, idx):
        image_fname = self.image_fnames[idx]
        image = self._image(image_fname)
        features = torch.from_numpy(np.load(os.path.join(self.features_dir, self.feature_fnames[idx])))
        label = torch.tensor(self.labels[idx])
        if self.text_embeds_dir is not None:
            text_fname = image_fname.replace(self._file_ext(image_fname), ".npy")          # dataset.py:82
            return image, features, label, torch.from_numpy(np.load(os.path.join(self.full_text_embeds_dir, text_fname)))
        return image, features, label, torch.zeros_like(features)


# --------------------------------------------------------------------------------------------------------------------
# posterior draw
# --------------------------------------------------------------------------------------------------------------------

def _per_channel(value, channels, device):
    """latents_scale / latents_bias as (device fp32 [C] or None, python float)."""
    if isinstance(value, torch.Tensor):
        if value.numel() == 1:
            return None, float(value)
        if value.numel() != channels:
            raise ValueError(f"latents scale/bias must be a scalar or hold {channels} per-channel values")
        return value.detach().to(device=device, dtype=torch.float32).reshape(channels).contiguous(), 0.0
    return None, float(value)


@torch.no_grad()
def sample_posterior(moments, latents_scale=1., latents_bias=0., noise=None):
    """train.py:84-91.  ``moments`` [B, 2C, H, W] fp32 on the GPU; returns ``(mean + std * randn) * scale + bias``.

    ``noise`` (optional, [B, C, H, W]) replaces the draw - used by the parity tests; by default the normals come from
    the same ``torch.randn_like(mean)`` call as the reference, so the device generator advances identically.
    """
    if not moments.is_cuda:
        raise RuntimeError("reed_b200 sample_posterior runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if moments.dim() < 2 or moments.shape[1] % 2:
        raise ValueError("moments must be [B, 2C, ...] (mean and std stacked on dim 1)")
    moments = moments.float().contiguous()
    B, C = moments.shape[0], moments.shape[1] // 2
    shape = (B, C) + tuple(moments.shape[2:])
    hw = 1
    for s in moments.shape[2:]:
        hw *= s
    if noise is None:
        noise = torch.randn_like(moments[:, :C])
    noise = noise.to(device=moments.device, dtype=torch.float32).contiguous()
    if tuple(noise.shape) != shape:
        raise ValueError(f"noise must have shape {shape}")
    scale_t, scale_s = _per_channel(latents_scale, C, moments.device)
    bias_t, bias_s = _per_channel(latents_bias, C, moments.device)
    out = torch.empty(shape, device=moments.device, dtype=torch.float32)
    ops._launch("reed_sample_posterior", ops._p(moments), ops._p(noise), ops._p(scale_t), ops._p(bias_t), scale_s, bias_s,
                ops._p(out), B, C, hw, ops._stream())
    return out


# --------------------------------------------------------------------------------------------------------------------
# pinned, double-buffered host -> device batches
# --------------------------------------------------------------------------------------------------------------------

class _Staging:
    """One set of pinned host buffers + their device twins for a batch of fixed shape."""

    def __init__(self, shapes, dtypes, device, pin):
        self.host = [torch.empty(s, dtype=d, pin_memory=pin) for s, d in zip(shapes, dtypes)]
        self.dev = [torch.empty(s, dtype=d, device=device) for s, d in zip(shapes, dtypes)]
        self.ready = None          # CUDA event: H2D copies of this slot have completed
        self.consumed = None       # CUDA event: the step that read this slot's device buffers has been enqueued and passed


class LatentBatchLoader:
    """Iterates ``(moments [B,2C,S,S], labels [B], text_embeds or None)`` device batches from a ``CustomDataset``.

    Sampling order follows ``DataLoader(shuffle=True, drop_last=True)`` semantics: a fresh ``torch.randperm`` of the
    whole dataset per epoch, of which the ranks take disjoint strided shards like accelerate's sharded sampler.  The
    permutation must be THE SAME on every rank for the shards to be disjoint, so it is drawn from a private generator
    seeded ``seed + epoch`` (accelerate synchronises the sampler's RNG across ranks to the same end) - never from the
    global CPU generator, which ``train.py:176`` seeds differently per rank and which SILoss draws its times from.  A
    caller-supplied ``generator`` is used as is: with ``world > 1`` it has to be seeded identically on all ranks.
    Each batch is collated into one of ``depth`` pinned staging slots and copied on ``copy_stream``; the consumer's
    stream waits on the slot's event, so the copy of batch i+1 overlaps the compute of batch i.  The device tensors of a
    slot are reused ``depth`` batches later - consume (or clone) a batch before asking for ``depth`` more.
    """

    def __init__(self, dataset, batch_size: int, device, *, rank: int = 0, world: int = 1, shuffle: bool = True,
                 generator: Optional[torch.Generator] = None, depth: int = 2, with_text: Optional[bool] = None,
                 seed: int = 0):
        self.dataset, self.batch_size, self.device = dataset, batch_size, torch.device(device)
        self.rank, self.world, self.shuffle, self.depth = rank, world, shuffle, max(2, depth)
        self.generator = generator
        self.seed, self.epoch = int(seed), 0
        self.with_text = (getattr(dataset, "text_embeds_dir", None) is not None) if with_text is None else with_text
        self.cuda = self.device.type == "cuda"
        self._slots: List[_Staging] = []
        self._copy_stream = torch.cuda.Stream(device=self.device) if self.cuda else None

    def __len__(self):
        return (len(self.dataset) // self.world) // self.batch_size

    def epoch_indices(self) -> List[int]:
        n = len(self.dataset)
        gen = self.generator
        if gen is None:
            gen = torch.Generator().manual_seed(self.seed + self.epoch)
        self.epoch += 1
        order = torch.randperm(n, generator=gen).tolist() if self.shuffle else list(range(n))
        shard = order[self.rank::self.world][: (n // self.world)]
        usable = len(shard) // self.batch_size * self.batch_size
        return shard[:usable]

    def _collate(self, indices: Sequence[int], slot_id: int):
        items = [self.dataset[i] for i in indices]
        moments = [it[1].squeeze(0) if it[1].dim() == 4 else it[1] for it in items]      # x.squeeze(dim=1) of train.py:334
        fields = [moments, [it[2] for it in items]]
        if self.with_text:
            fields.append([it[3] for it in items])
        if slot_id >= len(self._slots):
            shapes = [(len(items),) + tuple(f[0].shape) for f in fields]
            dtypes = [f[0].dtype for f in fields]
            self._slots.append(_Staging(shapes, dtypes, self.device, self.cuda))
        slot = self._slots[slot_id]
        if slot.consumed is not None:
            slot.consumed.synchronize()       # the step that read this slot's device tensors has finished with them
        for buf, vals in zip(slot.host, fields):
            torch.stack(list(vals), out=buf)
        if self.cuda:
            with torch.cuda.stream(self._copy_stream):
                for h, d in zip(slot.host, slot.dev):
                    d.copy_(h, non_blocking=True)
                slot.ready = torch.cuda.Event()
                slot.ready.record(self._copy_stream)
        else:
            for h, d in zip(slot.host, slot.dev):
                d.copy_(h)
        return slot

    def __iter__(self) -> Iterator:
        idx = self.epoch_indices()
        batches = [idx[i:i + self.batch_size] for i in range(0, len(idx), self.batch_size)]
        pending: List[_Staging] = []
        nxt = 0

        def refill():
            nonlocal nxt
            if nxt < len(batches):
                pending.append(self._collate(batches[nxt], nxt % self.depth))
                nxt += 1

        for _ in range(self.depth - 1):
            refill()
        while pending:
            slot = pending.pop(0)
            if self.cuda:
                torch.cuda.current_stream(self.device).wait_event(slot.ready)
            yield tuple(slot.dev) if self.with_text else (slot.dev[0], slot.dev[1], None)
            if self.cuda:                      # whatever the consumer enqueued on its stream for this batch
                slot.consumed = torch.cuda.Event()
                slot.consumed.record(torch.cuda.current_stream(self.device))
            # stage the next batch now that this step is enqueued: it lands in the slot read one step earlier, so the
            # host waits (if at all) on a step that has a successor queued behind it, and the copy overlaps compute
            refill()
