"""Drop-in for the part of the reference's ``image/dataset.py`` on the latent data path: ``from dataset import CustomDataset``
(train.py:25) with cwd = image/.  ``LatentBatchLoader`` / ``sample_posterior`` are the B200-side additions around it."""
import _reed_path  # noqa: F401

from reed_b200.image.dataset import CustomDataset, LatentBatchLoader, sample_posterior  # noqa: F401
