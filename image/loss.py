"""Drop-in for the reference's ``image/loss.py``: ``from loss import SILoss`` (train.py:22) with cwd = image/."""
import _reed_path  # noqa: F401

from reed_b200.image.loss import IMAGE_ENCODERS, SILoss, mean_flat, sum_flat  # noqa: F401
