"""Drop-in for the functions of the reference's ``image/utils.py`` that lie on the hot path's callers:
``load_legacy_checkpoints`` (generate.py:18,80-83) and ``load_encoders`` (train.py:23,225 - the frozen target encoders).
``download_model`` fetches published checkpoints over the network and is outside this package."""
import _reed_path  # noqa: F401

from reed_b200.image.generate import load_legacy_checkpoints  # noqa: F401


def load_encoders(enc_type, device, resolution=256):
    """train.py:225.  Returns (encoders, encoder_types, architectures) like utils.py:55-164 for the DINOv2 family, built on
    reed_b200.image.encoders (weights must be supplied locally: torch.hub is not reachable from an air-gapped box)."""
    from reed_b200.image.encoders import load_encoders as _load
    return _load(enc_type, device, resolution)


def download_model(model_name):
    raise NotImplementedError("download_model needs network access to the published checkpoints; load a local state_dict "
                              "with reed_b200.image.generate.load_sampling_weights instead")
