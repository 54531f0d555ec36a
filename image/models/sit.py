"""Drop-in for the reference's ``image/models/sit.py``: ``from models.sit import SiT_models`` (train.py:21, generate.py:9)
with cwd = image/ resolves here and gets the B200-native SiT (reed_b200.image.models.sit: same constructor kwargs, forward
signature, attributes and state_dict layout; every arithmetic step runs in libreed_sm100.so)."""
import _reed_path  # noqa: F401

from reed_b200.image.models.sit import (  # noqa: F401
    FinalLayer, LabelEmbedder, PatchEmbed, SiT, SiT_models, SiTBlock, TimestepEmbedder, build_mlp, get_1d_sincos_pos_embed_from_grid,
    get_2d_sincos_pos_embed, get_2d_sincos_pos_embed_from_grid, modulate)

# the reference also exposes one constructor per zoo entry (sit.py:373-407)
SiT_XL_2, SiT_XL_4, SiT_XL_8 = SiT_models["SiT-XL/2"], SiT_models["SiT-XL/4"], SiT_models["SiT-XL/8"]
SiT_L_2, SiT_L_4, SiT_L_8 = SiT_models["SiT-L/2"], SiT_models["SiT-L/4"], SiT_models["SiT-L/8"]
SiT_B_2, SiT_B_4, SiT_B_8 = SiT_models["SiT-B/2"], SiT_models["SiT-B/4"], SiT_models["SiT-B/8"]
SiT_S_2, SiT_S_4, SiT_S_8 = SiT_models["SiT-S/2"], SiT_models["SiT-S/4"], SiT_models["SiT-S/8"]
