"""Raw-image preprocessing in front of the target encoders (train.py:53-74): the tap-level oracle against outputs of the
reference function (tests/golden/preprocess.pt, made by oracle/make_golden.py --only-preprocess) and against torch's bicubic."""
import pytest
import torch
import torch.nn.functional as F

from oracle import preprocess_oracle


def _input(case):
    r = case["resolution"]
    return torch.randint(0, 256, (2, 3, r, r), generator=torch.Generator().manual_seed(case["seed"]), dtype=torch.uint8)


def test_oracle_matches_reference_function(golden):
    cases = golden("preprocess.pt")
    assert sorted(cases) == ["clip/256", "dinov1/256", "dinov2/256", "dinov2/512", "jepa/256", "mae/256", "mocov3/256",
                             "siglip/256"]
    for name, case in cases.items():
        x = _input(case)
        y = preprocess_oracle.preprocess_raw_image(x, case["enc_type"])
        assert tuple(y.shape) == case["shape"] and str(y.dtype) == case["dtype"], name
        # bar for this floating-point path: 1e-5 absolute on values of magnitude <= 2.7 (bicubic overshoot included)
        assert float((y[..., ::7, ::5].float() - case["sample"].float()).abs().max()) < 1e-5, name
        assert abs(float(y.double().sum()) - case["total"]) < 1e-6 * case["abs_total"] + 1e-3, name
    assert cases["dinov2/256"]["shape"] == (2, 3, 224, 224) and cases["dinov2/512"]["shape"] == (2, 3, 448, 448)
    assert cases["siglip/256"]["dtype"] == "torch.uint8"            # unknown encoder types pass through untouched


@pytest.mark.parametrize("size_in,size_out", [(256, 224), (512, 448), (37, 50), (64, 17)])
def test_tapwise_bicubic_is_torchs_bicubic(size_in, size_out):
    x = torch.rand(2, 3, size_in, size_in, generator=torch.Generator().manual_seed(size_in))
    want = F.interpolate(x, size_out, mode="bicubic")
    assert float((preprocess_oracle.bicubic_resize(x, size_out) - want).abs().max()) < 5e-6


def test_product_plan_and_cpu_refusal():
    from reed_b200.image import preprocess
    assert preprocess._plan("dinov2-vit-b", 256) == (preprocess.IMAGENET_DEFAULT_MEAN, preprocess.IMAGENET_DEFAULT_STD, 224, 0)
    assert preprocess._plan("dinov2-vit-l", 512)[2] == 448
    assert preprocess._plan("clip-vit-L", 256)[3] == 1 and preprocess._plan("clip-vit-L", 256)[0] == preprocess.CLIP_DEFAULT_MEAN
    assert preprocess._plan("mocov3-vit-b", 256)[2] == 256 and preprocess._plan("dinov1", 256)[2] == 256
    assert preprocess._plan("dinov1-vit-b", 256)[2] == 256          # the reference tests `'dinov1' in enc_type` (train.py:66)
    assert preprocess._plan("siglip", 256) is None
    x = torch.zeros(1, 3, 256, 256, dtype=torch.uint8)
    assert preprocess.preprocess_raw_image(x, "siglip") is x
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        preprocess.preprocess_raw_image(x, "dinov2")
    assert (preprocess.IMAGENET_DEFAULT_MEAN, preprocess.CLIP_DEFAULT_STD) == \
        (preprocess_oracle.IMAGENET_DEFAULT_MEAN, preprocess_oracle.CLIP_DEFAULT_STD)
