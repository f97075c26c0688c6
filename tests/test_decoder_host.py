"""CPU: host-side logic of the pixel-decoder mirror (reference: msdeformattn.py:165-312) and the oracle's sine
position embedding against the reference's own output (tests/golden/decoder/pixel_decoder_tiny.npz, made by
oracle/gen_golden_decoder.py from the unmodified reference classes)."""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden", "decoder", "pixel_decoder_tiny.npz")


def test_oracle_sine_embedding_matches_reference_golden():
    import msda_oracle
    z = np.load(GOLD)
    for key, (h, w) in (("pos_2x3", (2, 3)), ("pos_7x5", (7, 5))):
        got = msda_oracle.sine_position_embedding(h, w)
        assert got.shape == z[key].shape
        assert np.abs(got - z[key]).max() <= 2e-6


def test_mirror_state_dict_keys_equal_reference(built):
    import gen_golden_decoder as G
    from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
    shapes = {k: ShapeSpec(channels=c, stride=s) for k, (c, s) in G.CASE["input_shape"].items()}
    dec = MSDeformAttnPixelDecoder(shapes, **G.CASE["kwargs"])
    assert sorted(dec.state_dict().keys()) == list(np.load(GOLD)["state_dict_keys"])
    # msdeformattn.py:197-209, 260-261: features sorted by stride; encoder levels lowest resolution first; one FPN level
    assert dec.in_features == ["res2", "res3", "res4", "res5"]
    assert dec.transformer_in_features == ["res3", "res4", "res5"]
    assert [p[0].in_channels for p in dec.input_proj] == [320, 512, 256]
    assert dec.num_fpn_levels == 1 and dec.maskformer_num_feature_levels == 3
    assert dec.adapter_1.bias is None and dec.layer_1.bias is None          # norm == "GN" -> no conv bias


def test_from_config_uses_reference_keys(built):
    from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
    head = SimpleNamespace(IN_FEATURES=["res2", "res3", "res4", "res5"], CONVS_DIM=256, MASK_DIM=256, NORM="GN",
                           TRANSFORMER_ENC_LAYERS=6, DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES=["res3", "res4", "res5"],
                           COMMON_STRIDE=4)
    cfg = SimpleNamespace(MODEL=SimpleNamespace(SEM_SEG_HEAD=head, MASK_FORMER=SimpleNamespace(DROPOUT=0.0, NHEADS=8,
                                                                                            DIM_FEEDFORWARD=2048)))
    shapes = {f"res{i + 2}": ShapeSpec(channels=256 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    shapes["stem"] = ShapeSpec(channels=64, stride=2)
    kw = MSDeformAttnPixelDecoder.from_config(cfg, shapes)
    assert "stem" not in kw["input_shape"]
    assert kw["transformer_dim_feedforward"] == 1024          # hard-coded in the reference (msdeformattn.py:305-306)
    assert kw["transformer_enc_layers"] == 6 and kw["transformer_nheads"] == 8 and kw["common_stride"] == 4
