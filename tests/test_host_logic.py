"""CPU: host-side mirror of the reference interface — module surface, init, workloads."""
import math

import pytest
import torch

from bm2f_b200 import workloads as W


def test_module_surface_and_state_dict_keys(built):
    # reference: ops/modules/ms_deform_attn.py:35-62
    from bm2f_b200.ops.modules import MSDeformAttn
    m = MSDeformAttn()
    assert (m.d_model, m.n_levels, m.n_heads, m.n_points, m.im2col_step) == (256, 4, 8, 4, 128)
    keys = sorted(m.state_dict().keys())
    assert keys == sorted(f"{n}.{p}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
                          for p in ("weight", "bias"))
    m3 = MSDeformAttn(256, 3, 8, 4)
    assert m3.sampling_offsets.out_features == 8 * 3 * 4 * 2
    assert m3.attention_weights.out_features == 8 * 3 * 4
    assert sum(p.numel() for p in m3.parameters()) == 205600
    with pytest.raises(ValueError):
        MSDeformAttn(d_model=250, n_heads=8)


def test_module_init_matches_reference_formula(built):
    # reference: ms_deform_attn.py:64-80
    from bm2f_b200.ops.modules import MSDeformAttn
    m = MSDeformAttn(256, 3, 8, 4)
    assert not m.sampling_offsets.weight.any() and not m.attention_weights.weight.any()
    assert not m.attention_weights.bias.any() and not m.value_proj.bias.any()
    thetas = torch.arange(8, dtype=torch.float32) * (2.0 * math.pi / 8)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(8, 1, 1, 2).repeat(1, 3, 4, 1)
    for i in range(4):
        grid[:, :, i, :] *= i + 1
    assert torch.equal(m.sampling_offsets.bias.detach(), grid.view(-1))
    assert torch.equal(W.compass_offset_bias(8, 3, 4), grid)


def test_module_refuses_cpu(built):
    from bm2f_b200.ops.modules import MSDeformAttn
    m = MSDeformAttn(256, 3, 8, 4)
    shapes, start = W.level_tensors(((2, 2), (4, 4), (8, 8)))
    x = torch.zeros(1, 84, 256)
    with pytest.raises(RuntimeError, match="CPU"):
        m(x, W.reference_points(((2, 2), (4, 4), (8, 8)), 1), x, shapes, start)


def test_workload_shapes_match_baseline_table():
    # BASELINE.md §4
    assert W.WORKLOADS[1].S == 5376 and W.WORKLOADS[2].S == 21504 and W.WORKLOADS[3].S == 8400
    assert W.WORKLOADS[4].S == 43008 and W.WORKLOADS[5].S == 5040
    for cfg, starts in {1: [0, 256, 1280], 2: [0, 1024, 5120], 3: [0, 400, 2000], 4: [0, 2048, 10240],
                        5: [0, 240, 1200]}.items():
        _, start = W.level_tensors(W.WORKLOADS[cfg].levels)
        assert start.tolist() == starts


def test_workload_inputs_are_seeded_and_model_like():
    a = W.make_inputs(((4, 4), (8, 8)), 2, seed=5)
    b = W.make_inputs(((4, 4), (8, 8)), 2, seed=5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["loc"].shape == (2, 80, 8, 2, 4, 2) and a["attn"].shape == (2, 80, 8, 2, 4)
    assert torch.allclose(a["attn"].sum((-1, -2)), torch.ones(2, 80, 8), atol=1e-5)
    inp = W.workload_inputs(1)
    frac_out = ((inp["loc"] < 0) | (inp["loc"] > 1)).any(-1).float().mean().item()
    assert 0.05 < frac_out < 0.3          # 7.5-15 % of points have an outside coordinate


def test_reference_points_are_pixel_centres():
    ref = W.reference_points(((2, 3),), 1)
    assert ref.shape == (1, 6, 1, 2)
    assert torch.allclose(ref[0, :, 0, 0], torch.tensor([1, 3, 5, 1, 3, 5]) / 6.0)
    assert torch.allclose(ref[0, :, 0, 1], torch.tensor([1, 1, 1, 3, 3, 3]) / 4.0)


def test_byte_model_matches_survey():
    S = 21504
    assert W.hbm_bytes(S, "fwd") == 3200 * S
    assert W.hbm_bytes(S, "bwd") == 6400 * S
    assert W.hbm_bytes(S, "fwd", "bf16") == 2176 * S
    assert W.gather_bytes(S, "fwd") == 49152 * S
    assert W.gather_bytes(S, "bwd") == 98304 * S


def test_encoder_mirror_surface(built):
    # reference: msdeformattn.py:22-60 (EncoderOnly), 92-113 (layer): names -> state-dict keys
    from bm2f_b200.encoder import (MSDeformAttnTransformerEncoder, MSDeformAttnTransformerEncoderLayer,
                                   MSDeformAttnTransformerEncoderOnly)
    enc = MSDeformAttnTransformerEncoderOnly(d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024,
                                             dropout=0.0, activation="relu", num_feature_levels=3, enc_n_points=4)
    keys = set(enc.state_dict().keys())
    assert "level_embed" in keys and enc.level_embed.shape == (3, 256)
    for i in range(6):
        for name in ("self_attn.sampling_offsets", "self_attn.attention_weights", "self_attn.value_proj",
                     "self_attn.output_proj", "norm1", "linear1", "linear2", "norm2"):
            assert f"encoder.layers.{i}.{name}.weight" in keys and f"encoder.layers.{i}.{name}.bias" in keys
    assert len(keys) == 1 + 6 * 16
    layer = MSDeformAttnTransformerEncoderLayer()
    assert layer.linear1.out_features == 1024 and layer.self_attn.n_levels == 4 and layer.dropout1.p == 0.1
    ref = MSDeformAttnTransformerEncoder.get_reference_points([(2, 3)], torch.ones(1, 1, 2), "cpu")
    assert torch.allclose(ref, W.reference_points(((2, 3),), 1))


def test_bench_host_logic_config_shard_and_graph_rule():
    """bench.py host logic shared by both arms: the config builder, the strong-scaling shard size and the rule that turns
    CUDA-graph replay on for launch-bound per-GPU steps."""
    import importlib.util
    import os
    import types

    from bm2f_b200 import dist as D
    from bm2f_b200 import workloads as W
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod_host", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = W.WORKLOADS[2]
    for world in (1, 2, 4, 8, 3, 16):
        assert bench.b200_shard(wl.batch, world) == D.shard_batch(wl.batch, world, 0)[1]
    cfg1 = bench.workload_config(wl, 1, 16, "weak")
    assert cfg1["global_batch"] == 16 and cfg1["batch_per_gpu"] == 16 and cfg1["collective"] == "none"
    assert cfg1["l2_policy"].startswith("inputs larger than L2: 1101 MB")
    cfg8 = bench.workload_config(wl, 8, 2, "strong", "default", True)
    assert cfg8["global_batch"] == 16 and cfg8["parallelism"] == "dp8" and cfg8["collective"] != "none" and cfg8["cuda_graph"]
    assert bench.workload_config(wl, 8, 16, "weak")["global_batch"] == 128
    auto = types.SimpleNamespace(graph=None)
    assert not bench.graph_enabled(auto, wl, 16) and not bench.graph_enabled(auto, wl, 8)
    assert bench.graph_enabled(auto, wl, 4) and bench.graph_enabled(auto, wl, 2)          # 86 016 / 43 008 rows per GPU
    assert bench.graph_enabled(auto, W.WORKLOADS[1], 1)                                    # one 512^2 image: launch-bound
    assert not bench.graph_enabled(types.SimpleNamespace(graph=False), wl, 2)
    assert bench.graph_enabled(types.SimpleNamespace(graph=True), wl, 16)
    small = bench.workload_config(W.WORKLOADS[1], 1, 1, "weak")
    assert "smaller than L2" in small["l2_policy"]
