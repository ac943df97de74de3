"""CPU: the SAM2 engine's HOST logic (window-major token orders, q-pooling tables, attention work lists, padded / folded /
composed weights, prompt handling) run end to end on a plain-torch ops backend against the HF port of SAM2
(transformers.models.sam2, random-init, reduced depth / resolution): features, low-res mask logits and IoU predictions."""
import numpy as np
import pytest
import torch

from sam2_torch_ops import TorchOps
from socioreasoner_b200.sam2 import layout
from socioreasoner_b200.sam2.config import Sam2Cfg
from socioreasoner_b200.sam2.engine import Sam2Engine
from socioreasoner_b200.sam2.weights import pack


def small_hf_model(seed=0, image=256):
    from transformers.models.sam2 import configuration_sam2 as c, modeling_sam2 as m
    bc = dict(hidden_size=72, embed_dim_per_stage=[72, 144, 288, 576], num_attention_heads_per_stage=[1, 2, 4, 8], blocks_per_stage=[1, 2, 3, 2],
              global_attention_blocks=[4], window_size_per_stage=[8, 4, 16, 8], image_size=[image, image], num_attention_heads=1)
    g = image // 4
    vc = dict(backbone_config=bc, backbone_channel_list=[576, 288, 144, 72], backbone_feature_sizes=[[g, g], [g // 2, g // 2], [g // 4, g // 4]])
    cfg = c.Sam2Config(vision_config=vc, prompt_encoder_config=dict(image_size=image), mask_decoder_config={})
    torch.manual_seed(seed)
    model = m.Sam2Model(cfg).eval()
    with torch.no_grad():   # HF init leaves embeddings / biases at zero or tiny values: randomise so every term is exercised
        gen = torch.Generator().manual_seed(seed + 1)
        for n, p in model.named_parameters():
            if p.dim() == 1 or "embed" in n or "token" in n:
                p.copy_(torch.randn(p.shape, generator=gen) * (0.5 if "embed" in n or "token" in n else 0.1) + (1.0 if "norm" in n and n.endswith("weight") else 0.0))
    return model


def test_layout_tables():
    perm = layout.wm_perm(16, 4)
    assert sorted(perm.tolist()) == list(range(256)) and perm[:5].tolist() == [0, 1, 2, 3, 16] and perm[16] == 4
    idx = layout.pool_sources(16, 4)               # 8x8 outputs in window-major(2) order
    x = torch.arange(256.0).view(16, 16)
    pooled = torch.nn.functional.max_pool2d(x[None, None], 2)[0, 0]
    xw = x.reshape(-1)[torch.from_numpy(perm).long()]
    got = xw[torch.from_numpy(idx).long()].amax(1)
    want = pooled.reshape(-1)[torch.from_numpy(layout.wm_perm(8, 2)).long()]
    assert torch.equal(got, want)
    p = layout.window_plan(6, 16, 64)               # pooled first block of a stage: 8 windows of 16 query rows per tile
    assert p.tiles.tolist() == [[0, 96, 0, 384]] and p.row_lo[17] == 64 and p.row_hi[17] == 128
    p = layout.window_plan(3, 256, 256)
    assert p.tiles[:2].tolist() == [[0, 128, 0, 256], [128, 128, 0, 256]] and len(p.tiles) == 6


@pytest.fixture(scope="module")
def pair():
    hf = small_hf_model()
    cfg = Sam2Cfg.from_hf(hf.config)
    W = pack(hf.state_dict(), cfg, "cpu", act_dtype=torch.float32)
    return hf, cfg, Sam2Engine(cfg, W, TorchOps())


def test_trunk_neck_features_match_hf(pair):
    hf, cfg, eng = pair
    px = torch.randn(3, 256, 256, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        out = hf.get_image_features(px[None], return_dict=True)
        emb = hf(pixel_values=px[None], input_points=torch.tensor([[[[10.0, 20.0]]]]), input_labels=torch.tensor([[[1]]])).image_embeddings
    caps = []
    eng.set_image(px, capture=caps)
    s0, s1 = out.fpn_hidden_states[0][:, 0], out.fpn_hidden_states[1][:, 0]
    np.testing.assert_allclose(eng.feat_s0.numpy(), s0.numpy(), atol=2e-4, rtol=1e-3)
    np.testing.assert_allclose(eng.feat_s1.numpy(), s1.numpy(), atol=2e-4, rtol=1e-3)
    # image embedding: HF adds no_memory_embedding in forward, no_mask_embed (dense prompt) inside the decoder: both folded here
    want = emb[-1][0].flatten(1).t() + hf.prompt_encoder.no_mask_embed.weight
    np.testing.assert_allclose(eng.image_embed.numpy(), want.detach().numpy(), atol=2e-4, rtol=1e-3)


@pytest.mark.parametrize("prompt", ["point", "box", "box+points"])
def test_mask_logits_and_iou_match_hf(pair, prompt):
    hf, cfg, eng = pair
    px = torch.randn(3, 256, 256, generator=torch.Generator().manual_seed(4))
    eng.set_image(px)
    kw, coords, labels = {}, [], []
    if "box" in prompt:
        kw["input_boxes"] = torch.tensor([[[30.0, 40.0, 200.0, 180.0]]])
        coords += [[30.0, 40.0], [200.0, 180.0]]
        labels += [2, 3]
    if "point" in prompt:
        pts = [[100.0, 120.0], [50.0, 60.0]] if prompt != "point" else [[100.0, 120.0]]
        kw["input_points"] = torch.tensor([[pts]])
        kw["input_labels"] = torch.tensor([[[1, 0][:len(pts)]]])
        coords += pts
        labels += [1, 0][:len(pts)]
    coords.append([0.0, 0.0]); labels.append(-1)          # the padding "not a point" token both implementations append
    with torch.no_grad():
        out = hf(pixel_values=px[None], multimask_output=True, **kw)
    logits, iou = eng.predict_logits(np.asarray(coords, np.float32), np.asarray(labels, np.int64))
    np.testing.assert_allclose(logits.numpy(), out.pred_masks[0, 0].numpy(), atol=5e-4, rtol=2e-3)
    np.testing.assert_allclose(iou.numpy(), out.iou_scores[0, 0].numpy(), atol=1e-4)
