"""Image preprocessing: the host-side restatement of PIL's 8-bit bicubic tables and of the HF processor's fused
normalisation, pinned on PIL / the HF processor themselves (CPU), and the device kernel against that reference path (GPU)."""
import numpy as np
import pytest
import torch

from oracle import host_ref
from socioreasoner_b200 import preprocess as P

SIZES = [(1024, 1024), (1000, 1300), (300, 310), (448, 448), (97, 1500), (756, 756)]


def _emulate_resize(img, oh, ow):
    """numpy emulation of the two integer passes of sr_preprocess_tile (horizontal, then vertical)."""
    x = img.astype(np.int64)
    half = 1 << (P.PIL_PRECISION_BITS - 1)
    if ow != x.shape[1]:
        b, k = P.pil_bicubic_coeffs(x.shape[1], ow)
        out = np.zeros((x.shape[0], ow, 3), np.int64)
        for xx in range(ow):
            s = np.full((x.shape[0], 3), half, np.int64)
            for j in range(b[xx, 1]):
                s += x[:, b[xx, 0] + j, :] * k[xx, j]
            out[:, xx, :] = np.clip(s >> P.PIL_PRECISION_BITS, 0, 255)
        x = out
    if oh != x.shape[0]:
        b, k = P.pil_bicubic_coeffs(x.shape[0], oh)
        out = np.zeros((oh, x.shape[1], 3), np.int64)
        for yy in range(oh):
            s = np.full((x.shape[1], 3), half, np.int64)
            for j in range(b[yy, 1]):
                s += x[b[yy, 0] + j, :, :] * k[yy, j]
            out[yy] = np.clip(s >> P.PIL_PRECISION_BITS, 0, 255)
        x = out
    return x.astype(np.uint8)


@pytest.mark.parametrize("hw", SIZES[:5])
def test_pil_tables_reproduce_pil_resize(hw):
    rng = np.random.default_rng(hw[0])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    oh, ow = P.smart_resize(*hw)
    _, _, resized = host_ref.reference_image_path([img])
    assert resized[0].shape == (oh, ow, 3)
    assert np.array_equal(_emulate_resize(img, oh, ow), resized[0])           # bit-exact with PIL's own resampler


def test_normalise_and_patch_layout_match_hf_processor():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (56, 84, 3), dtype=np.uint8)
    pv, thw, _ = host_ref.reference_image_path([img])
    mean, std = P.normalize_constants()
    v = (img.astype(np.float32) - mean) / std
    gh, gw = 4, 6
    x = np.stack([v, v], 0).transpose(0, 3, 1, 2).reshape(1, 2, 3, gh // 2, 2, 14, gw // 2, 2, 14)
    x = x.transpose(0, 3, 6, 4, 7, 2, 1, 5, 8).reshape(gh * gw, 1176)
    assert thw.tolist() == [[1, gh, gw]] and np.array_equal(x, pv.numpy())     # bit-exact


def test_smart_resize_matches_hf():
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize
    for h, w in SIZES + [(30, 4000), (2000, 2000), (57, 57)]:
        for mx in (768 * 768, 1344 * 1344):
            assert P.smart_resize(h, w, max_pixels=mx) == smart_resize(h, w, factor=28, min_pixels=56 * 56, max_pixels=mx)


@pytest.mark.gpu
@pytest.mark.parametrize("hw", SIZES)
def test_gpu_preprocess_bit_exact(hw):
    rng = np.random.default_rng(hw[1])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    ref, thw, _ = host_ref.reference_image_path([img])
    pp = P.GpuPreprocessor("cuda")
    got, g = pp([img])
    assert g.tolist() == thw.tolist()
    assert torch.equal(got.cpu(), ref)                                          # fp32 rows, bit-exact
    got16, _ = pp([torch.from_numpy(img).cuda()], out_dtype=torch.bfloat16)
    assert torch.equal(got16.cpu(), ref.to(torch.bfloat16))
    n = ref.shape[0]
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0)).int()
    moved, _ = pp([img], row_maps=[perm.cuda()])
    assert torch.equal(moved.cpu()[perm.long()], ref)                           # row r of the processor order lands in row perm[r]


@pytest.mark.gpu
def test_gpu_preprocess_batch_of_mixed_sizes():
    rng = np.random.default_rng(7)
    imgs = [rng.integers(0, 256, (s, s, 3), dtype=np.uint8) for s in (448, 1024, 896)]
    ref, thw, _ = host_ref.reference_image_path(imgs, max_pixels=1344 * 1344)
    got, g = P.GpuPreprocessor("cuda", max_pixels=1344 * 1344)(imgs)
    assert g.tolist() == thw.tolist() and torch.equal(got.cpu(), ref)
