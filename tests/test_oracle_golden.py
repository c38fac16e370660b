"""Pin the CPU oracle (oracle/) against fixtures produced by the real reference
(tests/golden/make_golden.py).  CPU-only; runs on every `-m "not gpu"` pass."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import nets as onets
from conftest import GOLDEN

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def _load(tag):
    return dict(np.load(os.path.join(GOLDEN, tag + ".npz")))


def _weights_ok(sd, gold):
    w = torch.cat([v.double().flatten() for v in sd.values()])
    fp = np.array([w.sum().item(), w.abs().sum().item(), (w * w).sum().item()])
    return np.allclose(fp, gold["weights_fp"], rtol=1e-9, atol=1e-9)


NET_CASES = [
    # tag, kind, (cin, ncls), seed, N, spatial, loss, masks
    ("vnet3d_c2_16", "vnet3d", (1, 2), 0, 2, (16, 16, 16), "MutilDiceLoss", True),
    ("vnet3d_c2_32", "vnet3d", (1, 2), 0, 1, (32, 32, 32), "MutilCrossEntropyDiceLoss", False),
    ("vnet3d_c1_16", "vnet3d", (1, 1), 3, 1, (16, 16, 16), "BinaryDiceLoss", False),
    ("unet3d_c4_16", "unet3d", (1, 4), 1, 1, (16, 16, 16), "MutilCrossEntropyDiceLoss", True),
    ("unet2d_c1_32", "unet2d", (1, 1), 2, 2, (32, 32), "BinaryDiceFocalLoss", True),
    ("unet2d_c1_128", "unet2d", (1, 1), 2, 2, (128, 128), "BinaryDiceLoss", False),
]


def build_case(kind, chans, seed):
    cin, ncls = chans
    if kind == "vnet3d":
        spec = onets.vnet3d_state_spec(cin, ncls)
        fwd = lambda sd, x, masks=None: onets.vnet3d_forward(sd, x, masks)
        draw = lambda n: onets.draw_dropout_masks_vnet3d(n)
    else:
        dims = 3 if kind == "unet3d" else 2
        spec = onets.unet_state_spec(cin, ncls, dims)
        fwd = lambda sd, x, masks=None: onets.unet_forward(sd, x, dims, masks)
        draw = lambda n: onets.draw_dropout_masks_unet(n, dims)
    sd = onets.init_state_dict(spec, seed=seed, randomize_affine=True)
    return spec, sd, fwd, draw


@pytest.mark.parametrize("tag,kind,chans,seed,n,spatial,lossname,has_masks", NET_CASES)
def test_network_matches_reference(tag, kind, chans, seed, n, spatial, lossname, has_masks):
    gold = _load(tag)
    spec, sd, fwd, draw = build_case(kind, chans, seed)
    if not _weights_ok(sd, gold):
        pytest.skip("torch CPU RNG stream differs from the fixture generator's")
    x, y = oracle.make_inputs(n, chans[0], spatial, chans[1])
    assert np.allclose(gold["x_fp"][0], x.double().sum().item(), rtol=1e-9)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits, probs = fwd(sdg, x)
    alpha = torch.ones(chans[1])
    loss = oracle.loss_forward(lossname, logits, y, alpha)
    loss.backward()
    # the same algorithm on the same torch build: agreement to fp32 round-off
    if "logits" in gold:
        ref = torch.from_numpy(gold["logits"])
        rel = (logits.detach() - ref).norm() / ref.norm()
        assert rel < 2e-6, rel
        if chans[1] > 1:
            assert torch.equal(logits.argmax(1), ref.argmax(1))
    assert abs(loss.item() - float(gold["loss"])) < 2e-6 * max(1.0, abs(float(gold["loss"])))
    names = [nm for nm, _ in spec]
    gn = np.array([sdg[nm].grad.double().norm().item() for nm in names])
    assert np.allclose(gn, gold["grads"][:, 0], rtol=2e-4, atol=1e-7)
    amax = (probs > 0.5).sum() if chans[1] == 1 else probs.argmax(1).sum()
    assert int(amax) == int(gold["argmax_sum"])
    if has_masks:
        torch.manual_seed(int(gold["train_seed"]))
        masks = draw(n)
        lt, _ = fwd(sd, x, masks)
        ref = torch.from_numpy(gold["train_logits"])
        assert (lt - ref).norm() / ref.norm() < 2e-6


def test_state_spec_counts():
    spec = onets.vnet3d_state_spec(1, 2)
    assert len(spec) == 128
    assert sum(int(np.prod(s)) for _, s in spec) == 9492658
    spec = onets.unet_state_spec(1, 4, 3)
    assert len(spec) == 64 and sum(int(np.prod(s)) for _, s in spec) == 5646436
    spec = onets.unet_state_spec(1, 1, 2)
    assert len(spec) == 64 and sum(int(np.prod(s)) for _, s in spec) == 1942289
    assert len(onets.vnet3d_mask_channels()) == 34
    assert len(onets.unet_mask_channels()) == 18


LOSS_CASES = [
    ("BinaryDiceLoss", "BinaryDiceLoss", "b", None),
    ("BinaryCrossEntropyLoss", "BinaryCrossEntropyLoss", "b", None),
    ("BinaryFocalLoss", "BinaryFocalLoss", "b", None),
    ("BinaryCrossEntropyDiceLoss", "BinaryCrossEntropyDiceLoss", "b", None),
    ("MutilDiceLoss", "MutilDiceLoss", "m", None),
    ("MutilDiceLoss_absent", "MutilDiceLoss", "ma", None),
    ("MutilCrossEntropyLoss", "MutilCrossEntropyLoss", "m", None),
    ("MutilCrossEntropyLoss_absent", "MutilCrossEntropyLoss", "ma", None),
    ("MutilFocalLoss", "MutilFocalLoss", "m", 2),
    ("MutilFocalLoss_g3", "MutilFocalLoss", "m", 3),
    ("MutilCrossEntropyDiceLoss", "MutilCrossEntropyDiceLoss", "m", None),
]


@pytest.mark.parametrize("key,name,which,gamma", LOSS_CASES)
def test_loss_matches_reference(key, name, which, gamma):
    gold = _load("losses")
    if which == "b":
        z, t = torch.from_numpy(gold["zb"]), torch.from_numpy(gold["tb"])
    else:
        z = torch.from_numpy(gold["zm"])
        t = torch.from_numpy(gold["tm"] if which == "m" else gold["tm_absent"])
    z = z.clone().requires_grad_(True)
    alpha = torch.from_numpy(gold["alpha"])
    v = oracle.loss_forward(name, z, t, alpha, gamma)
    v.backward()
    assert abs(v.item() - float(gold[key + "_value"])) < 3e-6
    ref = torch.from_numpy(gold[key + "_grad"])
    assert (z.grad - ref).abs().max() < 1e-8 + 1e-4 * ref.abs().max()


def test_input_staging_restatement_equals_reference_dataset():
    """oracle/staging.py against what the reference's datasetModelSegwithopencv returned for the same 8-bit files
    (tests/golden/make_golden_staging.py): bit-equal images, equal labels"""
    from oracle import staging
    gold = _load("staging")
    x = staging.zscore_u8(gold["images_u8"])
    assert x.dtype == torch.float32 and tuple(x.shape) == gold["x"].shape
    assert torch.equal(x, torch.from_numpy(gold["x"]))
    y = staging.labels_from_u8(gold["labels_u8"])
    assert y.dtype == torch.int64 and torch.equal(y, torch.from_numpy(gold["y"]))
