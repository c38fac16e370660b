"""Host-side layer program + fused algebra, validated on CPU: the drop-in modules run on the
test-only PyTorch emulation of the C-ABI ops (tests/emu_backend.py) and are compared with the
oracle (autograd) -- logits, loss and every parameter gradient."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import nets as onets
import pytorchdeeplearing_b200 as b200
from pytorchdeeplearing_b200 import runtime
from emu_backend import EmuBackend

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


@pytest.fixture(autouse=True)
def _emu():
    runtime._set_backend_for_testing(EmuBackend())
    prev = runtime.get_precision()
    runtime.set_precision("fp32")
    yield
    runtime._set_backend_for_testing(None)
    runtime.set_precision(prev)


def _build(kind, cin, ncls, seed):
    if kind == "vnet3d":
        spec = onets.vnet3d_state_spec(cin, ncls)
        model = b200.VNet3d(cin, ncls)
        ofwd = lambda sd, x, masks=None: onets.vnet3d_forward(sd, x, masks)
        draw = lambda n: onets.draw_dropout_masks_vnet3d(n)
    else:
        dims = 3 if kind == "unet3d" else 2
        spec = onets.unet_state_spec(cin, ncls, dims)
        model = (b200.UNet3d if dims == 3 else b200.UNet2d)(cin, ncls)
        ofwd = lambda sd, x, masks=None: onets.unet_forward(sd, x, dims, masks)
        draw = lambda n: onets.draw_dropout_masks_unet(n, dims)
    sd = onets.init_state_dict(spec, seed=seed, randomize_affine=True)
    assert [n for n, _ in spec] == list(model.state_dict().keys())
    model.load_state_dict(sd, strict=True)
    return spec, sd, model, ofwd, draw


CASES = [
    ("vnet3d", 1, 2, (16, 16, 16), 2, "MutilDiceLoss", b200.MutilDiceLoss),
    ("vnet3d", 1, 1, (16, 16, 16), 1, "BinaryCrossEntropyDiceLoss", b200.BinaryCrossEntropyDiceLoss),
    ("vnet3d", 2, 3, (16, 16, 16), 1, "MutilCrossEntropyDiceLoss", b200.MutilCrossEntropyDiceLoss),
    ("unet3d", 1, 4, (16, 16, 16), 1, "MutilCrossEntropyDiceLoss", b200.MutilCrossEntropyDiceLoss),
    ("unet2d", 1, 1, (32, 32), 2, "BinaryDiceFocalLoss", b200.BinaryDiceFocalLoss),
    ("unet2d", 3, 2, (16, 32), 2, "MutilFocalLoss", b200.MutilFocalLoss),
]


@pytest.mark.parametrize("kind,cin,ncls,spatial,n,lossname,losscls", CASES)
@pytest.mark.parametrize("train", [False, True])
def test_forward_backward_vs_oracle(kind, cin, ncls, spatial, n, lossname, losscls, train):
    spec, sd, model, ofwd, draw = _build(kind, cin, ncls, seed=5)
    x, y = oracle.make_inputs(n, cin, spatial, ncls, seed=77)
    alpha = torch.linspace(0.5, 1.5, ncls)
    masks = None
    if train:
        torch.manual_seed(3)
        masks = draw(n)
        model.train()
        model.dropout_masks = masks
    else:
        model.eval()
    # oracle
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo, po = ofwd(sdg, x, masks)
    loss_o = oracle.loss_forward(lossname, lo, y, alpha)
    loss_o.backward()
    # drop-in on the emulated backend
    if lossname.startswith("Mutil"):
        lossfn = losscls(alpha)
    else:
        lossfn = losscls()
    logits, probs = model(x)
    assert logits.shape == lo.shape and probs.shape == po.shape
    loss = lossfn(logits, y)
    loss.backward()
    assert (logits - lo).norm() / lo.norm() < 5e-6
    assert (probs - po).abs().max() < 1e-5
    assert abs(loss.item() - loss_o.item()) < 1e-5 * max(1, abs(loss_o.item()))
    for name, p in model.named_parameters():
        go = sdg[name].grad
        assert p.grad is not None, name
        err = (p.grad - go).norm() / (go.norm() + 1e-12)
        assert err < 2e-4, (name, err.item())


def test_dropout_draw_contract_matches_reference_stream():
    """Train-mode mask draws: same call order / shapes / RNG calls as nn.Dropout3d:
    reproduces the reference's own train-mode logits from the golden fixture."""
    from conftest import GOLDEN
    gold = dict(np.load(os.path.join(GOLDEN, "vnet3d_c2_16.npz")))
    spec = onets.vnet3d_state_spec(1, 2)
    sd = onets.init_state_dict(spec, seed=0, randomize_affine=True)
    model = b200.VNet3d(1, 2)
    model.load_state_dict(sd)
    model.train()
    x, _ = oracle.make_inputs(2, 1, (16, 16, 16), 2)
    torch.manual_seed(int(gold["train_seed"]))
    with torch.no_grad():
        logits, _ = model(x)
    ref = torch.from_numpy(gold["train_logits"])
    assert (logits - ref).norm() / ref.norm() < 5e-6


def test_golden_eval_logits_and_argmax():
    from conftest import GOLDEN
    for tag, ctor, shape, ncls, seed in (("vnet3d_c2_16", lambda: b200.VNet3d(1, 2), (2, 1, 16, 16, 16), 2, 0),
                                         ("unet3d_c4_16", lambda: b200.UNet3d(1, 4), (1, 1, 16, 16, 16), 4, 1),
                                         ("unet2d_c1_32", lambda: b200.UNet2d(1, 1), (2, 1, 32, 32), 1, 2)):
        gold = dict(np.load(os.path.join(GOLDEN, tag + ".npz")))
        model = ctor().eval()
        spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        model.load_state_dict(onets.init_state_dict(spec, seed=seed, randomize_affine=True))
        x, _ = oracle.make_inputs(shape[0], shape[1], shape[2:], ncls)
        with torch.no_grad():
            logits, probs = model(x)
        ref = torch.from_numpy(gold["logits"])
        assert (logits - ref).norm() / ref.norm() < 5e-6
        if ncls > 1:
            assert torch.equal(logits.argmax(1), ref.argmax(1))


def test_state_dict_roundtrip_and_initialize_weights():
    m = b200.VNet3d(1, 2)
    m.apply(b200.initialize_weights)
    sd = m.state_dict()
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert torch.all(sd["in_tr.bn1.weight"] == 1) and torch.all(sd["in_tr.conv1.bias"] == 0)
    m2 = b200.VNet3d(1, 2)
    m2.load_state_dict(sd, strict=True)
    assert len(list(m.parameters())) == 128


def test_cpu_tensor_without_backend_fails_loudly():
    runtime._set_backend_for_testing(None)
    m = b200.UNet2d(1, 1).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200.BinaryDiceLoss()(torch.zeros(1, 1, 16, 16), torch.zeros(1, 16, 16, dtype=torch.long))


def test_graphed_step_split_phases_match_autograd_path():
    """GraphedStep's split (data-parallel) mode drives the engine directly; it must give the same loss and
    gradients as the autograd path (world size 1 gloo group so the collectives are identities)."""
    import torch.distributed as dist
    from pytorchdeeplearing_b200.graphed import GraphedStep
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29700 + os.getpid() % 200))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        spec, sd, model, ofwd, draw = _build("vnet3d", 1, 2, seed=5)
        model.eval()
        x, y = oracle.make_inputs(2, 1, (16, 16, 16), 2, seed=77)
        lossfn = b200.MutilCrossEntropyDiceLoss(torch.linspace(0.5, 1.5, 2))
        logits, _ = model(x)
        loss = lossfn(logits, y)
        loss.backward()
        ref = {n: p.grad.clone() for n, p in model.named_parameters()}
        b200.enable_data_parallel()
        step = GraphedStep(model, lossfn, x, y, warmup=1, use_graph=False)
        loss2 = step(x, y)
        b200.disable_data_parallel()
        assert abs(loss2.item() - loss.item()) < 1e-6
        for n, p in model.named_parameters():
            assert (p.grad - ref[n]).norm() <= 1e-5 * (ref[n].norm() + 1e-12), n
    finally:
        b200.disable_data_parallel()
        dist.destroy_process_group()


def test_graphed_step_prefetch_matches_direct_inputs():
    """GraphedStep.prefetch + step(prefetched=True) (the pipelined input path of bench.py's e2e loop) feeds the same
    inputs as step(x, y); on CPU the staging is a plain copy."""
    from pytorchdeeplearing_b200.graphed import GraphedStep
    spec, sd, model, ofwd, draw = _build("vnet3d", 1, 2, seed=3)
    model.eval()                                       # no dropout draws: the two paths must agree exactly
    lossfn = b200.MutilDiceLoss(torch.ones(2))
    x0, y0 = oracle.make_inputs(1, 1, (16, 16, 16), 2, seed=11)
    x1, y1 = oracle.make_inputs(1, 1, (16, 16, 16), 2, seed=12)
    step = GraphedStep(model, lossfn, x0, y0, warmup=1, use_graph=False)
    l_direct = float(step(x1, y1).detach())
    g_direct = [p.grad.clone() for p in model.parameters()]
    step(x0, y0)
    step.prefetch(x1, y1)
    l_pref = float(step(prefetched=True).detach())
    assert l_pref == l_direct
    for a, p in zip(g_direct, model.parameters()):
        assert torch.equal(a, p.grad)
