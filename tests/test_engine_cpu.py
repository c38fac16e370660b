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


# ----------------------------------------------------------------------------------------------- round 2 additions
def test_vnet2d_forward_backward_vs_oracle():
    """VNet2d (reference networks/VNet2d.py:102-160) through the same layer program with unit depth."""
    spec = onets.vnet3d_state_spec(1, 2, dims=2)
    sd = onets.init_state_dict(spec, seed=4, randomize_affine=True)
    model = b200.VNet2d(1, 2)
    assert [n for n, _ in spec] == list(model.state_dict().keys())
    model.load_state_dict(sd, strict=True)
    x, y = oracle.make_inputs(2, 1, (32, 48), 2, seed=9)
    torch.manual_seed(2)
    masks = onets.draw_dropout_masks_vnet3d(2, dims=2)
    model.train()
    model.dropout_masks = masks
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo, po = onets.vnet3d_forward(sdg, x, masks)
    alpha = torch.tensor([0.7, 1.3])
    loss_o = oracle.loss_forward("MutilCrossEntropyDiceLoss", lo, y, alpha)
    loss_o.backward()
    logits, probs = model(x)
    loss = b200.MutilCrossEntropyDiceLoss(alpha)(logits, y)
    loss.backward()
    assert logits.shape == lo.shape == (2, 2, 32, 48)
    assert (logits - lo).norm() / lo.norm() < 5e-6 and (probs - po).abs().max() < 1e-5
    assert abs(loss.item() - loss_o.item()) < 1e-5
    for name, p in model.named_parameters():
        go = sdg[name].grad
        assert (p.grad - go).norm() / (go.norm() + 1e-12) < 2e-4, name


def test_wide_net_takes_the_unfused_groupnorm_path():
    """init_features=64 -> 1024-channel bottleneck: beyond the fused-coefficient kernels' 512 channels the engine uses
    the finalize/apply form (the reference accepts any init_features)."""
    spec = onets.unet_state_spec(1, 1, 2, f=64)
    sd = onets.init_state_dict(spec, seed=1, randomize_affine=True)
    model = b200.UNet2d(1, 1, init_features=64)
    model.load_state_dict(sd, strict=True)
    model.eval()
    x, y = oracle.make_inputs(1, 1, (16, 16), 1, seed=3)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo, _ = onets.unet_forward(sdg, x, 2)
    loss_o = oracle.loss_forward("BinaryCrossEntropyDiceLoss", lo, y)
    loss_o.backward()
    logits, _ = model(x)
    loss = b200.BinaryCrossEntropyDiceLoss()(logits, y)
    loss.backward()
    assert (logits - lo).norm() / lo.norm() < 1e-5
    worst = max(((p.grad - sdg[n].grad).norm() / (sdg[n].grad.norm() + 1e-12)).item()
                for n, p in model.named_parameters())
    assert worst < 5e-4, worst


def test_autograd_contract_errors_are_explicit():
    spec, sd, model, ofwd, draw = _build("unet2d", 1, 2, seed=5)
    model.eval()
    x, y = oracle.make_inputs(1, 1, (16, 16), 2, seed=1)
    logits, probs = model(x)
    with pytest.raises(RuntimeError, match="second output"):
        (logits.sum() + probs.sum()).backward()              # gradient through probs: refused, not dropped
    logits, probs = model(x)
    loss = b200.MutilDiceLoss(torch.ones(2))(logits, y)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()
    with pytest.raises(RuntimeError, match="d/d\\(input\\)"):
        model(x.clone().requires_grad_(True))
    with pytest.raises(RuntimeError, match="fp32"):
        b200.UNet2d(1, 2).double()(x)
    with pytest.raises(RuntimeError, match="label"):
        bad = y.clone()
        bad[0, 0, 0] = 2
        b200.MutilCrossEntropyLoss(torch.ones(2))(model(x)[0], bad)


def test_metrics_match_reference_formulas():
    """dice_coeff / iou_coeff / multiclass_dice_coeff (reference model/metric.py:146-181) and the same numbers from the
    loss pass (``lossfn.last_dice()``)."""
    g = torch.Generator().manual_seed(0)

    def ref_dice(inp, tgt):
        inp = (inp > 0.5).float()
        num = tgt.size(0)
        a, b = inp.reshape(num, -1), tgt.reshape(num, -1).float()
        return ((2. * (a * b).sum(1) + 1e-5) / (a.sum(1) + b.sum(1) + 1e-5)).sum() / num

    def ref_iou(inp, tgt):
        inp = (inp > 0.5).float()
        num = tgt.size(0)
        a, b = inp.reshape(num, -1), tgt.reshape(num, -1).float()
        i = (a * b).sum(1)
        return ((i + 1e-5) / (a.sum(1) + b.sum(1) - i + 1e-5)).sum() / num

    z = 2 * torch.randn((3, 1, 8, 16, 16), generator=g)
    t = (torch.rand((3, 8, 16, 16), generator=g) > 0.6).long()
    p = torch.sigmoid(z)
    assert abs(b200.dice_coeff(p, t).item() - ref_dice(p, t).item()) < 1e-6
    assert abs(b200.iou_coeff(p, t).item() - ref_iou(p, t).item()) < 1e-6
    lf = b200.BinaryDiceLoss()
    lf(z, t)
    assert abs(lf.last_dice().item() - ref_dice(p, t).item()) < 1e-6
    zm = 2 * torch.randn((2, 4, 8, 8, 16), generator=g)
    tm = torch.randint(0, 4, (2, 8, 8, 16), generator=g)
    pm = torch.softmax(zm, 1)
    oh = torch.nn.functional.one_hot(tm, 4).permute(0, 4, 1, 2, 3)
    want = sum(ref_dice(pm[:, c], oh[:, c]) for c in range(1, 4)) / 3
    assert abs(b200.multiclass_dice_coeff(pm, tm).item() - want.item()) < 1e-6
    lm = b200.MutilCrossEntropyDiceLoss(torch.ones(4))
    lm(zm, tm)
    assert abs(lm.last_dice().item() - want.item()) < 1e-6


def test_predict_mask_and_sliding_window():
    """forward-only inference (reference predict, model/modelVNet.py:655-676): uint8 mask = argmax / threshold*255"""
    spec, sd, model, ofwd, draw = _build("unet3d", 1, 4, seed=2)
    x, _ = oracle.make_inputs(1, 1, (16, 16, 16), 4, seed=5)
    lo, po = ofwd(sd, x)
    mask = b200.predict(model, x[0].numpy())
    assert mask.dtype == np.uint8 and mask.shape == (16, 16, 16)
    assert np.array_equal(mask, po[0].argmax(0).numpy().astype(np.uint8))
    spec, sd, model, ofwd, draw = _build("unet2d", 1, 1, seed=2)
    x, _ = oracle.make_inputs(1, 1, (32, 32), 1, seed=5)
    lo, po = ofwd(sd, x)
    mask = b200.predict(model, x[0].numpy(), out_threshold=0.4)
    assert np.array_equal(mask, ((po[0, 0] > 0.4).numpy() * 255).astype(np.uint8))
    big = torch.randn(1, 32, 48)
    sw = b200.sliding_window_mask(model, big.numpy(), (16, 16), batch=3)
    want = np.zeros((32, 48), np.int64)
    for a in (0, 8, 16):
        for b in (0, 8, 16, 24, 32):
            _, pp = ofwd(sd, big[None, :, a:a + 16, b:b + 16])
            want[a:a + 16, b:b + 16] += (pp[0, 0] > 0.5).numpy()
    assert np.array_equal(sw, (want != 0).astype(np.uint8))


@pytest.mark.parametrize("cls,tcls,kw", [("FusedAdamW", torch.optim.AdamW, {}),
                                         ("FusedAdam", torch.optim.Adam, {"weight_decay": 0.05})])
def test_fused_adam_matches_torch_optim(cls, tcls, kw):
    """three zero_grad -> backward -> step iterations of the reference loop (model/modelVNet.py:590-593) with the fused
    optimizer vs torch.optim on the oracle"""
    spec, sd, model, ofwd, draw = _build("unet2d", 1, 1, seed=7)
    model.eval()
    x, y = oracle.make_inputs(2, 1, (16, 16), 1, seed=8)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    topt = tcls(list(sdg.values()), lr=1e-2, **kw)
    opt = getattr(b200, cls)(model.parameters(), lr=1e-2, **kw)
    lossfn = b200.BinaryCrossEntropyDiceLoss()
    for it in range(3):
        topt.zero_grad()
        lo, _ = ofwd(sdg, x)
        oracle.loss_forward("BinaryCrossEntropyDiceLoss", lo, y).backward()
        topt.step()
        opt.zero_grad()
        logits, _ = model(x)
        lossfn(logits, y).backward()
        opt.step()
    assert len(model.state_dict()) == 64                      # re-homed parameters keep the state_dict contract
    # three steps of size ~lr = 1e-2 each: agreement to ~1 % of the distance travelled (Adam's m/sqrt(v) amplifies the
    # fp32-level gradient differences between the two evaluations where a gradient is close to zero)
    for n, p in model.named_parameters():
        # (an element whose gradient is ~0 in one evaluation and exactly 0 in the other takes a full +-lr step in one
        #  of them: bound the mean and the 99th percentile, not the maximum)
        d = (p.detach() - sdg[n].detach()).abs().flatten()
        assert d.mean() < 3e-5 and d.kthvalue(max(1, int(0.99 * d.numel()))).values < 3e-4, (n, d.mean().item())
        assert (sd[n] - sdg[n].detach()).abs().mean() > 1e-3, n          # ... and they did move


@pytest.mark.parametrize("fused", [True, False])
def test_graphed_step_with_optimizer_updates_weights(fused):
    """GraphedStep(optimizer=...) must step the optimizer and keep p.grad bound to the static gradients even after the
    reference loop's ``zero_grad()`` (set_to_none) -- ADVICE r1: the optimizer used to be dropped silently."""
    from pytorchdeeplearing_b200.graphed import GraphedStep
    spec, sd, model, ofwd, draw = _build("unet2d", 1, 2, seed=7)
    model.eval()
    x, y = oracle.make_inputs(2, 1, (16, 16), 2, seed=8)
    alpha = torch.ones(2)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    topt = torch.optim.AdamW(list(sdg.values()), lr=1e-2)
    opt = b200.FusedAdamW(model.parameters(), lr=1e-2) if fused else torch.optim.AdamW(model.parameters(), lr=1e-2)
    lossfn = b200.MutilCrossEntropyDiceLoss(alpha)
    step = GraphedStep(model, lossfn, x, y, warmup=1, optimizer=opt, use_graph=False)    # warm-up = 1 real step
    losses = []
    for it in range(3):
        topt.zero_grad()
        lo, _ = ofwd(sdg, x)
        lo_loss = oracle.loss_forward("MutilCrossEntropyDiceLoss", lo, y, alpha)
        lo_loss.backward()
        topt.step()
        losses.append(lo_loss.item())
        if it > 0:                                              # the warm-up was step 0
            opt.zero_grad()                                     # set_to_none: p.grad must come back
            l = step(x, y)
            assert abs(l.item() - lo_loss.item()) < 2e-5
            assert all(p.grad is not None for p in model.parameters())
    assert losses[2] < losses[0]
    for n, p in model.named_parameters():
        d = (p.detach() - sdg[n].detach()).abs().flatten()
        assert d.mean() < 3e-5 and d.kthvalue(max(1, int(0.99 * d.numel()))).values < 3e-4, (n, d.mean().item())
    assert abs(step.dice.item() - b200.multiclass_dice_coeff(torch.softmax(lo, 1), y).item()) < 1e-6


def test_input_stager_host_logic_on_the_emulated_backend():
    """InputStager (staging.py): slot rotation, shape/dtype checks, label handling -- arithmetic by the emulated ops,
    checked against the oracle restatement of model/dataset.py:138-157"""
    import numpy as np
    from oracle import staging as ostaging
    from pytorchdeeplearing_b200.staging import InputStager, stage_batch
    rng = np.random.default_rng(3)
    st = InputStager("cpu", (3, 16, 24), backend=EmuBackend())
    batches = [(rng.integers(0, 256, (3, 16, 24), dtype=np.uint8), (rng.random((3, 16, 24)) > 0.6).astype(np.uint8) * 255)
               for _ in range(5)]
    st.put(*batches[0])
    st.put(*batches[1])
    with pytest.raises(RuntimeError):
        st.put(*batches[2])                       # both slots hold untaken batches
    for i in range(5):
        x, y = st.get()
        assert x.shape == (3, 1, 16, 24) and x.dtype == torch.float32 and y.dtype == torch.int64
        assert (x - ostaging.zscore_u8(batches[i][0])).abs().max() < 1e-6
        assert torch.equal(y, ostaging.labels_from_u8(batches[i][1]))
        st.release()
        if i + 2 < 5:
            st.put(*batches[i + 2])
    with pytest.raises(RuntimeError):
        st.get()
    with pytest.raises(ValueError):
        st.put(batches[0][0].astype(np.float32))
    with pytest.raises(ValueError):
        st.put(batches[0][0][:2])
    st.put(batches[0][0])                         # images only
    x, y = st.get()
    assert y is None
    x2, y2 = stage_batch(torch.from_numpy(batches[1][0]), torch.from_numpy(batches[1][1]), binarize_labels=False,
                         backend=EmuBackend())
    assert torch.equal(y2, torch.from_numpy(batches[1][1]).long()) and (x2 - ostaging.zscore_u8(batches[1][0])).abs().max() < 1e-6


@pytest.mark.parametrize("kind,cin,ncls,spatial,lossname,losscls", [
    ("vnet3d", 1, 2, (16, 16, 16), "MutilDiceLoss", b200.MutilDiceLoss),
    ("unet2d", 1, 1, (32, 32), "BinaryDiceFocalLoss", b200.BinaryDiceFocalLoss),
])
def test_overlap_options_change_launch_grouping_not_results(kind, cin, ncls, spatial, lossname, losscls, monkeypatch):
    """B200SEG_OVERLAP only regroups launches (two pack launches, an early unpack) and moves them between streams; on
    the emulated backend (no streams) every mask must give bit-identical logits, loss and gradients."""
    from pytorchdeeplearing_b200 import engine
    spec, sd, model, ofwd, draw = _build(kind, cin, ncls, seed=9)
    x, y = oracle.make_inputs(2, cin, spatial, ncls, seed=21)
    torch.manual_seed(4)
    model.train()
    model.dropout_masks = draw(2)
    lossfn = losscls(torch.ones(ncls)) if lossname.startswith("Mutil") else losscls()
    be = runtime._TEST_BACKEND
    calls = {"pack": 0, "unpack": 0}
    pack0, unpack0 = be.pack_many, be.unpack_many
    monkeypatch.setattr(be, "pack_many", lambda reqs: (calls.__setitem__("pack", calls["pack"] + 1), pack0(reqs))[1])
    monkeypatch.setattr(be, "unpack_many",
                        lambda items: (calls.__setitem__("unpack", calls["unpack"] + 1), unpack0(items))[1])
    results = {}
    for mask in (0, 15):
        monkeypatch.setenv("B200SEG_OVERLAP", str(mask))
        assert engine.overlap_mask() == mask
        calls["pack"] = calls["unpack"] = 0
        for p in model.parameters():
            p.grad = None
        logits, _ = model(x)
        loss = lossfn(logits, y)
        loss.backward()
        results[mask] = (logits.detach().clone(), loss.detach().clone(),
                         [p.grad.clone() for p in model.parameters()], dict(calls))
    # mask 15: a small pack launch + a planned (deferred) one, an early unpack + the final one
    assert results[0][3] == {"pack": 1, "unpack": 1}
    assert results[15][3] == {"pack": 2, "unpack": 2} and be.deferred_pack_launches == 1
    assert torch.equal(results[0][0], results[15][0]) and torch.equal(results[0][1], results[15][1])
    for g0, g1 in zip(results[0][2], results[15][2]):
        assert torch.equal(g0, g1)


def test_overlap_defaults_follow_the_measurements(monkeypatch):
    from pytorchdeeplearing_b200 import engine
    monkeypatch.delenv("B200SEG_OVERLAP", raising=False)
    assert engine.overlap_mask("vnet", 3) == engine.OV_ALL and engine.overlap_mask("unet", 2) == engine.OV_ALL
    assert engine.overlap_mask("unet", 3) == 0               # measured slower on UNet3d 128^3
    monkeypatch.setenv("B200SEG_OVERLAP", "5")
    assert engine.overlap_mask("unet", 3) == 5
