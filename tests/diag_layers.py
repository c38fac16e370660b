"""Diagnostic tool (not a test): run one fwd+bwd of a net with the CUDA backend and with the CPU
emulation, recording the outputs of every backend op, and print where they first / most diverge.

    python tests/diag_layers.py vnet3d 32 fp32
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from oracle import nets as onets  # noqa: E402
import pytorchdeeplearing_b200 as b200  # noqa: E402
from pytorchdeeplearing_b200 import runtime  # noqa: E402
from emu_backend import EmuBackend  # noqa: E402

OUT_ARGS = {
    "conv": [5, 6], "wgrad": [4], "gn_finalize": [7, 8], "apply": [5], "gn_bwd_reduce": [3],
    "gn_bwd_finalize": [6, 7, 8, 9], "gn_bwd_apply": [4], "colsum": [1], "pool_fwd": [1], "pool_bwd": [3],
    "head_probs": [1], "loss_partials": [4], "loss_finalize": [6, 7], "loss_bwd": [4], "unpack_wgrad": [1],
}


class Recorder:
    def __init__(self, inner):
        self.inner, self.log = inner, []

    def __getattr__(self, name):
        fn = getattr(self.inner, name)
        if name not in OUT_ARGS and name != "pack_weight":
            return fn

        def wrapped(*a, **k):
            r = fn(*a, **k)
            if name == "gn_bwd_apply":
                self.last_bwd_apply_inputs = [t.detach().cpu().clone() for t in a[:4]]
                self.bwd_apply_inputs = getattr(self, "bwd_apply_inputs", [])
                self.bwd_apply_inputs.append((len(self.log), self.last_bwd_apply_inputs))
            if name in OUT_ARGS:
                outs = [a[i] for i in OUT_ARGS[name] if i < len(a) and isinstance(a[i], torch.Tensor)]
                tag = name
                if name in ("conv", "wgrad"):
                    tag += f"[k{a[0]}]"
                self.log.append((tag, [o.detach().double().cpu().clone() for o in outs],
                                 [tuple(o.shape) for o in outs]))
            return r
        return wrapped


def run(backend, arch, size, mode, dev):
    runtime._set_backend_for_testing(backend)
    runtime.set_precision(mode)
    if arch == "vnet3d":
        model = b200.VNet3d(1, 2)
        x, y = oracle.make_inputs(2, 1, (size,) * 3, 2, seed=77)
        lossfn = b200.MutilDiceLoss(torch.linspace(0.5, 1.5, 2).to(dev))
    elif arch == "unet3d":
        model = b200.UNet3d(1, 4)
        x, y = oracle.make_inputs(1, 1, (size,) * 3, 4, seed=77)
        lossfn = b200.MutilCrossEntropyDiceLoss(torch.ones(4).to(dev))
    else:
        model = b200.UNet2d(1, 1)
        x, y = oracle.make_inputs(2, 1, (size,) * 2, 1, seed=77)
        lossfn = b200.BinaryDiceFocalLoss()
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(onets.init_state_dict(spec, seed=5, randomize_affine=True))
    model = model.to(dev).eval()
    logits, _ = model(x.to(dev))
    loss = lossfn(logits, y.to(dev))
    loss.backward()
    return {n: p.grad.detach().double().cpu() for n, p in model.named_parameters()}


def main():
    arch, size, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    from pytorchdeeplearing_b200._abi import CudaBackend
    rc = Recorder(CudaBackend())
    gc = run(rc, arch, size, mode, "cuda")
    torch.cuda.synchronize()
    re_ = Recorder(EmuBackend())
    ge = run(re_, arch, size, mode, "cpu")
    runtime._set_backend_for_testing(None)
    print(f"{len(rc.log)} ops (cuda) / {len(re_.log)} ops (emu)")
    for i, ((tc, oc, sc), (te, oe, se)) in enumerate(zip(rc.log, re_.log)):
        assert tc == te, (i, tc, te)
        errs = []
        for a, b in zip(oc, oe):
            errs.append(((a - b).norm() / (b.norm() + 1e-300)).item())
        worst = max(errs) if errs else 0.0
        flag = " <<<" if worst > (1e-4 if mode == "fp32" else 2e-2) else ""
        extra = ""
        if tc.startswith("gn_bwd_apply") and oc:
            d = (oc[0] - oe[0]).abs()
            big = int((d > 0.05 * oe[0].abs().max()).sum())
            extra = f"  [outliers>5%max: {big}, err^2 share of top-8: {(d.flatten().topk(8).values ** 2).sum() / (d ** 2).sum():.3f}]"
        print(f"{i:4d} {tc:18s} {str(sc[0]) if sc else '':28s} " + " ".join(f"{e:.2e}" for e in errs) + flag + extra)
    # kernel arithmetic in isolation: re-evaluate the first badly diverging gn_bwd_apply on the CUDA run's
    # own inputs with the emulation formula
    emu = EmuBackend()
    shown = 0
    for (idx, ins_c), (_, ins_e) in zip(rc.bwd_apply_inputs, re_.bwd_apply_inputs):
        oc, oe = rc.log[idx][1][0], re_.log[idx][1][0]
        err = ((oc - oe).norm() / oe.norm()).item()
        if err < 1e-4 or shown >= 2:
            continue
        shown += 1
        g, y, coef, coef3 = ins_c
        out = torch.empty_like(g)
        emu.gn_bwd_apply(g, y, coef, coef3, out)
        print(f"op {idx}: cuda-out vs emu(formula on cuda inputs): {((oc - out.double()).norm() / oc.norm()).item():.3e}")
        ge, ye, coefe, coef3e = ins_e
        n = y.shape[0]
        v = lambda t, i: t[..., i].view(n, 1, 1, 1, -1)
        pre_c = torch.addcmul(v(coef, 1), y.float(), v(coef, 0))
        pre_e = torch.addcmul(v(coefe, 1), ye.float(), v(coefe, 0))
        flips = (pre_c > 0) != (pre_e > 0)
        print(f"   mask flips between runs: {int(flips.sum())}; |pre| at flips: {pre_c[flips].abs().tolist()[:8]}")
        print(f"   exact zeros in pre-activation (cuda inputs): {int((pre_c == 0).sum())}; y==0: {int((y == 0).sum())}")
        d = (oc - oe).abs().flatten()
        top = d.topk(8)
        for val, ix in zip(top.values.tolist(), top.indices.tolist()):
            print(f"   idx {ix}: |diff| {val:.3e} cuda {oc.flatten()[ix]:.4e} emu {oe.flatten()[ix]:.4e} "
                  f"pre_c {pre_c.flatten()[ix]:.4e} pre_e {pre_e.flatten()[ix]:.4e} g {g.flatten()[ix]:.4e}")
    print("---- parameter gradients (cuda vs emu)")
    for n in gc:
        e = ((gc[n] - ge[n]).norm() / (ge[n].norm() + 1e-300)).item()
        if e > (1e-4 if mode == "fp32" else 5e-2):
            print(f"{n:40s} {e:.3e}")


if __name__ == "__main__":
    main()
