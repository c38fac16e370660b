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
        print(f"{i:4d} {tc:18s} {str(sc[0]) if sc else '':28s} " + " ".join(f"{e:.2e}" for e in errs) + flag)
    print("---- parameter gradients (cuda vs emu)")
    for n in gc:
        e = ((gc[n] - ge[n]).norm() / (ge[n].norm() + 1e-300)).item()
        if e > (1e-4 if mode == "fp32" else 5e-2):
            print(f"{n:40s} {e:.3e}")


if __name__ == "__main__":
    main()
