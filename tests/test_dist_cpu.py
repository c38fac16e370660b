"""Data-parallel host logic on CPU (gloo, world_size 2) with the emulated backend: batch sharding
with (a) the all-reduce of the loss partial sums (global-batch-exact Dice) and
(b) the SUM all-reduce of the flat gradient bucket must reproduce the single-process result on the
global batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, arch, q, ov=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if ov is not None:
        os.environ["B200SEG_OVERLAP"] = ov           # side-stream scheduling options forced on under data parallelism
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import nets as onets
    import pytorchdeeplearing_b200 as b200
    from pytorchdeeplearing_b200 import runtime
    from emu_backend import EmuBackend
    runtime._set_backend_for_testing(EmuBackend())
    runtime.set_precision("fp32")
    if arch == "vnet3d":
        model, ncls, sp, lossfn = b200.VNet3d(1, 2), 2, (16, 16, 16), b200.MutilCrossEntropyDiceLoss(torch.ones(2))
    else:
        model, ncls, sp, lossfn = b200.UNet2d(1, 1), 1, (32, 32), b200.BinaryDiceFocalLoss()
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(onets.init_state_dict(spec, seed=4, randomize_affine=True))
    model.eval()
    x, y = oracle.make_inputs(2 * world, 1, sp, ncls, seed=9)
    # single-process result on the GLOBAL batch (no collectives)
    logits, _ = model(x)
    loss_g = lossfn(logits, y)
    loss_g.backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = None
    # sharded: rank r takes samples [2r, 2r+2)
    b200.enable_data_parallel()
    sl = slice(2 * rank, 2 * rank + 2)
    logits, _ = model(x[sl])
    loss = lossfn(logits, y[sl])
    loss.backward()
    b200.disable_data_parallel()
    err = max(((p.grad - ref[n]).norm() / (ref[n].norm() + 1e-12)).item() for n, p in model.named_parameters())
    # train mode: every rank takes ITS rows of the global-batch dropout draw (same seed on all ranks), and
    # GraphedStep (engine-driven, bucketed all-reduce) must agree with the autograd path on the global batch
    from pytorchdeeplearing_b200.graphed import GraphedStep
    model.train()
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(21)
    logits, _ = model(x)
    loss_t = lossfn(logits, y)
    loss_t.backward()
    ref_t = {n: p.grad.clone() for n, p in model.named_parameters()}
    b200.enable_data_parallel()
    torch.manual_seed(21)
    step = GraphedStep(model, lossfn, x[sl], y[sl], warmup=1, use_graph=False)
    b200.disable_data_parallel()
    err_t = max(((p.grad - ref_t[n]).norm() / (ref_t[n].norm() + 1e-12)).item() for n, p in model.named_parameters())
    q.put((rank, max(abs(loss.item() - loss_g.item()), abs(step.loss.item() - loss_t.item())), max(err, err_t)))
    dist.destroy_process_group()


@pytest.mark.parametrize("arch,ov", [("vnet3d", None), ("unet2d", None), ("vnet3d", "15")])
def test_two_rank_sharding_matches_global_batch(arch, ov):
    """``ov``: by default data-parallel runs keep the plain schedule (engine.overlap_mask); the forced case checks that
    the options' host logic (split pack, gradient bucket allocated and zeroed during forward, bucket flush in place of
    the early unpack) still reproduces the global-batch result"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if ov else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, arch, q, ov)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dloss, gerr in res:
        assert dloss < 1e-6, (rank, dloss)          # loss equals the global-batch loss on every rank
        assert gerr < 1e-4, (rank, gerr)            # summed gradients equal the global-batch gradients
