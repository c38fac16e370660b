"""Data-parallel path on real GPUs (NCCL, world size 2): sharded forward/backward with the loss partial-sum
all-reduce and the flat gradient all-reduce must reproduce the single-GPU result on the global batch.
Needs >= 2 CUDA devices (skipped otherwise).  `-m gpu`."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import oracle
    from oracle import nets as onets
    import pytorchdeeplearing_b200 as b200
    b200.set_precision("fp32")
    model = b200.VNet3d(1, 2)
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(onets.init_state_dict(spec, seed=4, randomize_affine=True))
    model = model.to(dev).eval()
    lossfn = b200.MutilCrossEntropyDiceLoss(torch.ones(2, device=dev))
    x, y = oracle.make_inputs(2 * world, 1, (32, 32, 32), 2, seed=9)
    x, y = x.to(dev), y.to(dev)
    logits, _ = model(x)
    loss_g = lossfn(logits, y)
    loss_g.backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = None
    b200.enable_data_parallel()
    sl = slice(2 * rank, 2 * rank + 2)
    logits, _ = model(x[sl])
    loss = lossfn(logits, y[sl])
    loss.backward()
    torch.cuda.synchronize()
    errs = sorted(((p.grad - ref[n]).norm() / (ref[n].norm() + 1e-12)).item() for n, p in model.named_parameters())
    b200.disable_data_parallel()
    # train mode through GraphedStep (captured, NCCL inside the graph when the stack allows it): every rank takes ITS
    # rows of the global-batch dropout draw (same seed everywhere), so the sharded step equals the single-GPU step on
    # the global batch -- loss and all-reduced gradients
    from pytorchdeeplearing_b200.graphed import GraphedStep
    model.train()
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(31)
    logits, _ = model(x)
    loss_t = lossfn(logits, y)
    loss_t.backward()
    ref_t = {n: p.grad.clone() for n, p in model.named_parameters()}
    b200.enable_data_parallel()
    step = GraphedStep(model, lossfn, x[sl], y[sl], warmup=1)
    torch.manual_seed(31)
    loss_s = step(x[sl], y[sl])
    torch.cuda.synchronize()
    errs_t = sorted(((p.grad - ref_t[n]).norm() / (ref_t[n].norm() + 1e-12)).item()
                    for n, p in model.named_parameters())
    ngraphs = len(step.graphs)
    del step
    torch.cuda.synchronize()
    b200.disable_data_parallel()
    q.put((rank, max(abs(loss.item() - loss_g.item()), abs(loss_s.item() - loss_t.item())),
           max(errs[len(errs) // 2], errs_t[len(errs_t) // 2]), max(errs[-1], errs_t[-1]), ngraphs))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_two_rank_matches_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, dloss, med, worst, ngraphs in res:
        assert dloss < 1e-5, (rank, dloss)
        assert med < 5e-3 and worst < 5e-2, (rank, med, worst)
        print(f"rank {rank}: graphs per step = {ngraphs} (collectives between the graphs; 1 = NCCL captured inside)")
