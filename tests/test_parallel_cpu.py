"""CPU, world_size 2 over gloo: the gradient bucketer averages exactly like a single process on the concatenated batch,
and fedavg_allreduce reproduces fedavg_aggregate."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_model():
    torch.manual_seed(3)
    return nn.Sequential(nn.Linear(6, 16), nn.GELU(), nn.LayerNorm(16), nn.Linear(16, 4))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssl_vit_video_analytics_b200.federated.fed_loop import fedavg_allreduce
    from ssl_vit_video_analytics_b200.parallel import GradientBucketer

    model = _make_model()
    if rank == 1:  # ranks start from different weights: the bucketer must broadcast rank 0's
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    ddp = GradientBucketer(model, bucket_mb=0.0002, exclude=())
    assert len(ddp.buckets) >= 2
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 4, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for _ in range(2):  # second iteration exercises zero_grad / re-arming
        ddp.zero_grad()
        loss = ((ddp(xs) - ys) ** 2).mean()
        loss.backward()
        ddp.finish_gradient_sync()
    grads = [p.grad.clone() for p in model.parameters()]
    # deferred communication (the CUDA-graph mode): backward only fills the buckets, allreduce_now() averages them
    ddp.defer_comm = True
    ddp.zero_grad()
    ((ddp(xs) - ys) ** 2).mean().backward()
    ddp.finish_gradient_sync()
    local_only = [p.grad.clone() for p in model.parameters()]
    ddp.allreduce_now()
    for p, a in zip(model.parameters(), grads):
        assert torch.allclose(p.grad, a, rtol=1e-6, atol=1e-8)
    assert any(not torch.allclose(l, a, rtol=1e-3, atol=1e-6) for l, a in zip(local_only, grads))   # nothing was sent before
    ddp.defer_comm = False
    # FedAvg: rank r plays client r with weight r + 1
    fmodel = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4))
    torch.manual_seed(100 + rank)
    for p in fmodel.parameters():
        nn.init.normal_(p)
    fmodel[1].num_batches_tracked.fill_(5 + rank)
    before = {k: v.clone() for k, v in fmodel.state_dict().items()}
    fedavg_allreduce(fmodel, client_weight=float(rank + 1))
    torch.save({"grads": grads, "fed_before": before, "fed_after": {k: v.clone() for k, v in fmodel.state_dict().items()}},
               os.path.join(tmp, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_bucketed_allreduce_and_fedavg_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    # single-process reference on the full batch
    model = _make_model()
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 4, generator=g)
    ((model(x) - y) ** 2).mean().backward()
    for p, a, b in zip(model.parameters(), r0["grads"], r1["grads"]):
        assert torch.equal(a, b)                                  # both ranks hold the same averaged gradient
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-7)    # == gradient of the global-batch mean loss
    # FedAvg vs the drop-in aggregate
    from ssl_vit_video_analytics_b200.federated.fed_loop import fedavg_aggregate

    glob = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4))
    ref = fedavg_aggregate(glob, [r0["fed_before"], r1["fed_before"]], [1.0, 2.0])
    for k, v in ref.items():
        assert torch.equal(r0["fed_after"][k], r1["fed_after"][k]), k
        if torch.is_floating_point(v):
            assert torch.allclose(r0["fed_after"][k], v, rtol=1e-6, atol=1e-7), k
        else:
            assert torch.equal(r0["fed_after"][k], v), k
    assert int(r0["fed_after"]["1.num_batches_tracked"]) == 6


def test_bucketer_detects_missing_gradient():
    from ssl_vit_video_analytics_b200.parallel import GradientBucketer

    model = nn.ModuleDict({"used": nn.Linear(3, 3), "unused": nn.Linear(3, 3)})
    ddp = GradientBucketer(model, exclude=(), broadcast=False)
    ddp.zero_grad()
    model["used"](torch.randn(2, 3)).sum().backward()
    with pytest.raises(RuntimeError, match="received no gradient"):
        ddp.finish_gradient_sync()
    ddp2 = GradientBucketer(nn.ModuleDict({"used": nn.Linear(3, 3), "unused": nn.Linear(3, 3)}), exclude=("unused.",), broadcast=False)
    ddp2.zero_grad()
    ddp2.module["used"](torch.randn(2, 3)).sum().backward()
    ddp2.finish_gradient_sync()
