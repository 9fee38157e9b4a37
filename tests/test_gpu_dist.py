"""2-GPU NCCL data parallelism on real devices (skipped with fewer than 2 GPUs): bucketed all-reduce launched from
the backward stages gives every rank the SUM of the per-rank gradients, FusedAdamW applies 1/world, replicas stay
identical, and the result equals a single process that sees both batches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, sync):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), VITK_DP_SYNC=sync)
    try:
        from vision_transformers_torch_xla_b200 import optim_factory, utils
        from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
        from vision_transformers_torch_xla_b200.models import create_model
        from vision_transformers_torch_xla_b200.parallel import DataParallel

        assert utils.init_distributed_mode(None, backend="nccl")
        dev = torch.device("cuda", rank)
        torch.manual_seed(100 + rank)  # replicas start different; DataParallel broadcasts rank 0
        model = create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg").to(dev)

        class Args:
            opt, lr, weight_decay, opt_eps, opt_betas = "adamw", 1e-3, 0.05, 1e-8, None

        opt = optim_factory.create_optimizer(Args, model)
        dp = DataParallel(model, optimizer=opt)
        g = torch.Generator().manual_seed(7)
        xs = [torch.randn(4, 3, 224, 224, generator=g) for _ in range(world)]
        ys = [torch.softmax(torch.randn(4, 1000, generator=g) * 3, -1) for _ in range(world)]
        crit = SoftTargetCrossEntropy()
        dp.train()
        crit(dp(xs[rank].to(dev)), ys[rank].to(dev)).backward()
        dp.finish_gradient_sync()
        torch.cuda.synchronize()
        grad_sum = dp.store.grad.clone()
        # every rank holds the same summed gradient
        other = [torch.zeros_like(grad_sum) for _ in range(world)]
        dist.all_gather(other, grad_sum)
        assert all(torch.equal(other[0], o) for o in other)
        if rank == 0:
            # single-process reference: same weights, both batches, gradients accumulate
            ref = create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg").to(dev)
            ref.load_state_dict(model.state_dict())
            ref.train()
            for x, y in zip(xs, ys):
                crit(ref(x.to(dev)), y.to(dev)).backward()
            from vision_transformers_torch_xla_b200.store import get_store

            rg = get_store(ref).grad
            err = float((grad_sum - rg).abs().max() / rg.abs().max())
            assert err < 1e-3, err  # identical kernels; only the atomic accumulation order differs
        opt.step()
        opt.zero_grad()
        torch.cuda.synchronize()
        flat = [torch.zeros_like(dp.store.flat) for _ in range(world)]
        dist.all_gather(flat, dp.store.flat)
        assert all(torch.equal(flat[0], f) for f in flat), "replicas diverged after the optimizer step"
        assert opt.grad_scale == 1.0 / world
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("sync", ["step", "block", "tail:1"])
def test_data_parallel_two_gpus(sync):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, sync)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: "ok", 1: "ok"}, results
