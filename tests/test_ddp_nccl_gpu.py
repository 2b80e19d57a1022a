"""Data-parallel path on real GPUs (needs >= 2 devices; skipped otherwise): two NCCL ranks, each a full replica on its own
batch shard, gradients averaged by the bucketed all-reducer that TrainStep uses (started stage by stage from the segmentor's
explicit backward, on a side stream).  With BatchNorm in eval mode the shards are independent, so the averaged gradient must
equal the mean of the per-shard gradients computed by ONE process (SURVEY.md §4 iv, §8e)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _build(seed=0):
    from octave_b200 import config, network
    config.set_compute_dtype("bf16")
    # one writer per output in the split-K weight gradients: without it two evaluations of the SAME gradient already differ
    # by up to ~1e-3 on small-norm parameters (fp32 atomics), which is the size of the effect under test
    config.set_deterministic(True)
    torch.manual_seed(seed)
    net = network.ResnestUNet(2, False)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    return net.cuda().eval()


def _shard_grads(net, x, ys):
    from octave_b200 import losses
    for p in net.parameters():
        p.grad = None
    att, agg, _ = net(x.cuda())
    res = losses.FusedSegmentorLoss().total(agg, ys.cuda(), att, None, 1.0, 0.1, 0.0)
    res['total'].backward()
    return float(res['total'])


def _worker(rank, world, port, q, grad_dtype="fp32", tol=2e-3):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from octave_b200 import synth
        from octave_b200.train import GradAllReducer
        net = _build()
        shards = [synth.octa_batch(2, 96, 96, seed=40 + r, n_ridges=8)[:2] for r in range(world)]
        red = GradAllReducer(bucket_bytes=4 << 20, grad_dtype=grad_dtype)
        net._grad_ready_hook = lambda params, grads: red.reduce(params, grads)
        loss = _shard_grads(net, *shards[rank])
        red.finish()
        torch.cuda.synchronize()
        got = {n: p.grad.detach().float().clone() for n, p in net.named_parameters() if p.grad is not None}
        ok, worst = True, 0.0
        if rank == 0:
            # the same shards, one after the other, in this one process (no reducer): mean of the per-shard gradients
            net._grad_ready_hook = None
            ref = None
            for r in range(world):
                _shard_grads(net, *shards[r])
                cur = {n: p.grad.detach().float().clone() for n, p in net.named_parameters() if p.grad is not None}
                ref = cur if ref is None else {n: ref[n] + cur[n] for n in ref}
            for n in ref:
                a, b = got[n], ref[n] / world
                err = float((a - b).norm() / b.norm().clamp_min(1e-20))
                worst = max(worst, err)
            ok = set(got) == set(ref) and worst < tol
        q.put((rank, ok, worst, loss))
        dist.barrier()
    finally:
        dist.destroy_process_group()


# fp32 buckets: exact up to summation order (2e-3 on the worst parameter, measured 2e-4); bf16 buckets (the default of the
# bf16 product path): one bf16 rounding per rank + NCCL's bf16 sum, i.e. ~2^-8 relative -> 1e-2 on the worst parameter
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("grad_dtype,tol", [("fp32", 2e-3), ("bf16", 1e-2)])
def test_nccl_averaged_gradients_equal_single_process_mean_of_shards(grad_dtype, tol):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, grad_dtype, tol)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(timeout=120)
    print("rank results (rank, ok, worst relative L2 error of a parameter gradient, loss):", res)
    assert all(r[1] for r in res), res
    assert all(p.exitcode == 0 for p in ps)
