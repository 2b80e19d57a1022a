"""N>1 host logic on CPU: world_size-2 gloo run of the bucketed gradient all-reducer used by TrainStep (the data path
has no other collective: batch shards are independent, BatchNorm statistics stay per replica — SURVEY.md §8e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from octave_b200.train import GradAllReducer
    torch.manual_seed(0)
    shapes = [(64, 32, 3, 3), (64,), (128, 64, 1, 1), (2, 32), (1024, 512, 3, 3), (7,)]
    base = [torch.randn(s) for s in shapes]
    grads = [b * (rank + 1) for b in base]          # rank r holds (r+1) * base
    red = GradAllReducer(bucket_bytes=1 << 16)      # small buckets: several flushes, one oversized tensor
    red.reduce(grads[:2]); red.reduce([None] + grads[2:4]); red.reduce(grads[4:])
    red.finish()
    mean = sum(range(1, world + 1)) / world
    ok = all(torch.allclose(g, b * mean, rtol=1e-6, atol=1e-6) for g, b in zip(grads, base))
    # second use of the same reducer (next step) must start clean
    g2 = [torch.full((10,), float(rank))]
    red.reduce(g2); red.finish()
    ok = ok and torch.allclose(g2[0], torch.full((10,), (world - 1) / 2))
    # parameter-owning form: .grad is re-pointed at the averaged bucket slice (what TrainStep uses)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    g3 = [b * (rank + 1) for b in base]
    red.reduce(params[:3], g3[:3]); red.reduce(params[3:], g3[3:]); red.finish()
    ok = ok and all(p.grad is not None and p.grad.shape == p.shape and torch.allclose(p.grad, b * mean, rtol=1e-6, atol=1e-6)
                    for p, b in zip(params, base))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_process_is_a_no_op():
    from octave_b200.train import GradAllReducer
    red = GradAllReducer()
    g = [torch.ones(3)]
    red.reduce(g); red.finish()
    assert torch.equal(g[0], torch.ones(3))


def test_single_process_reducer_still_owns_param_grad():
    """With a grad-ready hook installed the segmentor's autograd node hands its gradients to the hook only (returning them
    to autograd as well would make AccumulateGrad clone every tensor the reducer references): a reducer of world size 1
    must therefore assign param.grad itself."""
    from octave_b200.train import GradAllReducer
    red = GradAllReducer()
    p = [torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4))]
    g = [torch.full((2, 3), 2.0), torch.full((4,), 3.0)]
    red.reduce(p, g); red.finish()
    assert torch.equal(p[0].grad, g[0]) and torch.equal(p[1].grad, g[1])


def test_reducer_rejects_unknown_bucket_dtype():
    import pytest
    from octave_b200.train import GradAllReducer
    with pytest.raises(ValueError, match="grad_dtype"):
        GradAllReducer(grad_dtype="fp8")
