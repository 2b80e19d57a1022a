"""Does a bucketed all-reduce on a side stream capture into a CUDA graph and replay (2 ranks)?"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200.train import GradAllReducer
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "reducer"
params = [torch.nn.Parameter(torch.zeros(1 << 20, device=dev)) for _ in range(8)]
src = [torch.full((1 << 20,), float(rank + 1 + i), device=dev) for i in range(8)]
red = GradAllReducer(bucket_bytes=8 << 20)
def step():
    grads = [s * 2.0 for s in src]
    if mode == "plain":
        for g in grads: dist.all_reduce(g, op=dist.ReduceOp.AVG)
        return grads
    red.reduce(params[:4], grads[:4]); red.reduce(params[4:], grads[4:]); red.finish()
    return [p.grad for p in params]
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): out = step()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
print(rank, "eager ok", float(out[0][0]), flush=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = step()
torch.cuda.synchronize(); print(rank, "captured", flush=True)
for _ in range(3): g.replay()
torch.cuda.synchronize()
print(rank, "replayed", float(out[0][0]), float(out[7][0]), flush=True)
ctl = dist.new_group(backend="gloo")
dist.barrier(group=ctl); print(rank, "gloo barrier ok", flush=True)
t = torch.tensor([float(rank)]); dist.all_reduce(t, op=dist.ReduceOp.MAX, group=ctl); print(rank, "gloo max", float(t), flush=True)
for _ in range(2): g.replay()
torch.cuda.synchronize(); print(rank, "replayed again", float(out[0][0]), flush=True)
if len(sys.argv) > 2:
    e = torch.ones(4, device=dev); dist.all_reduce(e); torch.cuda.synchronize(); print(rank, "eager nccl after graph ok", float(e[0]), flush=True)
sys.stdout.flush(); os._exit(0)
