"""Latency of the gradient exchange alone (torchrun, one process per GPU): the peer-memory all-reduce of csrc/dbw_comm.cu
(one-shot path at the scene-tensor payload, two-shot path at the leaf bucket) against ncclAllReduce, back to back on one
stream (lockstep: no rank skew in the number), eagerly and as 20 launches inside one CUDA graph."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import dbw_b200  # noqa: F401,E402
from dbw_b200.parallel import PeerAllReduce  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
CAP = 2_360_000
comm = PeerAllReduce(CAP, dev)
comm.flat.zero_()


def timed(fn, n=200):
    for _ in range(10):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


out = {}
for name, n in (('scene tensors (40k floats)', 40_000), ('leaf bucket (2.36M floats)', CAP)):
    buf = comm.flat[:n]
    other = torch.zeros(n, device=dev)
    out[name] = {'peer eager us': timed(lambda: comm.all_reduce(buf)), 'nccl eager us': timed(lambda: dist.all_reduce(other))}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        comm.all_reduce(buf)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            comm.all_reduce(buf)
    out[name]['peer in-graph us'] = timed(g.replay, 20) / 20
err = comm.error()
if rank == 0:
    print({'world': world, 'error': err, **out})
dist.barrier()
comm.close()
dist.destroy_process_group()
