"""One GPU, no NCCL: time the graph-replayed step of RANK 0's shard of the 49-view workload for world sizes 1,2,4,8
(7 views at N=8) -- the per-rank compute of the strong-scaling run, i.e. how much of a step does not shrink with the
shard (small launches, texture prep, zero-fills).  python scripts/shard_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from copy import deepcopy
import bench
import dbw_b200
from dbw_b200 import _lib
from dbw_b200.dbw import DifferentiableBlocksWorld
from dbw_b200.parallel import ViewParallel, shard_views
from dbw_b200.graph import GraphedStep

dev = torch.device('cuda:0')
W = bench.WORKLOAD
B, H, Wd, K = W['n_views'], W['height'], W['width'], W['faces_per_pixel']
host = bench.synthetic_inputs(B, H, Wd)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for world in (1, 2, 4, 8):
    torch.manual_seed(bench.SEED)
    model = DifferentiableBlocksWorld((H, Wd), **deepcopy(bench.MODEL_CFG)).to(dev); model.train()
    vp = ViewParallel(model, seed=bench.SEED)
    sl = shard_views(B, world, 0)
    loc = {k: v[sl].contiguous().to(dev) for k, v in host.items()}
    g = GraphedStep(vp, loc, B)
    for _ in range(5):
        g.run()
    torch.cuda.synchronize()
    evs = []
    for _ in range(50):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.run(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    _lib.lib().dbw_timing_reset(); _lib.lib().dbw_timing_enable(1)
    for _ in range(10):
        flush.zero_(); vp.forward_backward(loc, None, already_sharded=True, n_total_views=B)
    torch.cuda.synchronize(); _lib.lib().dbw_timing_enable(0)
    kt = sum(_lib.kernel_time_ms(kind, kk)[0] for kind in (0, 1) for kk in (1, K)) / 10
    _lib.lib().dbw_timing_reset()
    nv = sl.stop - sl.start
    if world == 1:
        base = B / ms
        try:                                            # GPU activities of one eager step, by name
            import collections
            from torch.profiler import profile, ProfilerActivity
            from torch.autograd import DeviceType
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                vp.forward_backward(loc, None, already_sharded=True, n_total_views=B)
                torch.cuda.synchronize()
            names = collections.Counter(e.name[:70] for e in prof.events() if e.device_type == DeviceType.CUDA)
            print(f'GPU activities in one eager step: {sum(names.values())}')
            for k, n in names.most_common(40):
                print(f'   {n:3d}  {k}')
        except Exception as exc:
            print('profiler unavailable:', exc)
    print(f'world {world}: {nv:2d} views on rank 0: step {ms:.3f} ms, raster kernels {kt:.3f} ms, rest {ms - kt:.3f} ms '
          f'-> {B / ms:.1f} k views/s if every rank took this long (x{(B / ms) / base:.2f})')
