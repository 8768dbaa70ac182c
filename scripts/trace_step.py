"""Which op launches which kernel in one eager training step (dtu shape, 4 views): prints, in launch order, every CUDA kernel
with the innermost aten / autograd op that issued it -- the list the launch-count work in DESIGN.md section 6 went through."""
import os
import sys
from copy import deepcopy

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402
import dbw_b200  # noqa: F401,E402
from dbw_b200.dbw import DifferentiableBlocksWorld  # noqa: E402
from dbw_b200.parallel import ViewParallel  # noqa: E402

w = dict(bench.WORKLOADS['dtu'], n_views=4)
dev = torch.device('cuda:0')
torch.manual_seed(bench.SEED)
model = DifferentiableBlocksWorld((w['height'], w['width']), **deepcopy(bench.model_cfg(w))).to(dev)
model.train()
vp = ViewParallel(model, seed=bench.SEED)
inp = {k: v.to(dev) for k, v in bench.synthetic_inputs(w).items()}
for _ in range(3):
    vp.forward_backward(inp)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    vp.forward_backward(inp)
    torch.cuda.synchronize()
evs = prof.events()
cpu = [e for e in evs if e.device_type == torch.autograd.DeviceType.CPU]
kernels = sorted((e for e in evs if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
print(f'{len(kernels)} device activities in one step')
for k in kernels:
    # the innermost CPU op whose interval contains the kernel's launch (correlated by the profiler: k.linked_correlation... fallback: name)
    print(f'{k.time_range.elapsed_us():8.1f} us  {k.name[:90]}')
print()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=70))
