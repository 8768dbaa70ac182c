#!/usr/bin/env python
"""bench.py -- render+backward views/sec of the DBW render hot path (BASELINE.json metric).

A "step" = one optimisation step's render work on one batch of B synthetic views: build the scene from the leaf
parameters, render the environment pass (K=1, sigma=0) and the blocks pass (K=10, sigma=1e-4, per-face opacities),
composite + MSE against the target images, and back-propagate to every leaf parameter
(S, R_6d, T, sq_eps, alpha_logit, textures, texture_bkg, texture_ground, R_6d_ground, T_ground).  SURVEY.md 8d.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (torchrun launches N ranks for N > 1)
  python bench.py --impl reference ...                            the reference's algorithm on the host cores (oracle port)
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = {'name': 'dtu_scan24_shape', 'n_blocks': 10, 'n_views': 49, 'height': 400, 'width': 400,
            'faces_per_pixel': 10, 'txt_size': 256}
SEED = 227391          # configs/dtu/default.yml:42

MODEL_CFG = {
    'mesh': {'n_blocks': WORKLOAD['n_blocks'], 'S_world': 0.5, 'R_world': [115, 0, 0], 'txt_size': WORKLOAD['txt_size']},
    'renderer': {'faces_per_pixel': WORKLOAD['faces_per_pixel'], 'cameras': {'name': 'perspective'}, 'detach_bary': True,
                 'z_clip': 0.001},
    'rend_optim': {'coarse_learning': 1500, 'decimate_txt': 750, 'decimate_factor': 8, 'kill_blocks': True,
                   'decouple_rendering': True, 'opacity_noise': True},
    'loss': {'rgb_weight': 1},
}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this very
# workload at N=1 (profiles/ncu_full_r1f_summary.txt); None for other shard sizes
NCU_TRAFFIC_BYTES = {'raster_forward[env K=1]': 108.9e6, 'raster_forward[blocks K=10]': 1402.5e6,
                     'raster_backward[blocks K=10]': 1001.7e6, 'raster_backward[env K=1]': 162.3e6}


def algorithmic_bytes_per_view(K, H, W):
    """RGBA out + grad RGBA in + int32 top-K ids written forward and read backward = H*W*(32 + 8K)."""
    return H * W * (32 + 8 * K)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line): one background
    `nvidia-smi -lms 100` process, started before the region and stopped after it."""
    QUERY = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)                      # let the first sample land before the timed region starts
        except Exception:
            self.proc = None

    def summary(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ''
        rows = [[c.strip() for c in l.split(',')] for l in out.strip().splitlines() if l.count(',') >= 6]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        num = lambda x: float(x) if x.replace('.', '', 1).isdigit() else None
        sm = sorted(v for v in (num(r[0]) for r in rows) if v is not None)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in rows)]
        pw = [v for v in (num(r[2]) for r in rows) if v is not None]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': num(rows[0][1]), 'reasons': reasons,
                'samples': len(rows), 'power_w_max': max(pw) if pw else None}


def synthetic_inputs(B, H, W, device=None, pin=False):
    from dbw_b200.synthetic import ring_cameras
    R, T, K = ring_cameras(B)
    g = torch.Generator().manual_seed(SEED)
    imgs = torch.rand(B, 3, H, W, generator=g)
    inp = {'imgs': imgs, 'R': R, 'T': T, 'K': K[None].expand(B, -1, -1).contiguous()}
    if pin:
        inp = {k: v.pin_memory() for k, v in inp.items()}
    if device is not None:
        inp = {k: v.to(device) for k, v in inp.items()}
    return inp


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    assert world == args.gpus or world == 1, f'WORLD_SIZE={world} but --gpus {args.gpus}'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import dbw_b200
    from dbw_b200 import _lib
    from dbw_b200.dbw import DifferentiableBlocksWorld
    from dbw_b200.parallel import ViewParallel, shard_views
    from dbw_b200.graph import GraphedStep, PipelinedGraphedStep
    from copy import deepcopy

    B, H, W, K = WORKLOAD['n_views'], WORKLOAD['height'], WORKLOAD['width'], WORKLOAD['faces_per_pixel']
    torch.manual_seed(SEED)
    model = DifferentiableBlocksWorld((H, W), **deepcopy(MODEL_CFG)).to(dev)
    model.train()
    vp = ViewParallel(model, seed=SEED)
    host = synthetic_inputs(B, H, W, pin=True)
    sl = shard_views(B, world, rank)
    host_local = {k: v[sl].contiguous().pin_memory() for k, v in host.items()}
    B_local = sl.stop - sl.start
    dev_local = {k: v.to(dev) for k, v in host_local.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step_eager():
        return vp.forward_backward(dev_local, None, already_sharded=True, n_total_views=B)

    car = args.allreduce_in_graph
    graphed = None if args.no_graph else GraphedStep(vp, dev_local, B, capture_all_reduce=car)
    piped = None if args.no_graph else PipelinedGraphedStep(vp, dev_local, B, capture_all_reduce=car)

    def step_resident():
        return graphed.run() if graphed is not None else step_eager()

    def step_e2e():
        if piped is not None:
            # H2D of this step's inputs (pinned) into the graph's static buffers; the NEXT step's inputs are prefetched on
            # a copy stream while this step computes (every step's copy is inside the timed region)
            losses = piped.run(host_local, host_local)
        else:
            inp = {k: v.to(dev, non_blocking=True) for k, v in host_local.items()}
            losses = vp.forward_backward(inp, None, already_sharded=True, n_total_views=B)
        return float(losses['rgb'].item())                                        # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    # ---- device-resident timing: per-step CUDA event pairs, L2 flushed between steps (outside the pairs)
    launches0 = _lib.launch_count()
    step_eager()
    launches_per_step = _lib.launch_count() - launches0      # our kernels per step (a graph replays exactly these)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_resident()
        b.record()
        evs.append((a, b))
    barrier()
    launches = launches_per_step * args.steps
    ms_local = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.summary() if sampler else None

    # ---- per-kernel durations (roofline): the same steps run eagerly with CUDA events around the raster kernels on their
    # launch stream (events cannot be read back from inside a replayed graph; kernel durations do not depend on how
    # the launch was submitted)
    _lib.lib().dbw_timing_reset()
    _lib.lib().dbw_timing_enable(1)
    for _ in range(args.steps):
        flush.zero_()
        step_eager()
    barrier()
    _lib.lib().dbw_timing_enable(0)
    kt = {(kind, kk): _lib.kernel_time_ms(kind, kk) for kind in (0, 1) for kk in (1, K)}
    _lib.lib().dbw_timing_reset()

    # ---- end-to-end timing through the public API with host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    evs2 = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_e2e()
        b.record()
        evs2.append((a, b))
    barrier()
    ms_e2e_local = sum(a.elapsed_time(b) for a, b in evs2)

    t = torch.tensor([ms_local, ms_e2e_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e_total = t.tolist()

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = B * args.steps / (ms_total / 1e3)
        e2e_value = B * args.steps / (ms_e2e_total / 1e3)
        peak, peak_src = peaks()
        # dominant kernel = the raster kernel with the largest summed duration over the timed region
        names = {(0, 1): 'raster_forward[env K=1]', (0, K): f'raster_forward[blocks K={K}]',
                 (1, 1): 'raster_backward[env K=1]', (1, K): f'raster_backward[blocks K={K}]'}
        dom = max(kt, key=lambda k: kt[k][0])
        dom_ms = kt[dom][0] / max(kt[dom][1], 1)
        kk = dom[1]
        # per launch: each direction moves half of H*W*(32+8K) per view (16 B RGBA + 4K B ids), B_local views per launch
        alg_bytes = B_local * H * W * (16 + 4 * kk)
        achieved = alg_bytes / (dom_ms / 1e3) / 1e9
        h2d = sum(v.numel() * v.element_size() for v in host_local.values())
        line = {
            'metric': 'render+backward views/sec (400x400, 10 blocks)', 'value': value, 'unit': 'views/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'DTU scan24 shape: 10 superquadric blocks (800 faces) + env (448 faces), 49 views 400x400, '
                                   'K=10, 256^2 textures, coarse phase (sigma=1e-4, per-face opacities); views sharded over ranks '
                                   '(7,6,6,..), one NCCL all-reduce of the flat gradient bucket',
                       'views_per_step': B, 'l2': 'flushed (256 MB memset) between steps, outside the per-step event pairs',
                       'loss': 'rgb (MSE) only; LPIPS excluded (SURVEY 8d)', 'seed': SEED,
                       'submission': 'eager' if graphed is None else 'whole step captured once in a CUDA graph and replayed',
                       'e2e_pipeline': 'none' if piped is None else 'inputs of step i+1 copied H2D on a side stream during step i (double-buffered)'},
            'e2e': {'value': e2e_value, 'unit': 'views/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': names[dom], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': NCU_TRAFFIC_BYTES.get(names[dom]) if world == 1 else None,
                         'traffic_source': 'profiles/ncu_full_r1f_summary.txt (bytes per launch at N=1)', 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': alg_bytes, 'avg_launch_ms': dom_ms,
                         'kernels_ms_per_step': {names[k]: kt[k][0] / args.steps for k in kt}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(n_views=16)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU legs (oracle port)
def oracle_step(n_views, threads=None):
    """One render+backward pass of the reference's algorithm (oracle port) over n_views views of the workload."""
    from oracle import dbw_path as D
    H, W, K = WORKLOAD['height'], WORKLOAD['width'], WORKLOAD['faces_per_pixel']
    tpl = D.SceneTemplate(n_blocks=WORKLOAD['n_blocks'], txt_size=WORKLOAD['txt_size'])
    p = {k: v.requires_grad_(True) for k, v in D.init_params(WORKLOAD['n_blocks'], WORKLOAD['txt_size'], seed=SEED).items()}
    R, T, Km = D.ring_cameras(WORKLOAD['n_views'])
    g = torch.Generator().manual_seed(SEED)
    imgs = torch.rand(n_views, 3, H, W, generator=g)
    keep = torch.sigmoid(p['alpha_logit'].detach()) > 0.01
    t0 = time.perf_counter()
    rec = D.predict(tpl, p, R[:n_views], T[:n_views], Km, (H, W), sigma=1e-4, faces_per_pixel=K, z_clip=0.001, keep=keep, decimate=8)
    loss = D.mse_loss(imgs, rec)
    loss.backward()
    return time.perf_counter() - t0


def cpu_threads():
    """threads for the CPU legs: all host cores up to 32 (beyond that the small torch ops of the path get slower, not
    faster; measured on the 128-core GPU box)"""
    return min(os.cpu_count(), 32)


def cpu_baseline(n_views=16):
    torch.set_num_threads(cpu_threads())
    oracle_step(1)                                   # warm-up (page in the library, thread pools)
    dt = oracle_step(n_views)
    return {'value': n_views / dt, 'unit': 'views/s', 'cores': cpu_threads(), 'kind': 'port',
            'sample': f'{n_views} of the 49 views (400x400, 10 blocks, K=10), forward+backward once, '
                      f'OpenMP rasterizer + torch ops on {cpu_threads()} of {os.cpu_count()} host threads; {dt:.1f} s'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    torch.set_num_threads(cpu_threads())
    n = 2
    for _ in range(min(args.warmup, 1)):
        oracle_step(1)
    tot = 0.0
    for _ in range(args.steps):
        tot += oracle_step(n)
    value = n * args.steps / tot
    sample = f'{n} of the 49 views per step (400x400, 10 blocks, K=10), forward+backward'
    print(json.dumps({
        'impl': 'reference', 'metric': 'render+backward views/sec (400x400, 10 blocks)', 'value': value, 'unit': 'views/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': tot / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'DTU scan24 shape (same as our arm); bounded sample: ' + sample, 'seed': SEED},
        'cpu_baseline': {'value': value, 'unit': 'views/s', 'cores': cpu_threads(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'views/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'PyTorch3D (the reference dependency that owns this arithmetic) is not installable here; this arm times the '
                'CPU restatement of its algorithm (oracle/, "port") on the host cores',
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='submit the step eagerly instead of replaying a CUDA graph')
    ap.add_argument('--allreduce-in-graph', action='store_true', help='capture the NCCL all-reduce into the CUDA graph (experimental)')
    args = ap.parse_args()
    if args.impl == 'reference':
        if args.steps > 5:
            args.steps = 5          # each step is a bounded CPU sample; keep the whole run within minutes
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
