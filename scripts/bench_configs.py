"""Timing of the other BASELINE.json shapes on one GPU (not bench lines: sanity that nothing pathological happens
when faces, K or resolution grow).  python scripts/bench_configs.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from copy import deepcopy
import bench
import dbw_b200
from dbw_b200 import _lib
from dbw_b200.dbw import DifferentiableBlocksWorld
from dbw_b200.synthetic import ring_cameras

dev = torch.device('cuda:0')
CASES = {
    'cfg2 dtu 400x400 N=10 K=10 B=49': dict(H=400, W=400, N=10, K=10, txt=256, up=1, B=49),
    'cfg2 FINE phase (sigma=5e-6, hard opacities) B=49': dict(H=400, W=400, N=10, K=10, txt=256, up=1, B=49, fine=True),
    'cfg4 bmvs 576x768 N=10 K=10 B=8': dict(H=576, W=768, N=10, K=10, txt=256, up=1, B=8),
    'cfg5 stress 800x800 N=50 K=25 B=4': dict(H=800, W=800, N=50, K=25, txt=128, up=2, B=4),
}
for name, c in CASES.items():
    cfg = deepcopy(bench.MODEL_CFG)
    cfg['mesh'].update(n_blocks=c['N'], txt_size=c['txt'], txt_bkg_upscale=c['up'])
    cfg['renderer']['faces_per_pixel'] = c['K']
    torch.manual_seed(0)
    model = DifferentiableBlocksWorld((c['H'], c['W']), **cfg).to(dev); model.train()
    if c.get('fine'):
        model.set_cur_epoch(1600)
        with torch.no_grad():
            model.alpha_logit.copy_(torch.linspace(-1, 3, c['N']))
    R, T, K = ring_cameras(c['B'])
    inp = {'imgs': torch.rand(c['B'], 3, c['H'], c['W'], device=dev), 'R': R.to(dev), 'T': T.to(dev), 'K': K[None].expand(c['B'], -1, -1).to(dev)}
    for _ in range(3):
        model.zero_grad(); model(inp, None)['total'].backward()
    torch.cuda.synchronize()
    _lib.lib().dbw_timing_reset(); _lib.lib().dbw_timing_enable(1)
    t0 = time.perf_counter(); n = 5
    for _ in range(n):
        model.zero_grad(); model(inp, None)['total'].backward()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    _lib.lib().dbw_timing_enable(0)
    kt = {f'{"fwd" if kind == 0 else "bwd"}[K={kk}]': round(_lib.kernel_time_ms(kind, kk)[0] / n, 3) for kind in (0, 1) for kk in (1, c['K'])}
    _lib.lib().dbw_timing_reset()
    print(f'{name}: {dt * 1e3:.2f} ms/step eager -> {c["B"] / dt:.0f} views/s; raster kernels ms/step {kt}; us/view {sum(kt.values()) / c["B"] * 1e3:.0f}')
