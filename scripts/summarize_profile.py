"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep} into the committed summaries under profiles/."""
import collections, csv, os, subprocess, sys
tag = sys.argv[1]
steps_in_run = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # 0: infer from the once-per-step scene geometry kernel
os.makedirs('profiles', exist_ok=True)
lines = [l for l in open(f'gpurun_out/launches_{tag}.csv') if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[u]
    a = agg.setdefault(row['Kernel Name'][:90], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
if not steps_in_run:
    steps_in_run = max(1, sum(a[0] for k, a in agg.items() if k.startswith('scene_geometry_forward_kernel')))
with open(f'profiles/launches_{tag}_summary.txt', 'w') as f:
    f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n')
    f.write(f'# {sum(a[0] for a in agg.values())} launches over {steps_in_run} steps; per-step averages; cold-cache serialised times: compare SHARES\n')
    f.write(f'# total {tot / steps_in_run:.3f} ms/step, {sum(a[0] for a in agg.values()) / steps_in_run:.0f} launches/step\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write(f'{a[1] / steps_in_run:9.3f} ms/step {a[0] / steps_in_run:7.1f} launches/step {100 * a[1] / tot:5.1f}%  {k}\n')
raw = subprocess.run(['ncu', '-i', f'gpurun_out/prof_{tag}.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
idx = {h: i for i, h in enumerate(hdr)}
with open(f'profiles/ncu_full_{tag}_summary.txt', 'w') as f:
    f.write(f'# ncu --set full --clock-control none --import-source on -k regex:raster_ -s 16 -c 4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n')
    f.write('# one step at cfg 2 (49 views 400x400, 10 blocks): env forward (K=1), blocks forward (K=10, with the loss epilogue), blocks backward, env backward\n')
    for w in want:
        if w in idx:
            f.write(f'{w:70s} [{units[idx[w]]:>14s}] ' + ' | '.join(r[idx[w]][:24] for r in data) + '\n')
# bytes per launch of the four raster kernels for bench.py's roofline.traffic, tied to the kernel sources they were measured on
import json
sys.path.insert(0, os.getcwd())
import bench
names = ['raster_forward[env K=1]', 'raster_forward[blocks K=10]', 'raster_backward[blocks K=10]', 'raster_backward[env K=1]']
def col(name):
    i = idx[name]; u = units[i]
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    return [float(r[i].replace(',', '')) * mult for r in data]
rd, wr = col('dram__bytes_read.sum'), col('dram__bytes_write.sum')
if len(rd) == 4 and '--no-traffic-json' not in sys.argv:
    json.dump({'kernel_fingerprint': bench.kernel_fingerprint(), 'workload': 'dtu', 'source': f'profiles/ncu_full_{tag}_summary.txt',
               'bytes_per_launch': {n: rd[i] + wr[i] for i, n in enumerate(names)}}, open('profiles/ncu_traffic.json', 'w'), indent=1)
    print('wrote profiles/ncu_traffic.json', bench.kernel_fingerprint())
print(open(f'profiles/launches_{tag}_summary.txt').read()[:1800])
print(open(f'profiles/ncu_full_{tag}_summary.txt').read())
