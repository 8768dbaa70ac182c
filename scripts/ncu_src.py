"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda --kernel-id :::N` :
instructions executed and stall samples per CUDA source line (file:line), plus totals per stall reason."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = open(path).read().splitlines()
inst = collections.Counter(); smp = collections.Counter(); reasons = collections.Counter()
cur = '?'; hdr = None
for l in lines:
    if l.startswith('"File Path"'):
        cur = l.split(',', 1)[1].strip('"').split('/')[-1]; hdr = None; continue
    if l.startswith('"Function Name"'):
        continue
    if l.startswith('"Line No"'):
        hdr = next(csv.reader([l]))
        ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
        isrc = hdr.index('Source'); stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None:
        continue
    r = next(csv.reader([l]))
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ie]); s = int(r[isamp])
    except ValueError:
        continue
    key = (cur, r[0], r[isrc].strip()[:105])
    inst[key] += n; smp[key] += s
    for i in stall:
        try: reasons[hdr[i]] += int(r[i])
        except ValueError: pass
ti, ts = sum(inst.values()), sum(smp.values())
print(f'total warp-instructions {ti:,}  samples {ts:,}')
print('stall reasons:', ', '.join(f'{k[6:]}={100*v/max(ts,1):.1f}%' for k, v in reasons.most_common(9)))
print('--- by instructions executed')
for k, n in inst.most_common(topn):
    print(f'{100*n/ti:5.1f}% inst {100*smp[k]/max(ts,1):5.1f}% smp  {k[0][:13]:13s}:{k[1]:>4s} {k[2]}')
print('--- by stall samples')
for k, n in smp.most_common(topn // 2):
    print(f'{100*inst[k]/ti:5.1f}% inst {100*n/max(ts,1):5.1f}% smp  {k[0][:13]:13s}:{k[1]:>4s} {k[2]}')
