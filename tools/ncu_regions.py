#!/usr/bin/env python
"""Per-region stall profile of one kernel from `ncu --page source --csv` (tools/ncu_sum.sh writes /tmp/ncu_<rep>.src.csv):
windows of W SASS instructions with their share of the warp-stall samples, samples per executed warp instruction, the top
stall reasons and the top opcodes.  usage: tools/ncu_regions.py SRC.csv [W]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def f(x):
    try: return float(x)
    except ValueError: return 0.0
tot = sum(f(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
print('first-instr  exec(M) samples share  samples/kinstr  top stalls | top opcodes')
for a in range(0, len(data), W):
    blk = data[a:a + W]
    ie = sum(f(r[ix['Instructions Executed']]) for r in blk)
    s = sum(f(r[ix['# Samples']]) for r in blk)
    if not s: continue
    d = {h: sum(f(r[ix[h]]) for r in blk) for h in st}
    top = sorted(d.items(), key=lambda kv: -kv[1])[:5]
    ops = {}
    for r in blk:
        w = r[ix['Source']].split()
        if not w: continue
        op = (w[1] if w[0].startswith('@') else w[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    topo = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print(f"{a:5d} {ie/1e6:8.2f} {s:7.0f} {100*s/tot:5.1f}% {s/max(ie,1)*1e3:7.3f}  " + ' '.join(f"{k[6:]}={v/s*100:.0f}%" for k, v in top) + '  | ' + ' '.join(f"{k}:{v}" for k, v in topo))
