"""Per-region summary of an `ncu --page source --csv` dump (several kernels): instructions, stall samples, top stall reasons."""
import csv, sys
fn, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(fn)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
for ki in range(len(starts) - 1):
    blk = rows[starts[ki]:starts[ki + 1]]
    kname = blk[0][1]
    if want and want not in kname: continue
    hdr = blk[1]; idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in blk[2:] if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[idx['# Samples']] or 0) for r in data)
    toti = sum(int(r[idx['Instructions Executed']] or 0) for r in data)
    print('=== kernel', ki, kname[:150], 'samples', tot, 'instr %.1fM' % (toti / 1e6))
    for b in range(0, len(data), B):
        chunk = data[b:b + B]
        ie = sum(int(r[idx['Instructions Executed']] or 0) for r in chunk)
        s = sum(int(r[idx['# Samples']] or 0) for r in chunk)
        if s < tot * 0.004 and ie < toti * 0.004: continue
        st = {}
        for r in chunk:
            for h in stalls: st[h] = st.get(h, 0) + int(r[idx[h]] or 0)
        top = [(k.replace('stall_', ''), v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3]]
        ops = {}
        for r in chunk:
            t = r[idx['Source']].strip().split()
            op = t[1] if t[0].startswith('@') else t[0]
            ops[op] = ops.get(op, 0) + int(r[idx['Instructions Executed']] or 0)
        topops = [k for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4]]
        print(f"{b:5d} instr {100*ie/toti:5.1f}% samples {100*s/tot:5.1f}% {top} {topops}")
