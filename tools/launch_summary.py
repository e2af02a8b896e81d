"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (name, template mode, grid)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
MODES = {0: "MSG0", 1: "MSG", 2: "GATE", 3: "EU1", 4: "EU2", 5: "LIN", 6: "MSGA"}
for row in rows:
    k = row['Kernel Name']
    name = k.split('<')[0].replace('void ', '').split('(')[0].replace('fm::', '')
    m = re.search(r'>, \(?(?:fm::EgMode|int)?\)?(\d+)(?:, (?:\(int\))?(\d+))?(?:, (?:\(int\))?(\d+))?>\(', k)
    if "egemm_c" in name and m:
        name += "<EU>"
    elif "egemm" in name and m:
        name += f"<{MODES.get(int(m.group(1)), m.group(1))}" + (f",img{m.group(2)}" if 'egemm_p' in name and m.group(2) not in (None, '0') else "") + ">"
    name += " g=" + row.get('Grid Size', '?').strip('()').split(',')[0]
    v = float(row['Metric Value'])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.3f} ms over {len(rows)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} n={n:4d} total_ms={t/1e6:9.3f} share={100*t/tot:5.1f}% avg_us={t/n/1e3:9.1f}")
