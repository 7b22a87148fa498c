"""Per-source-line instruction / stall-sample totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
seen, end = set(), len(rows)
for i, r in enumerate(rows):
    if r and r[0] == 'File Path':
        if r[1] in seen:
            end = i
            break
        seen.add(r[1])
def num(x):
    try: return int(x)
    except ValueError: return 0
cur, agg, tot, tots = None, {}, 0, 0
for r in rows[:end]:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 7 and r[0].isdigit():
        key = (cur, int(r[0]), r[1][:105])
        a = agg.setdefault(key, [0, 0])
        a[0] += num(r[7]); a[1] += num(r[6]); tot += num(r[7]); tots += num(r[6])
print('total warp-instructions', tot, 'samples', tots)
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{str(k[0])[-10:]:10s} {k[1]:4d} {n:9d} {100*n/tot:5.1f}% samp {s:4d}  {k[2]}")
