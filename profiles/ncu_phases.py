"""Phase profile of one kernel from an `ncu --page source --csv` export: the SASS listing is cut at barriers
(BAR.SYNC / __syncwarp's WARPSYNC are kept inside) and each segment gets its share of the stall samples, of the
shared-memory wavefronts and of the executed warp instructions.   python profiles/ncu_phases.py file.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
h = rows[1]
si, src, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
wf = h.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in h else ie
wfx = h.index('L1 Wavefronts Shared Excessive') if 'L1 Wavefronts Shared Excessive' in h else ie
stalls = [k for k in h if k.startswith('stall_') and '(Not Issued)' not in k]
body = []
for r in rows[2:]:                      # the first launch only (every launch repeats the two header rows)
    if r and r[0] == 'Kernel Name': break
    if len(r) == len(h): body.append(r)
def I(r, i):
    try: return int(r[i] or 0)
    except ValueError: return 0
tot, twf, tex = sum(I(r, si) for r in body), max(1, sum(I(r, wf) for r in body)), sum(I(r, ie) for r in body)
print(rows[0][1][:90])
print("samples %d  warp-instr %.1fM  smem wavefronts %.1fM (excess %.1f%%)" % (tot, tex / 1e6, twf / 1e6, 100 * sum(I(r, wfx) for r in body) / twf))
agg = {k: sum(I(r, h.index(k)) for r in body) for k in stalls}
print("stalls: " + "  ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
seg, a = [], 0
for i, r in enumerate(body):
    if 'BAR.SYNC' in r[src] or i == len(body) - 1:
        seg.append((a, i)); a = i + 1
for a, b in seg:
    rs = body[a:b + 1]
    s, w, n = sum(I(r, si) for r in rs), sum(I(r, wf) for r in rs), sum(I(r, ie) for r in rs)
    if 100 * s / tot < minpct and 100 * w / twf < minpct: continue
    ops = {}
    for r in rs:
        t = r[src].split()
        op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')).split('.')[0]
        ops[op] = ops.get(op, 0) + I(r, si)
    top = " ".join("%s:%.1f" % (k, 100 * v / tot) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:5])
    print("%5d-%5d len %4d  time %5.1f%%  wf %5.1f%%  instr %5.1f%%   %s" % (a, b, b - a + 1, 100 * s / tot, 100 * w / twf, 100 * n / tex, top))
