"""dev: aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: ncu_lines.py src.csv <warp-plies (divisor)> [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
per = collections.Counter(); smp = collections.Counter(); src = {}
f = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); ns = r.index("# Samples"); continue
    if r[0] in ("Function Name",) or hdr is None or r[0] == "": continue
    try: n = int(r[ie])
    except ValueError: continue
    k = (f, int(r[0])); per[k] += n; src[k] = r[1].strip()[:120]
    try: smp[k] += int(r[ns])
    except ValueError: pass
tot = sum(per.values()); ts = max(1, sum(smp.values()))
print("total warp instructions %d  = %.1f per unit" % (tot, tot / div))
for k, n in per.most_common(top):
    print("%7.1f %5.1f%% stall-samples %5.1f%%  %s:%d  %s" % (n / div, 100 * n / tot, 100 * smp[k] / ts, k[0], k[1], src[k]))
