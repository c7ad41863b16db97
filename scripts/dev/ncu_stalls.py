"""dev: rank CUDA source lines of an `ncu --page source --csv --print-source cuda,sass` export by warp-stall samples."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None; f = None
smp = collections.Counter(); ins = collections.Counter(); src = {}
reasons = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ns = r.index("# Samples"); ie = r.index("Instructions Executed"); continue
    if hdr is None or r[0] in ("Function Name", ""): continue
    try: n = int(r[ns])
    except ValueError: continue
    k = (f, int(r[0])); smp[k] += n; src[k] = r[1].strip()[:90]
    try: ins[k] += int(r[ie])
    except ValueError: pass
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: reasons[k][h] += int(r[j])
            except ValueError: pass
tot = sum(smp.values())
print("total samples", tot, "instructions", sum(ins.values()))
for k, n in smp.most_common(top):
    rs = ", ".join("%s %d" % (a.replace("stall_", ""), b) for a, b in reasons[k].most_common(3))
    print("%5.1f%% inst %8d %s:%d %s | %s" % (100 * n / tot, ins[k], k[0], k[1], src[k], rs))
