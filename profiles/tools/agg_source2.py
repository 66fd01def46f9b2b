"""Aggregate `ncu -i REP --page source --print-source cuda,sass --csv --kernel-name regex:K` per CUDA source line.
Usage: python agg_source2.py FILE.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur = None; hdr = None
agg = collections.OrderedDict(); src = {}; stalls = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) <= 2: continue
    if r[0] == "Line No":
        hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
        st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if not r[0]: continue
    try: samples = int(r[iS] or 0); inst = int(r[iI] or 0); tinst = int(r[iT] or 0)
    except ValueError: continue
    key = (cur, int(r[0]))
    if key not in agg: agg[key] = [0, 0, 0]; src[key] = r[1]
    a = agg[key]; a[0] += samples; a[1] += inst; a[2] += tinst
    for i, h in st:
        try: stalls[h] += int(r[i] or 0)
        except ValueError: pass
tot = [sum(a[i] for a in agg.values()) for i in range(3)]
print("total samples %d warp-inst %d thread-inst %d (avg lanes %.1f)" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
s = sum(stalls.values())
print("## stall reasons"); 
for k, v in stalls.most_common(10): print("  %-26s %6.2f%%" % (k, 100 * v / s))
print("## by file")
byf = collections.Counter(); byfs = collections.Counter()
for (f, l), a in agg.items(): byf[f] += a[1]; byfs[f] += a[0]
for f, v in byf.most_common(): print("  %-22s inst %5.2f%% samples %5.2f%%" % (f, 100 * v / tot[1], 100 * byfs[f] / tot[0]))
print("## top lines by warp instructions")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print("%-20s %4d  samp %5.2f%%  inst %5.2f%% lanes %4.1f | %s" % (f, l, 100*a[0]/tot[0], 100*a[1]/tot[1], a[2]/max(a[1],1), src[(f, l)].strip()[:110]))
