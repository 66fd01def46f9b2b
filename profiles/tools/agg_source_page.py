import csv, sys, collections
fn = sys.argv[1]
rows = list(csv.reader(open(fn)))
cur = None; hdr = None
agg = collections.OrderedDict(); src = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) <= 2: continue
    if r[0] == "Line No": hdr = r; continue
    if not r[0]: continue
    try: samples = int(r[6] or 0); inst = int(r[7] or 0); tinst = int(r[8] or 0)
    except ValueError: continue
    key = (cur, int(r[0]))
    if key not in agg: agg[key] = [0, 0, 0]; src[key] = r[1]
    a = agg[key]; a[0] += samples; a[1] += inst; a[2] += tinst
tot = [sum(a[i] for a in agg.values()) for i in range(3)]
print("total samples %d inst %d thread-inst %d" % tuple(tot))
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
top = sorted(agg.items(), key=lambda kv: -kv[1][k])[:int(sys.argv[3]) if len(sys.argv) > 3 else 50]
for (f, l), a in top:
    print("%-20s %4d  samp %5.2f%%  inst %5.2f%% thr/inst %4.1f | %s" % (f, l, 100*a[0]/tot[0], 100*a[1]/tot[1], a[2]/max(a[1],1), src[(f, l)][:100]))
