import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr=None; tot=collections.Counter()
for r in rows:
    if len(r)<=2: continue
    if r[0]=="Line No": hdr=r; continue
    if not r[0]: continue
    for i,h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: tot[h]+=int(r[i] or 0)
            except: pass
s=sum(tot.values())
for k,v in tot.most_common(): print("%-28s %6.2f%%"%(k,100*v/s))
