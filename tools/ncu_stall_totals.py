import csv, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cols=None; tot=collections.Counter(); byfile=collections.defaultdict(collections.Counter); fname=''
for r in rows:
    if not r: continue
    if r[0] in ("File Path","File Name"): fname=r[1].split('/')[-1]; continue
    if r[0]=="Line No": cols=r; continue
    if cols is None or len(r)<len(cols) or r[0]: continue   # SASS rows only: source text with quotes or commas shifts the columns of a source row
    for k,c in enumerate(cols):
        if c.startswith("stall_") and "Not Issued" not in c and r[k] not in ("","0"):
            try: v=int(float(r[k]))
            except: continue
            tot[c[6:]]+=v; byfile[fname][c[6:]]+=v
s=sum(tot.values())
print("total",s)
for k,v in tot.most_common(): print(f"  {k:20s} {100*v/s:5.1f}%")
for f,c in byfile.items():
    ss=sum(c.values()); print(f, f"{100*ss/s:.1f}%", ", ".join(f"{k} {100*v/ss:.0f}%" for k,v in c.most_common(5)))
