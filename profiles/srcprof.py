import csv, sys, subprocess, collections
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file="?"; agg=[]
hdr=None
for r in rows:
    if len(r)>=2 and r[0].strip()=="File Path": cur_file=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)>10 and r[0] not in ("",):
        try: ln=int(r[0])
        except: continue
        d=dict(zip(hdr[4:],r[4:]))
        if not d["# Samples"].isdigit(): continue
        agg.append((cur_file,ln,r[1].strip()[:90],int(d["# Samples"]),int(d["Instructions Executed"]),int(d["Thread Instructions Executed"])))
tot_s=sum(a[3] for a in agg); tot_i=sum(a[4] for a in agg); tot_t=sum(a[5] for a in agg)
print(f"total samples {tot_s} warp-instr {tot_i} thread-instr {tot_t} simt-eff {tot_t/max(1,tot_i)/32:.2f}")
byfile=collections.Counter()
for a in agg: byfile[a[0]]+=a[3]
print("samples by file:", dict(byfile))
for a in sorted(agg,key=lambda a:-a[3])[:topn]:
    print(f"{a[0]:12s}:{a[1]:4d} smp {a[3]/tot_s:5.1%} inst {a[4]/tot_i:5.1%} thr/inst {a[5]/max(1,a[4]):4.1f} | {a[2]}")
