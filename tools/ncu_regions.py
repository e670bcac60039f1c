"""Samples / warp instructions / SIMT efficiency of one kernel by source region (file, first line, last line, name).

  python tools/ncu_regions.py report.ncu-rep k_orca tools/ncu_regions_k_orca.txt
"""
import csv, sys, subprocess, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file="?"; agg=[]; hdr=None
for r in rows:
    if len(r)>=2 and r[0].strip()=="File Path": cur_file=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)>10 and r[0] not in ("",):
        try: ln=int(r[0])
        except: continue
        d=dict(zip(hdr[4:],r[4:]))
        if not d["# Samples"].isdigit(): continue
        agg.append((cur_file,ln,int(d["# Samples"]),int(d["Instructions Executed"]),int(d["Thread Instructions Executed"])))
tot_s=sum(a[2] for a in agg); tot_i=sum(a[3] for a in agg)
regions = eval(open(sys.argv[3]).read())
res=collections.OrderedDict()
for f,lo,hi,name in regions: res[name]=[0,0,0]
res["other"]=[0,0,0]
for f,ln,s,i,t in agg:
    for rf,lo,hi,name in regions:
        if f==rf and lo<=ln<=hi:
            res[name][0]+=s; res[name][1]+=i; res[name][2]+=t; break
    else:
        res["other"][0]+=s; res["other"][1]+=i; res["other"][2]+=t
for k,(s,i,t) in res.items():
    print(f"{k:28s} samples {s/tot_s:6.1%}  warp-instr {i/tot_i:6.1%}  simt {t/max(1,i)/32:4.2f}")
