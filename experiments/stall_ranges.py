"""Warp-state samples of one kernel of an `ncu --page source --csv --print-source sass` dump, summed over ranges of N SASS instructions
(usage: stall_ranges.py dump.csv <kernel substring> [N]).  Used to see which warp role of a warp-specialised kernel is busy and which waits."""
import csv, sys
path=sys.argv[1]; want=sys.argv[2]; step=int(sys.argv[3]) if len(sys.argv)>3 else 100
rows=list(csv.reader(open(path))); secs=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={"name":r[1],"hdr":None,"data":[]}; secs.append(cur)
    elif cur is not None and r and r[0]=="Address": cur["hdr"]=r
    elif cur is not None and cur["hdr"] and len(r)==len(cur["hdr"]): cur["data"].append(r)
s=[x for x in secs if want in x["name"]][0]
hdr,data=s["hdr"],s["data"]; ix={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[ix["# Samples"]]) for r in data)
print(s["name"][:80], "samples", tot)
stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for b in range(0,len(data),step):
    blk=data[b:b+step]
    n=sum(int(r[ix["# Samples"]]) for r in blk)
    ex=sum(int(r[ix["Instructions Executed"]]) for r in blk)
    st={}
    for r in blk:
        for h in stalls:
            v=int(r[ix[h]]); 
            if v: st[h[6:]]=st.get(h[6:],0)+v
    top=sorted(st.items(), key=lambda kv:-kv[1])[:4]
    mx=max(blk,key=lambda r:int(r[ix["Instructions Executed"]]))
    print("%5d-%5d samples %6d %5.1f%% exec %9d  %s" % (b,b+len(blk)-1,n,100.0*n/tot,ex,top))
