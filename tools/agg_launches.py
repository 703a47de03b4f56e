import csv,sys,collections,re
rows=[l for l in open(sys.argv[1]) if not l.startswith("==")]
r=list(csv.DictReader(rows))
agg=collections.OrderedDict()
tot=0
for x in r:
    n=re.sub(r"\(.*","",x["Kernel Name"]).replace("wd::","").replace("void ","").replace("<unnamed>::","")
    t=float(x["Metric Value"].replace(",",""))
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=t; tot+=t
print(f"total {tot/1e6:.2f} ms over {len(r)} launches")
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print(f"{n:60s} {c:5d} {t/1e3:10.1f} us {t/tot:6.3f}")
