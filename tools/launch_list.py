import csv, collections, sys
rows=[l for l in open(sys.argv[1]) if not l.startswith("==")]
agg=collections.OrderedDict()
for r in csv.DictReader(rows):
    if r["Metric Name"]!="gpu__time_duration.sum": continue
    k=r["Kernel Name"].split("(")[0][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r["Metric Value"].replace(",",""))
for k,a in agg.items(): print("%-62s %3d %10.1f us"%(k,a[0],a[1]/1e3))
