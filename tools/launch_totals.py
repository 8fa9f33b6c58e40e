import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split('(')[0]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e3
for k,(n,t) in agg.items(): print("%-28s n=%3d total %9.1f us  avg %8.1f us"%(k,n,t,t/n))
