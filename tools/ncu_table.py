"""Last value of every metric per kernel from an `ncu --csv` log.  usage: python tools/ncu_table.py file.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
agg = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    agg.setdefault(d["Kernel Name"].split("(")[0].split("::")[-1][:28], collections.OrderedDict())[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
for k, v in agg.items():
    print(k, " ".join("%s=%.4g" % (m.split("__")[-1][:34], x) for m, x in v.items()))
