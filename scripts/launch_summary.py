"""Summarise an ncu launch list CSV (gpu__time_duration.sum) by kernel: python scripts/launch_summary.py in.csv out.txt 'header...'"""
import csv, collections, re, sys
src, dst = sys.argv[1], sys.argv[2]
hdr = sys.argv[3:] 
lines=[l for l in open(src) if not l.startswith('==')]
tot=collections.defaultdict(float); cnt=collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    v = v/1e6 if unit=='ns' else v/1e3 if unit=='us' else v*1e3 if unit=='s' else v
    name=re.sub(r'\(.*','',row['Kernel Name'])[:90]
    tot[name]+=v; cnt[name]+=1
T=sum(tot.values()); n=sum(cnt.values())
out=["# "+h for h in hdr]+[f"# {n} launches, {T:.2f} ms total (cold-cache, serialised timings: compare shares, not absolutes)"]
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:40]:
    out.append(f"{v:9.3f} ms {100*v/T:5.1f}%  n={cnt[k]:5d}  avg={1000*v/cnt[k]:9.1f} us  {k}")
open(dst,'w').write("\n".join(out)+"\n")
print("\n".join(out[:34]))
