"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,...` launch list: python scripts/launch_summary.py file.csv [max_rows]"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi, mi, ii = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name'), H.index('ID')
per = collections.OrderedDict()
for r in data:
    per.setdefault(r[ii], {'name': r[ki][:78]})[r[mi]] = float(r[vi].replace(',', ''))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10**9
for i, d in list(per.items())[:n]:
    if d.get('gpu__time_duration.sum', 0) < 20000:
        continue
    print(f"{i:>3} {d['gpu__time_duration.sum']/1e6:8.3f} ms  rd {d.get('dram__bytes_read.sum',0)/1e9:7.3f} GB wr {d.get('dram__bytes_write.sum',0)/1e9:7.3f} GB"
          f"  L2hit {d.get('lts__t_sector_hit_rate.pct',0):5.1f}% warps {d.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f}% {d['name']}")
