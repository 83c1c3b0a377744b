#!/usr/bin/env python
"""Dev helper: one line per launch of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log: duration, DRAM bytes, DRAM bandwidth."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
ci = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[ci['ID']], r[ci['Kernel Name']][:44], r[ci['Grid Size']]), {})[r[ci['Metric Name']]] = float(r[ci['Metric Value']].replace(',', ''))
tot = sum(m.get('gpu__time_duration.sum', 0) for m in agg.values()) / 1e6
for (i, k, g), m in agg.items():
    t = m.get('gpu__time_duration.sum', 0) / 1e6
    b = m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0)
    print(f"{i:>3} {k:44s} grid {g:>16s} {t:8.3f} ms {100 * t / tot:5.1f}%  DRAM {b / 1e9:7.3f} GB  {b / 1e9 / (t / 1e3 + 1e-12) / 1e3:5.2f} TB/s")
print(f"total {tot:.3f} ms")
