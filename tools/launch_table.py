"""Print an ncu --csv launch list (several metrics per launch) as a table (development aid).  python tools/launch_table.py file.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii], {'k': r[ki]})[r[mi]] = r[vi]
f = lambda v, m: float(v.get(m, '0').replace(',', ''))
tot = 0.0
for i, v in d.items():
    k = v['k'].replace('<unnamed>::', '').replace('void ', '')[:64]
    if k.startswith('at::'): continue
    t = f(v, 'gpu__time_duration.sum') / 1e3; tot += t
    print('%-64s %9.1f us  rd %8.1f MB wr %8.1f MB dmma %5.1f%%' % (k, t, f(v, 'dram__bytes_read.sum') / 1e6, f(v, 'dram__bytes_write.sum') / 1e6,
          f(v, 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active')))
print('total %.1f us' % tot)
