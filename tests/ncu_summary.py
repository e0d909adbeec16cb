"""One line per profiled kernel of an `ncu --set full` report:
    ncu -i gpurun_out/r2p_top.ncu-rep --page raw --csv > /tmp/raw.csv; python tests/ncu_summary.py /tmp/raw.csv > profiles/<name>.txt
Not a test."""
import csv
import re
import sys

COLS = [('dur_us', 'gpu__time_duration.sum', 1e-3), ('sm_MHz', 'smsp__cycles_elapsed.avg.per_second', 1e-6),
        ('tensor_pipe_active%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 1),
        ('tensor_pipe_elapsed%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 1),
        ('sm_throughput%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('dram_throughput%', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('dram_read_MB', 'dram__bytes_read.sum', None), ('dram_write_MB', 'dram__bytes_write.sum', None),
        ('l2_to_sm_read_MB', 'l1tex__m_xbar2l1tex_read_bytes.sum', None),
        ('grid', 'launch__grid_size', 1), ('block', 'launch__block_size', 1), ('cluster', 'launch__cluster_size', 1),
        ('regs', 'launch__registers_per_thread', 1), ('dyn_smem_KB', 'launch__shared_mem_per_block_dynamic', None)]
UNIT = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'Tbyte': 1e6}

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
head, units, data = rows[h], rows[h + 1], rows[h + 2:]
ix = {n: i for i, n in enumerate(head)}
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(r'pack_filter|elementwise|philox|fill|cast')
print('ncu --set full --clock-control none (B200, sm_100a); one launch per kernel, shapes as in tests/profile_kernels.py.')
print('tensor_pipe_active% = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active (the counter VERDICT r1 asked for);')
print('l2_to_sm_read_MB = l1tex__m_xbar2l1tex_read_bytes.sum; times are ncu-serialised (cold caches), not bench numbers.\n')
for r in data:
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '')
    if skip.search(name):
        continue
    out = []
    for label, metric, scale in COLS:
        if metric not in ix or r[ix[metric]] in ('', 'n/a'):
            continue
        v = float(r[ix[metric]].replace(',', ''))
        u = units[ix[metric]]
        if scale is None:
            v *= UNIT.get(u.split('/')[0], 1.0) * (1e3 if label.endswith('KB') else 1.0)
        elif label == 'dur_us':
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(u, 1e-3)
        elif label == 'sm_MHz':
            v *= {'Ghz': 1e3, 'Mhz': 1.0, 'hz': 1e-6, 'cycle/nsecond': 1e3, 'cycle/usecond': 1.0, 'cycle/second': 1e-6}.get(u, 1.0)
        out.append('%s=%s' % (label, ('%.2f' % v) if v != int(v) else '%d' % v))
    print('%-44s %s' % (name[:44], '  '.join(out)))
