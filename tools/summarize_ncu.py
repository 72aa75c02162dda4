"""Turn the ncu CSVs of tools/gpu_profile.sh into the small summaries committed under profiles/.
   python tools/summarize_ncu.py TAG ROUND      (reads gpurun_out/launches_TAG.csv, prof_TAG_raw.csv, bench_TAG.json)"""
import collections, csv, json, re, sys

tag, rnd = sys.argv[1], sys.argv[2]
G = 'gpurun_out/'

# ---- launch list of the last profiled eager step: duration + DRAM bytes per launch
rows = []
with open(G + f'launches_{tag}.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    k = int(r['ID'])
    e = per.setdefault(k, dict(kernel=re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('pg::', ''),
                               grid=r['Grid Size']))
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    name = r['Metric Name']
    if name == 'gpu__time_duration.sum':
        e['us'] = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    else:
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
        e[name.split('.')[0]] = v * scale
launches = list(per.values())
# the last step = the last quarter of the launches of this library (4 identical steps were profiled)
ours = [l for l in launches if not l['kernel'].startswith('at::')]
n = len(ours) // 4
step = ours[-n:]
tot = sum(l['us'] for l in step)
agg = collections.OrderedDict()
for l in step:
    a = agg.setdefault(l['kernel'], dict(launches=0, us=0.0, dram_bytes=0.0))
    a['launches'] += 1
    a['us'] += l['us']
    a['dram_bytes'] += l.get('dram__bytes_read', 0) + l.get('dram__bytes_write', 0)
summary = dict(what='one eager G+D step (cfg3, graphs and side streams off), ncu --metrics gpu__time_duration.sum,'
                    'dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; per-launch times are cold-cache '
                    'and serialised: compare shares',
               launches_in_step=n, sum_us=round(tot, 1),
               kernels=[dict(kernel=k, launches=v['launches'], us=round(v['us'], 1), share=round(v['us'] / tot, 4),
                             dram_MB_per_launch=round(v['dram_bytes'] / v['launches'] / 1e6, 2),
                             dram_GBps=round(v['dram_bytes'] / v['us'] / 1e3, 1))
                        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['us'])],
               step=[dict(kernel=l['kernel'], grid=l['grid'], us=round(l['us'], 2),
                          dram_MB=round((l.get('dram__bytes_read', 0) + l.get('dram__bytes_write', 0)) / 1e6, 2)) for l in step])
json.dump(summary, open(f'profiles/r{rnd}_ncu_step_launches.json', 'w'), indent=0)
print('launch list:', n, 'launches', round(tot, 1), 'us')
for k in summary['kernels'][:12]:
    print(k)

# ---- full capture of the dominant kernel
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'lts__t_bytes.sum', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor']
try:
    rows = list(csv.reader(open(G + f'prof_{tag}_raw.csv')))
    hdr, units = rows[0], rows[1]
    out = dict(what='ncu --set full --clock-control none of the conv_tc kernel (persistent variant), discriminator conv 256->512 stride 1, B32, 32x32 '
                    '(tools/conv_probe.py conv 1 32 32 256 512), 3 launches', launches=[])
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in want or h in ('Kernel Name', 'Grid Size', 'Block Size'):
                d[h + (f' [{units[i]}]' if units[i] else '')] = r[i]
        out['launches'].append(d)
    json.dump(out, open(f'profiles/r{rnd}_ncu_full_conv_tc_d3.json', 'w'), indent=0)
    for k, v in out['launches'][-1].items():
        print(k, v)
except (FileNotFoundError, IndexError):
    print('no full capture')
