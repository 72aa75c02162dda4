"""Turn the ncu CSVs of tools/gpu_profile.sh into the small summaries committed under profiles/.
   python tools/summarize_ncu.py TAG ROUND      (reads gpurun_out/launches_TAG.csv, full_TAG_raw.csv, prof_TAG_raw.csv)"""
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


# ---- ncu --set full of the step's main kernels (tools/gpu_profile.sh: full_TAG_raw.csv): one line per launch
HBM_PEAK_GBS = 6558.0        # MEASURED_PEAKS.json hbm_gbs of this pool (copy bandwidth)
try:
    rows = list(csv.reader(open(G + f'full_{tag}_raw.csv')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, to=None):
        i = col.get(name)
        if i is None or r[i] in ('', 'n/a'):
            return None
        v = float(r[i].replace(',', ''))
        if to == 'MB':
            v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[units[i]]
        elif to == 'us':
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}[units[i]]
        elif to == 'KB':
            v *= {'byte': 1 / 1024, 'Kbyte': 1.0, 'Mbyte': 1024.0}[units[i].split('/')[0]]
        return v

    out = dict(what='ncu --set full --clock-control none --import-source on over the first eager G+D step of tools/step_once.py '
                    '(cfg3, graphs and side streams off), the launches of the step\'s main kernel families; hbm_frac = DRAM '
                    f'bytes / duration / {HBM_PEAK_GBS:.0f} GB/s (measured copy bandwidth)',
               made_by=f'tools/gpu_profile.sh (gpurun_out/full_{tag}_raw.csv) + tools/summarize_ncu.py', launches=[])
    for r in rows[2:]:
        us = val(r, 'gpu__time_duration.sum', 'us')
        rd_, wr_ = val(r, 'dram__bytes_read.sum', 'MB') or 0.0, val(r, 'dram__bytes_write.sum', 'MB') or 0.0
        gbps = (rd_ + wr_) / us * 1e3 if us else None
        rnd2 = lambda v, n=1: None if v is None else round(v, n)
        out['launches'].append(dict(
            kernel=re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('pg::', ''),
            grid=r[col['Grid Size']], block=r[col['Block Size']], us=rnd2(us, 2), dram_read_MB=rnd2(rd_, 2),
            dram_write_MB=rnd2(wr_, 2), dram_GBps=rnd2(gbps), hbm_frac=rnd2(gbps / HBM_PEAK_GBS if gbps else None, 3),
            tensor_pipe_pct=rnd2(val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')),
            warps_active_pct=rnd2(val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')),
            regs=rnd2(val(r, 'launch__registers_per_thread'), 0),
            dyn_smem_KB=rnd2(val(r, 'launch__shared_mem_per_block_dynamic', 'KB')),
            l2_hit_pct=rnd2(val(r, 'lts__t_sector_hit_rate.pct')),
            xbar2l1_read_MB=rnd2(val(r, 'l1tex__m_xbar2l1tex_read_bytes.sum', 'MB'), 2)))
    json.dump(out, open(f'profiles/r{rnd}_ncu_full_step_kernels.json', 'w'), indent=0)
    print('full capture:', len(out['launches']), 'launches')
except FileNotFoundError:
    print('no full capture of the step kernels')
