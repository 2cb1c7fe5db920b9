#!/usr/bin/env python
"""Turn the outputs of tools/gpu_profile.sh (gpurun_out/<tag>_*) into the tracked
summaries under profiles/: per-kernel ncu metrics, hot source lines, launch-list
shares, the bench JSON lines and profiles/ncu_traffic.json (DRAM bytes per launch
that bench.py reports as roofline.traffic).

  python tools/summarize_profiles.py r01
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
STALL = 'smsp__average_warps_issue_stalled_'


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def raw_page(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def short(name):
    n = name.replace('brcnn::', '').replace('void ', '')
    return n.split('(')[0].split('<')[0]


def summarize(tag, mode, traffic):
    rep = os.path.join(OUT, f'{tag}_full_{mode}.ncu-rep')
    if not os.path.exists(rep):
        return
    hdr, units, rows = raw_page(rep)
    lines = [f'# ncu --set full --clock-control none, bench.py mode={mode} (one step after 3 warm-ups)',
             f'# source: gpurun_out/{tag}_full_{mode}.ncu-rep (tools/gpu_profile.sh)', '']
    seen = {}
    for r in rows:
        name = r[hdr.index('Kernel Name')]
        k = short(name)
        seen[k] = seen.get(k, 0) + 1
        lines.append(f'## {k}  (launch {seen[k]})')
        rd = wr = 0.0
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith(STALL) and h.endswith('_per_issue_active.ratio')):
                lines.append(f'  {h} = {r[i]} {units[i]}')
            if h == 'dram__bytes_read.sum':
                rd = to_bytes(r[i], units[i])
            if h == 'dram__bytes_write.sum':
                wr = to_bytes(r[i], units[i])
        traffic.setdefault(k, []).append(rd + wr)
        lines.append('')
    # hot lines of the main kernels
    tool = os.path.join(ROOT, 'tools', 'ncu_hot_lines.py')
    for k in seen:
        res = subprocess.run([sys.executable, tool, rep, k, '12'], capture_output=True, text=True)
        lines.append(f'## hot source lines: {k}')
        lines.append(res.stdout.rstrip())
        lines.append('')
    open(os.path.join(PROF, f'{tag}_ncu_full_{mode}.txt'), 'w').write('\n'.join(lines) + '\n')


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    os.makedirs(PROF, exist_ok=True)
    tool = os.path.join(ROOT, 'tools', 'ncu_launch_summary.py')
    for mode in ('infer', 'train'):
        src = os.path.join(OUT, f'{tag}_launches_{mode}.csv')
        if os.path.exists(src):
            res = subprocess.run([sys.executable, tool, src], capture_output=True, text=True)
            open(os.path.join(PROF, f'{tag}_launches_{mode}.summary.txt'), 'w').write(
                f'# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py mode={mode}\n'
                + res.stdout)
            # keep the raw list too (kernel name, duration)
            rows = [l for l in open(src) if not l.startswith('==')]
            with open(os.path.join(PROF, f'{tag}_launches_{mode}.csv'), 'w') as f:
                rd = csv.DictReader(rows)
                f.write('id,kernel,duration,unit\n')
                for row in rd:
                    f.write(f"{row['ID']},\"{row['Kernel Name'][:90]}\",{row['Metric Value']},{row['Metric Unit']}\n")
    traffic_all = {}
    if os.path.exists(os.path.join(PROF, 'ncu_traffic.json')):
        traffic_all = json.load(open(os.path.join(PROF, 'ncu_traffic.json')))
    for mode, key in (('infer', 'utdac_b16'), ('train', 'coco_train_b2')):
        traffic = {}
        summarize(tag, mode, traffic)
        if traffic:
            traffic_all[key] = {k: sum(v) / len(v) for k, v in traffic.items()}
            traffic_all[key]['_source'] = f'profiles/{tag}_ncu_full_{mode}.txt'
    json.dump(traffic_all, open(os.path.join(PROF, 'ncu_traffic.json'), 'w'), indent=1, sort_keys=True)
    for name in ('bench_infer.json', 'bench_train.json', 'bench_stress.json',
                 'bench_stress_clustered.json', 'bench_stress_anchors.json', 'bench_voc1000.json',
                 'bench_coco_infer.json', 'bench_reference.json', 'nms_microbench.txt',
                 'roi_microbench_infer.json', 'roi_microbench_train.json'):
        src = os.path.join(OUT, f'{tag}_{name}')
        if os.path.exists(src) and os.path.getsize(src) > 0:
            text = open(src).read()
            if name.endswith('.json'):      # keep the JSON line only (NCCL / warnings go to stdout too)
                js = [l for l in text.splitlines() if l.startswith('{')]
                text = (js[-1] + '\n') if js else text
            open(os.path.join(PROF, f'{tag}_{name}'), 'w').write(text)
    smi = os.path.join(OUT, f'{tag}_nvidia_smi.csv')
    if os.path.exists(smi):
        open(os.path.join(PROF, f'{tag}_nvidia_smi.csv'), 'w').write(open(smi).read())
    print('profiles/ updated:', sorted(os.listdir(PROF)))


if __name__ == '__main__':
    main()
