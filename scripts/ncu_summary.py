"""Summarise an ncu launch list (csv) and a --set full raw page (csv) into profiles/<name>.md + profiles/traffic.json.
usage: python scripts/ncu_summary.py launches.csv raw.csv out.md "<command>" """
import collections
import csv
import json
import os
import sys

launch_csv, raw_csv, out_md, cmd = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
out.append('# ncu evidence (B200, sm_100a)\n')
out.append(f'Command (plain run exited 0 first): `{cmd}`\n')
out.append(f'## Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` ({os.path.basename(launch_csv)})\n')
lines = [l for l in open(launch_csv) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    try:
        t = float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    u = row['Metric Unit']
    t = t / 1e3 if u == 'ns' else (t * 1e3 if u == 'ms' else t)
    agg.setdefault(row['Kernel Name'][:70], []).append(t)
tot = sum(sum(v) for v in agg.values())
out.append('| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|')
mlp_share = 0.
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:12]:
    out.append(f'| `{k}` | {len(v)} | {sum(v) / 1e3:.2f} | {sum(v) / len(v):.1f} | {sum(v) / tot * 100:.1f}% |')
    if 'mlp_' in k:
        mlp_share += sum(v) / tot
out.append(f'\nThe tcgen05 MLP kernels hold {mlp_share * 100:.1f}% of the device time (cold-cache, serialised timings: shares, not absolutes).\n')
out.append('## `ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 18 -c 6` (one train step: coarse net 262,144 points, fine net 786,432 points)\n')
rows = list(csv.reader(open(raw_csv)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = [('gpu__time_duration.sum', 'ms'), ('dram__bytes_read.sum', 'GB read'), ('dram__bytes_write.sum', 'GB written'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('launch__registers_per_thread', 'regs')]
out.append('| kernel | ' + ' | '.join(w[1] for w in want) + ' |\n|' + '---|' * (len(want) + 1))
traffic = collections.defaultdict(list)
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    out.append(f'| `{name}` | ' + ' | '.join(f'{float(r[idx[w[0]]]):.3f}' for w in want) + ' |')
    traffic[name.split('<')[0]].append((float(r[idx['dram__bytes_read.sum']]) + float(r[idx['dram__bytes_write.sum']])) * 1e9)
open(out_md, 'w').write('\n'.join(out) + '\n')
tj = {'source': os.path.basename(out_md) + ': dram__bytes_read.sum + dram__bytes_write.sum, average of the coarse (262,144 pts) and fine (786,432 pts) launches'}
for k, v in traffic.items():
    tj[k + '_bytes_per_launch'] = sum(v) / len(v)
tj['mlp_fwd_chain_kernel_train_bytes_per_launch'] = tj.get('mlp_fwd_chain_kernel_bytes_per_launch')
json.dump(tj, open(os.path.join(root, 'profiles', 'traffic.json'), 'w'), indent=1)
print(open(out_md).read())
