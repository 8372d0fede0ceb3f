"""SASS evidence for the tcgen05 / TMA kernels: per kernel of libnerf_b200.so, the count of the mnemonics that prove the
Blackwell-native path (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UBLKCP
(cp.async.bulk), UTMALDG (tensor-map loads), SYNCS (mbarrier), plus registers.  usage: python scripts/sass_summary.py > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'nerf_pytorch_paeng_b200', 'libnerf_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pat = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM.x32', 'UTCBAR', 'UBLKCP.S.G', 'UBLKCP.S.G.MULTICAST', 'UTMALDG.2D.2CTA', 'SYNCS', 'FADD2', 'STG.E.128', 'HMMA', 'FFMA']
cur = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(anonymous namespace\)::', '', cur).split('(')[0].replace('void ', '')
        counts[cur] = collections.Counter()
        continue
    if cur is None or '/*' not in line:
        continue
    ins = line.split('*/')[1] if '*/' in line else line
    toks = re.findall(r'[A-Z][A-Z0-9_]*(?:\.[A-Za-z0-9_]+)*', ins)
    op = toks[0] if toks else ''
    for p in pat:
        if op == p or (p in ('UTCHMMA', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA', 'FADD2') and op.split('.')[0] == p):
            counts[cur][p] += 1
print('# SASS summary of libnerf_b200.so (cuobjdump -sass, sm_100a)\n')
print('| kernel | ' + ' | '.join(pat) + ' |\n|---|' + '---|' * len(pat))
for k, c in counts.items():
    if 'mlp_' in k or sum(c[p] for p in pat[:7]):
        print(f'| `{k}` | ' + ' | '.join(str(c[p]) for p in pat) + ' |')
print('\nOther kernels (CUDA-core, HBM-bound): ' + ', '.join(f'`{k}`' for k in counts if not ('mlp_' in k or sum(counts[k][p] for p in pat[:7]))))
res = subprocess.run(['cuobjdump', '-res-usage', so], capture_output=True, text=True).stdout
print('\n## Resource usage (cuobjdump -res-usage), MLP kernels\n')
lines = res.splitlines()
for i, l in enumerate(lines):
    if 'Function' in l and 'mlp_' in l:
        name = subprocess.run(['c++filt', l.split('Function ')[1].rstrip(':')], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', name).split('(')[0].replace('void ', '')
        print(f'* `{name}`: {lines[i + 1].strip()}')
