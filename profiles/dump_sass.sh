#!/bin/bash
# SASS of the step's kernels from the built library (cuobjdump), with an instruction histogram per kernel.
# usage: profiles/dump_sass.sh <tag>     -> profiles/<tag>_sass_<kernel>.txt (+ <tag>_sass_mnemonics.txt)
set -e
cd "$(dirname "$0")/.."
LIB=mars_ode_physics_b200/lib/libmars_ode_b200.so
TAG=${1:-r2}
cuobjdump -sass "$LIB" > /tmp/b2p_all.sass
python3 - "$TAG" <<'PY'
import re, sys, collections
tag = sys.argv[1]
txt = open('/tmp/b2p_all.sass').read().split('\t\tFunction : ')
want = {'sor_tmem_8_8': r'b2p_sor_tmem_kernelILi8ELi8E', 'assemble_islands': r'b2p_assemble_kernelILb1E',
        'integrate': r'b2p_integrate_kernelE', 'collide': r'b2p_collide_kernelILb0E'}
hist_out = []
for key, pat in want.items():
    for blk in txt[1:]:
        name = blk.split('\n', 1)[0]
        if re.search(pat, name):
            lines = []
            hist = collections.Counter()
            for l in blk.split('\n'):
                m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
                if m:
                    ins = m.group(2).strip()
                    lines.append(f'/*{m.group(1)}*/ {ins}')
                    op = re.sub(r'^@!?U?P\d+\s+', '', ins).split()[0]
                    hist[op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDG', 'STG', 'LDS', 'STS', 'LDTM', 'STTM', 'UTMA', 'UBLKCP', 'SYNCS')) and '.' in op else '')] += 1
            open(f'profiles/{tag}_sass_{key}.txt', 'w').write(f'// {name}\n' + '\n'.join(lines) + '\n')
            top = ', '.join(f'{k} {v}' for k, v in hist.most_common(40))
            hist_out.append(f'{key} ({name}): {len(lines)} instructions\n  {top}\n')
            break
open(f'profiles/{tag}_sass_mnemonics.txt', 'w').write('\n'.join(hist_out))
print('\n'.join(hist_out))
PY
