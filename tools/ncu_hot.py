"""Instruction mix and most-sampled SASS of the first kernel in an ncu report:
python tools/ncu_hot.py <file.ncu-rep> [top_n]"""
import collections
import csv
import io
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
sec = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
print(sec[0][1][:120])
hdr = sec[1]
ia, ie, isamp, iad = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address')
body = [r for r in sec[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body)
tots = sum(int(r[isamp]) for r in body)
byop, samp = collections.Counter(), collections.Counter()
for r in body:
    src = r[ia]
    op = (src.split()[1] if src.startswith('@') else src.split()[0]).split('.')[0]
    byop[op] += int(r[ie])
    samp[op] += int(r[isamp])
print('warp instructions', tot, 'samples', tots)
for op, n in byop.most_common(14):
    print('%-8s %5.1f%% instr  %5.1f%% samples' % (op, 100 * n / tot, 100 * samp[op] / tots))
print('--- most sampled')
for r in sorted(body, key=lambda r: -int(r[isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('%s %9d %6s  %s' % (r[iad][-5:], int(r[ie]), r[isamp], r[ia][:90]))
