#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: per-opcode executed-instruction histogram and the
hottest SASS instructions (by executed count and by stall samples) of the first kernel instance."""
import csv
import subprocess
import sys
from collections import Counter


def main(rep, topn=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # split into kernel instances
    inst, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = []
            inst.append((r[1], cur))
        elif r and r[0] == 'Address':
            hdr = r
        elif cur is not None and r and r[0].startswith('0x'):
            cur.append(r)
    name, body = inst[0]
    iS, iE, iT, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in body)
    totT = sum(int(r[iT]) for r in body)
    print(f'{name[:80]}\n  SASS lines {len(body)}  warp-instr executed {tot}  thread-instr {totT}  avg active {totT / max(tot, 1):.1f}')
    ops = Counter()
    for r in body:
        op = r[iS].split()[0] if not r[iS].strip().startswith('@') else r[iS].split()[1]
        ops[op.split('.')[0]] += int(r[iE])
    print('  by opcode:', ', '.join(f'{k} {v * 100 / tot:.1f}%' for k, v in ops.most_common(18)))
    # cumulative profile by address order, in 10 chunks of executed instructions
    print('  hottest by samples:')
    for r in sorted(body, key=lambda r: -int(r[iSamp]))[:topn]:
        print(f'    {r[0][-5:]} samp={r[iSamp]:>5} exec={r[iE]:>9} {r[iS].strip()[:90]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
