#!/usr/bin/env python
"""Executed-instruction / stall-sample share per block of B consecutive SASS instructions (first kernel instance)."""
import csv
import subprocess
import sys

rep, B = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 48
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
body, n = [], 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        n += 1
        if n == 2:
            break
    elif r and r[0] == 'Address':
        hdr = r
    elif r and r[0].startswith('0x'):
        body.append(r)
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iE]) for r in body)
ts = sum(int(r[iSamp]) for r in body)
for i in range(0, len(body), B):
    blk = body[i:i + B]
    e = sum(int(r[iE]) for r in blk)
    s = sum(int(r[iSamp]) for r in blk)
    mx = max(int(r[iE]) for r in blk)
    print(f'{i:5d} exec {e * 100 / tot:5.1f}%  samples {s * 100 / ts:5.1f}%  maxexec {mx:8d}  first: {blk[0][iS].strip()[:60]}')
print(len(body), tot, ts)
