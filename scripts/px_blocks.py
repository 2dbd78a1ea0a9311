"""Summarises an ncu source-page CSV of k_stats_big_exact_px: samples / executed instructions per 48-instruction block
with the markers that identify the warp role (chain: STS.128 + LDS.64, verifier: FSETP.NEU, variance: VIADDMNMX)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc = hdr.index('Address'), hdr.index('Source')
isamp, iex = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = rows[2:]
print('total samples', sum(int(r[isamp] or 0) for r in data), 'instructions', sum(int(r[iex] or 0) for r in data))
blk = 48
for b in range(0, len(data), blk):
    ch = data[b:b + blk]
    smp = sum(int(r[isamp] or 0) for r in ch)
    ex = sum(int(r[iex] or 0) for r in ch)
    if smp == 0 and ex == 0:
        continue
    txt = ' '.join(r[isrc] for r in ch)
    cnt = {m: txt.count(m) for m in ['LDGSTS', 'STS.128', 'FSETP.NEU', 'VIADDMNMX', 'BAR.SYNC', 'MEMBAR', 'ATOMS',
                                      'LDS.128', 'LDS.64', 'FFMA'] if m in txt}
    agg = {}
    for r in ch:
        for i, h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[i] or 0)
    top = dict(sorted(agg.items(), key=lambda kv: -kv[1])[:3])
    print(hex(int(ch[0][ia], 16) & 0xfffff), str(smp).rjust(6), str(ex).rjust(10), cnt, top)
