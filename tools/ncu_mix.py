"""Summarises an `ncu --page source --csv` dump: instruction mix per opcode and stall samples."""
import csv, collections, sys
path = sys.argv[1]; samples = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = rows[1]
isrc, ie, istall = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
cnt = collections.Counter(); tot = 0; st = collections.Counter(); stall_by_op = collections.Counter()
for r in rows[2:]:
    try: n = int(r[ie])
    except Exception: continue
    toks = r[isrc].strip().split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    cnt[op] += n; tot += n
    for i in stall_cols:
        try: st[hdr[i]] += int(r[i])
        except Exception: pass
    try: stall_by_op[op] += int(r[istall])
    except Exception: pass
print('total warp instr', tot, ' per sample', round(tot * 32 / samples, 2))
for op, n in cnt.most_common(24):
    print(f"{op:10s} {n:>12d} {n*32/samples:8.2f}/sample   stall-samples {stall_by_op[op]}")
tot_st = sum(st.values())
print('stall reasons:', ', '.join(f"{k[6:]} {100*v/tot_st:.1f}%" for k, v in st.most_common(10)))
