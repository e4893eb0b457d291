"""Per-opcode / per-region view of `ncu --page source --csv --print-source sass` for one kernel launch."""
import csv, sys, subprocess, collections
rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = []; blocks.append(cur); continue
    if r and r[0] == 'Address': hdr = r; continue
    if cur is not None and len(r) > 6: cur.append(r)
b = blocks[which]
i_s = hdr.index('# Samples'); i_e = hdr.index('Instructions Executed'); i_src = hdr.index('Source')
tot_s = sum(int(r[i_s]) for r in b); tot_e = sum(int(r[i_e]) for r in b)
print('instructions', len(b), 'samples', tot_s, 'warp-inst executed', tot_e)
op = collections.Counter(); ops = collections.Counter()
for r in b:
    o = r[i_src].split()
    o = [x for x in o if not x.startswith('@')][0].split('.')[0]
    op[o] += int(r[i_e]); ops[o] += int(r[i_s])
for o, e in op.most_common(18):
    print(f'  {o:10s} exec {100*e/tot_e:5.1f}%  samples {100*ops[o]/tot_s:5.1f}%')
if len(sys.argv) > 3:
    step = int(sys.argv[3])
    for a in range(0, len(b), step):
        e = sum(int(r[i_e]) for r in b[a:a+step]); s = sum(int(r[i_s]) for r in b[a:a+step])
        print(f'  [{a:5d}] exec {100*e/tot_e:5.1f}% samples {100*s/tot_s:5.1f}%  {b[a][i_src].strip()[:50]}')
