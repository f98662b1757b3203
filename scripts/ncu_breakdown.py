"""Source-level breakdown of an ncu report: python scripts/ncu_breakdown.py <rep> [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]; ia = h.index('Source'); ie = h.index('Instructions Executed'); iad = h.index('Address'); isamp = h.index('# Samples')
data = []
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit(): continue
    data.append((r[iad][-5:], r[ia].strip(), int(r[ie]), int(r[isamp])))
tot_s = sum(d[3] for d in data); tot_e = sum(d[2] for d in data)
b = collections.Counter(); bs = collections.Counter(); bn = collections.Counter()
for a, t, e, s in data:
    b[e] += e; bs[e] += s; bn[e] += 1
print('total samples', tot_s, 'total exec', tot_e, 'static', len(data))
for k, v in sorted(bs.items(), key=lambda x: -x[1])[:12]:
    print(f'exec/instr={k:9d} n_static={bn[k]:5d} exec_share={100*b[k]/tot_e:5.1f}% sample_share={100*v/tot_s:5.1f}%')
print('--- top sampled instructions')
for a, t, e, s in sorted(data, key=lambda d: -d[3])[:top]:
    print(a, f'{100*s/tot_s:5.2f}%', e, t[:90])
