"""Group a compute-sanitizer racecheck log by (hazard kind, first access site, second access site).
usage: python scripts/racecheck_summary.py LOG > summary"""
import collections, re, sys
t = open(sys.argv[1]).read()
c = collections.Counter()
for b in t.split('========= Error:')[1:] + t.split('========= Warning:')[1:]:
    kind = re.search(r'Potential (\w+) hazard', b)
    sites = re.findall(r'(Write|Read) Thread .*? at (?:void )?(?:pbb::)?(\w+).*? in (\S+:\d+)', b)
    c[(kind.group(1) if kind else '?',) + tuple('%s %s (%s)' % (a, loc, fn) for a, fn, loc in sites[:2])] += 1
m = re.search(r'RACECHECK SUMMARY: (.*)', t)
print('compute-sanitizer --tool racecheck summary line:', m.group(1) if m else 'none')
print('hazard reports grouped by access pair (count, kind, first access, second access):')
for k, v in c.most_common():
    print('%8d  %s' % (v, '  |  '.join(k)))
