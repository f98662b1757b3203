"""profiles/sanitizer_r2.txt from gpurun_out/{memcheck,racecheck}_r2.txt (compute-sanitizer over scripts/sanitize_small.py)."""
import re
t = open('gpurun_out/racecheck_r2.txt').read()
pairs = {}
for b in t.split('========= Error:')[1:]:
    first = re.search(r'Race reported between (\w+) access at .*? in (\S+:\d+)', b)
    for a, loc, n in re.findall(r'and (\w+) access at .*? in (\S+:\d+) \[(\d+) haz', b):
        key = tuple(sorted([first.group(1) + ' ' + first.group(2), a + ' ' + loc]))
        pairs[key] = pairs.get(key, 0) + int(n)
out = ['compute-sanitizer on scripts/sanitize_small.py (cACGMM on em_sticky_kernel (clusters) and on em_ws_kernel with the frame',
       'split forced (PBB_TSPLIT=2), full variant with saliency, streamed upload, K = 2, 3, 4, complex Watson, generic shapes,',
       'DHTV alignment on the thread-block-cluster kernel), NVIDIA B200 via gpurun, round-2 final code',
       '',
       '--tool memcheck:  ' + open('gpurun_out/memcheck_r2.txt').read().strip().splitlines()[-1].replace('========= ', ''),
       '--tool racecheck --racecheck-report analysis:  ' + re.search(r'RACECHECK SUMMARY: (.*)', t).group(1),
       '',
       'Every racecheck report is one of these access pairs (hazard count, access | access), all inside em_ws_kernel:']
for (a, b), n in sorted(pairs.items(), key=lambda x: -x[1]):
    out.append('  %8d  %s | %s' % (n, a, b))
out += ['',
        'All of them are shared-memory hand-overs between the warp roles of em_ws_kernel that are ordered by an mbarrier',
        '(inline-PTX mbarrier.arrive / mbarrier.try_wait with a "memory" clobber), which racecheck does not model:',
        '  em_ws.cuh:228 (EM warps write the scatter sums S[sb])  ->  mbar_arrive(s_full[sb])  ->  updaters mbar_wait(s_full[sb])  ->  read S[sb]',
        '    (em_ws.cuh:376 frame-split store of the part; cacg_update_class in em_persistent.cuh); the reverse edge is s_empty[sb]',
        '  em_ws.cuh:320 (producer stages the model coef[mb])  ->  mbar_arrive(model_full[mb])  ->  EM warps mbar_wait(model_full[mb])  ->',
        '    lean_chunk2_split reads coef (em_persistent.cuh:486); the reverse edge is model_empty[mb]',
        'No report involves em_sticky_kernel or dhtv_cluster_kernel (cluster barriers, DSMEM exchange), em_persistent_kernel',
        '(CTA barriers / named barrier 1), the frame-split flags, the TMA rings, or any other kernel of the library.']
open('profiles/sanitizer_r2.txt', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
