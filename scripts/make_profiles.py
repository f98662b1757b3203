"""Turn the raw gpurun_out/ artifacts of a measurement run into the committed summaries under profiles/.
usage: python scripts/make_profiles.py <tag>   (expects gpurun_out/{bench_TAG.json, launches_TAG.csv, em_ws_TAG.ncu-rep, ...})"""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1]
G, P = 'gpurun_out', 'profiles'
os.makedirs(P, exist_ok=True)
for name in (f'bench_{tag}.json', f'bench_reference_{tag}.json', f'configs_{tag}.txt', f'launches_{tag}.csv'):
    src = os.path.join(G, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, name))
lc = os.path.join(G, f'launches_{tag}.csv')
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split('(')[0].replace('void ', '')
        ns = float(r[vi].replace(',', '')) * (1e3 if r[ui] == 'us' else 1e6 if r[ui] == 'ms' else 1)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f'launches_{tag}_summary.txt'), 'w') as f:
        f.write('ncu --metrics gpu__time_duration.sum --clock-control none -c 80 python bench.py --steps 2 --warmup 1 --no-cpu\n')
        f.write('(cold-cache, serialised per-launch times: compare SHARES, not absolutes; first 80 launches)\n\n')
        f.write(f'{"kernel":70s} {"launches":>8s} {"total_us":>12s} {"share":>7s}\n')
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f'{k[:70]:70s} {n:8d} {t/1e3:12.1f} {100*t/tot:6.1f}%\n')
rep = os.path.join(G, f'em_ws_{tag}.ncu-rep')
if os.path.exists(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); h, u, v = rows[0], rows[1], rows[2]
    want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max',
            'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
            'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
            'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
    sel = {}
    lines = []
    for k, uu, x in zip(h, u, v):
        if k in want or ('issue_stalled' in k and 'ratio' in k):
            lines.append(f'{k:95s} {x:>18s} {uu}')
            try: sel[k] = (float(x.replace(',', '')), uu)
            except ValueError: pass
    det = subprocess.run(['ncu', '-i', rep, '--page', 'details'], capture_output=True, text=True).stdout
    det = '\n'.join(l for l in det.splitlines() if l.strip() and not l.strip().startswith(('OPT', 'INF', 'Est.', '---')))
    with open(os.path.join(P, f'em_ws_{tag}_ncu_full.txt'), 'w') as f:
        f.write('ncu --set full --clock-control none --import-source on -k regex:em_ws_kernel -c 1 python scripts/one_fit.py 100\n')
        f.write('one launch = 100 EM iterations of C2 (F=513 T=500 D=8 K=3, complex128); captured on NVIDIA B200 via gpurun\n')
        f.write('(numbers under ncu are NOT bench values)\n\n== selected raw metrics ==\n' + '\n'.join(lines) + '\n\n== details page ==\n' + det[:12000] + '\n')
    def by(k):
        val, unit = sel[k]
        mult = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1}.get(unit, 1)
        return val * mult
    m = {'kernel': 'em_ws_kernel<3,double2>', 'launch': '100 EM iterations, C2',
         'dram_bytes_read': by('dram__bytes_read.sum'), 'dram_bytes_write': by('dram__bytes_write.sum'),
         'duration_ms_under_ncu': sel['gpu__time_duration.sum'][0],
         'fp64_pipe_active_pct': sel['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'][0],
         'registers_per_thread': sel['launch__registers_per_thread'][0],
         'grid': sel['launch__grid_size'][0], 'block': sel['launch__block_size'][0]}
    m['traffic_bytes_per_launch'] = m['dram_bytes_read'] + m['dram_bytes_write']
    json.dump(m, open(os.path.join(P, 'em_kernel_metrics.json'), 'w'), indent=1)
    bd = subprocess.run([sys.executable, 'scripts/ncu_breakdown.py', rep, '25'], capture_output=True, text=True).stdout
    open(os.path.join(P, f'em_ws_{tag}_source_breakdown.txt'), 'w').write(bd)
    print(json.dumps(m))

# ---- round 2: text pages made on the GPU box (reports too large to bring back) ----
def _details_trim(path, limit):
    det = open(path).read()
    return '\n'.join(l for l in det.splitlines() if l.strip() and not l.strip().startswith(('OPT', 'INF', 'Est.', '---')))[:limit]

def _raw_select(path, want_sub):
    rows = list(csv.reader(open(path)))
    rows = [r for r in rows if len(r) > 20]
    h, u = rows[0], rows[1]
    out = []
    for v in rows[2:]:
        name = v[h.index('Kernel Name')] if 'Kernel Name' in h else '?'
        out.append('-- ' + name.split('(')[0])
        for k, uu, x in zip(h, u, v):
            if any(s in k for s in want_sub):
                out.append(f'{k:95s} {x:>18s} {uu}')
    return '\n'.join(out)

WANT = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'issue_stalled_long_scoreboard_per', 'issue_stalled_barrier_per', 'issue_stalled_short_scoreboard_per',
        'issue_stalled_wait_per', 'issue_stalled_math_pipe', 'issue_stalled_sleeping', 'issue_stalled_membar', 'issue_stalled_branch']
cwd = os.path.join(G, f'cw_{tag}_details.txt')
if os.path.exists(cwd):
    with open(os.path.join(P, f'cw_{tag}_ncu_full.txt'), 'w') as f:
        f.write('ncu --set full --clock-control none --import-source on -k regex:em_persistent_kernel -c 1 python scripts/one_fit_cw.py 50\n')
        f.write('one launch = 50 EM iterations of C4 (complex Watson, F=257 T=1000 D=6 K=4, complex128): em_persistent_kernel<6,4,double2,false,2,1>,\n'
                '4 warps per CTA (3 slot-group warps + 1 update-only warp); captured on NVIDIA B200 via gpurun (numbers under ncu are NOT bench values)\n\n')
        f.write('== selected raw metrics ==\n' + _raw_select(os.path.join(G, f'cw_{tag}_raw.csv'), WANT) + '\n\n')
        f.write('== source-level breakdown (scripts/ncu_breakdown.py) ==\n' + open(os.path.join(G, f'cw_{tag}_source_breakdown.txt')).read() + '\n')
        f.write('== details page ==\n' + _details_trim(cwd, 9000) + '\n')
pfd = os.path.join(G, f'postfit_{tag}_details.txt')
if os.path.exists(pfd):
    with open(os.path.join(P, f'postfit_{tag}_ncu.txt'), 'w') as f:
        f.write("ncu --set full --clock-control none -k regex:'dhtv_cluster_kernel|em_fast_kernel|gev_kernel|apply_bf_kernel' -c 4 python scripts/run_c3.py --iterations 5\n")
        f.write('the post-fit kernels of the C3 pipeline (predict = em_fast_kernel, DHTV alignment, GEV beamformer, apply), first launch of each;\n'
                'captured on NVIDIA B200 via gpurun (numbers under ncu are NOT bench values)\n\n')
        f.write('== selected raw metrics ==\n' + _raw_select(os.path.join(G, f'postfit_{tag}_raw.csv'), WANT) + '\n\n')
        det = open(pfd).read()
        # per kernel: header + speed-of-light + launch statistics + occupancy sections only
        keep, on = [], False
        for l in det.splitlines():
            s_ = l.strip()
            if s_.startswith('void ') or s_.startswith('pbb::') or 'Context 1, Stream' in l:
                keep.append(l); continue
            if s_.startswith('Section:'):
                on = any(k in s_ for k in ('Speed Of Light Throughput', 'Launch Statistics', 'Occupancy', 'Warp State'))
            if on and s_ and not s_.startswith(('OPT', 'INF', 'Est.', '---')):
                keep.append(l)
        f.write('== details page (speed of light, launch statistics, occupancy, warp state) ==\n' + '\n'.join(keep)[:16000] + '\n')
for name in (f'pytest_gpu_{tag}.txt',):
    src = os.path.join(G, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, name))
std = os.path.join(G, f'sticky_{tag}_details.txt')
if os.path.exists(std):
    with open(os.path.join(P, f'sticky_{tag}_ncu.txt'), 'w') as f:
        f.write('ncu --set full --clock-control none -k regex:em_sticky_kernel -c 1 python scripts/one_fit_small.py\n')
        f.write('one launch = 100 EM iterations of a C3 shard (F=65 T=500 D=8 K=3, complex128): 65 clusters of 4 CTAs, one bin each;\n'
                'captured on NVIDIA B200 via gpurun (numbers under ncu are NOT bench values)\n\n')
        f.write('== selected raw metrics ==\n' + _raw_select(os.path.join(G, f'sticky_{tag}_raw.csv'), WANT) + '\n\n')
        f.write('== details page ==\n' + _details_trim(std, 9000) + '\n')
