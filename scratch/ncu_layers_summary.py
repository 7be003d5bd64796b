"""profiles/<tag>_ncu_full_layers_summary.csv and profiles/ncu_traffic.json from the `ncu --set full` capture of
scratch/prof_layers.py (gpurun_out/<tag>_layers.ncu-rep): per-launch DRAM traffic of the headline layers (second
repetition of each), which bench.py copies into roofline.traffic.   usage: ncu_layers_summary.py <tag>"""
import csv, io, json, subprocess, sys
tag = sys.argv[1]
rep = 'gpurun_out/%s_layers.ncu-rep' % tag
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg',
        'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum']
idx = [h.index(k) for k in keep if k in h]
with open('profiles/%s_ncu_full_layers_summary.csv' % tag, 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
labels = ['SubpixelConvolution/conv2x*TransitionLast/conv', 'SubpixelConvolution/conv2x', 'ResidualBlock6/conv2']
sizes = ['64x64', '32x32', '32x32']
ik, ir, iw, it = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'), h.index('gpu__time_duration.sum')
assert rows[1][ir] == 'Mbyte' and rows[1][it] == 'us', (rows[1][ir], rows[1][it])
body = rows[2:]
out = {'_comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from the round-2 closing `ncu --set full` capture of '
                   'scratch/prof_layers.py (batch 64, tf32x3; second repetition of each layer); bench.py copies the entry of its '
                   'dominant kernel into roofline.traffic',
       '_source': 'profiles/%s_ncu_full_layers_summary.csv (gpurun_out/%s_layers.ncu-rep)' % (tag, tag),
       '_duration_us': {}, '_kernel': {}}
for li, (lab, sz) in enumerate(zip(labels, sizes)):
    base = li * 6 + 3                       # second repetition: fwd, wgrad, dgrad
    for j, kind in enumerate(('fwd', 'wgrad', 'dgrad')):
        r = body[base + j]
        key = '%s:%s@%s' % (lab, kind, sz)
        out[key] = int(round((float(r[ir]) + float(r[iw])) * 1e6))
        out['_duration_us'][key] = float(r[it])
        out['_kernel'][key] = r[ik].replace('void ', '').split('(')[0]
json.dump(out, open('profiles/ncu_traffic.json', 'w'), indent=1)
print(json.dumps(out, indent=1)[:1500])
