"""Turns the raw ncu outputs brought back in gpurun_out/ into the small tracked files of profiles/.

  python profiles/summarise.py launches gpurun_out/launches.csv profiles/rNN_launches.md
  python profiles/summarise.py full gpurun_out/prof.ncu-rep profiles/rNN_ncu_full.csv
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration', 'dram__bytes', 'dram__throughput', 'sm__throughput', 'warps_active',
        'registers_per_thread', 'pipe_fma', 'pipe_tensor', 'bank_conflicts', 'wavefronts_mem_shared', 'grid_size', 'block_size',
        'occupancy', 'issue_active', 'lts__t_sector_hit', 'l1tex__t_sector_hit', 'shared_mem_per_block', 'sm__cycles_elapsed.max',
        'launch__waves', 'smsp__inst_executed.sum', 'lts__throughput', 'l1tex__throughput']


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) < 15:
            continue
        name = re.sub(r'\(umma::GemmArgs.*', '', r[4])
        name = re.sub(r'\(.*', '', name).replace('<unnamed>::', '')
        tot[name][0] += 1
        tot[name][1] += float(r[-1])
    total = sum(v[1] for v in tot.values())
    with open(dst, 'w') as f:
        f.write('ncu --metrics gpu__time_duration.sum --clock-control none (serialised: compare SHARES; see the file header in profiles/README.md for the cache-control mode)\n\n')
        f.write('%d launches, %.3f ms total\n\n| launches | total ms | share | avg us | kernel |\n|---|---|---|---|---|\n' % (sum(v[0] for v in tot.values()), total / 1e6))
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write('| %d | %.3f | %.1f %% | %.1f | `%s` |\n' % (v[0], v[1] / 1e6, 100 * v[1] / total, v[1] / v[0] / 1e3, k[:150]))


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEEP)]
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in keep])
        w.writerow([units[i] for i in keep])
        for r in rows[2:]:
            w.writerow([r[i] for i in keep])


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
