#!/usr/bin/env python
"""Join the ncu per-launch metrics of one bench step (tools/gpu_ncu.sh pass 1) with the CUDA-event profile
(APTP_PROFILE_DUMP of bench.py) so every grouped-GEMM / attention launch carries its label, event time,
kept FLOPs, DRAM bytes and tensor-pipe activity. Usage: merge_launch_metrics.py launch_metrics.csv kernel_profile.tsv"""
import collections
import csv
import sys

U = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}


def load(metrics, profile):
    rows = [r for r in csv.DictReader(l for l in open(metrics) if not l.startswith('=='))]
    L = collections.OrderedDict()
    for r in rows:
        d = L.setdefault(int(r['ID']), {'name': r['Kernel Name'].split('(')[0].replace('void ', '')})
        d[r['Metric Name'].split('.')[0]] = float(r['Metric Value'].replace(',', '')) * U.get(r['Metric Unit'], 1)
    prof = [l.rstrip('\n').split('\t') for l in open(profile)]
    gem = [p for p in prof if p[0] == 'gemm' and ' tiles0 ' not in p[4]]
    att = [p for p in prof if p[0] == 'attn']
    gl = [d for d in L.values() if 'grouped_gemm' in d['name']]
    al = [d for d in L.values() if 'attention_kernel' in d['name']]
    assert len(gem) == len(gl) and len(att) == len(al), (len(gem), len(gl), len(att), len(al))
    out = []
    for kind, ps, ds in (('gemm', gem, gl), ('attn', att, al)):
        for p, d in zip(ps, ds):
            by = d['dram__bytes_read'] + d['dram__bytes_write']
            out.append(dict(kind=kind, ncu_ms=d['gpu__time_duration'] / 1e6, ev_ms=float(p[1]), gflop=float(p[2]),
                            dram_mb=by / 1e6, tensor_pct=d['sm__pipe_tensor_cycles_active'], label=p[4]))
    return out, L


def main():
    out, L = load(sys.argv[1], sys.argv[2])
    print("kind\tncu_ms\tev_ms\tGFLOP\tTFLOP/s(ev)\tdram_MB\tdram_TB/s(ev)\ttensor%\tlabel")
    for o in sorted(out, key=lambda x: -x['ev_ms']):
        print(f"{o['kind']}\t{o['ncu_ms']:.3f}\t{o['ev_ms']:.3f}\t{o['gflop']:.1f}\t{o['gflop'] / o['ev_ms']:.0f}\t"
              f"{o['dram_mb']:.1f}\t{o['dram_mb'] / o['ev_ms'] / 1e3:.2f}\t{o['tensor_pct']:.1f}\t{o['label']}")
    for kind in ('gemm', 'attn'):
        s = [o for o in out if o['kind'] == kind]
        t = sum(o['ev_ms'] for o in s)
        print(f"# {kind}: {len(s)} launches, {t:.3f} ms (events), {sum(o['ncu_ms'] for o in s):.3f} ms (ncu), "
              f"{sum(o['gflop'] for o in s) / t:.1f} TFLOP/s, DRAM traffic {sum(o['dram_mb'] for o in s) / 1e3:.2f} GB "
              f"({sum(o['dram_mb'] for o in s) / len(s):.1f} MB/launch), time-weighted tensor-pipe active "
              f"{sum(o['ncu_ms'] * o['tensor_pct'] for o in s) / sum(o['ncu_ms'] for o in s):.1f}%")
    oth = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in L.values():
        if 'grouped_gemm' in d['name'] or 'attention_kernel' in d['name']:
            continue
        a = oth[d['name'][:40]]
        a[0] += 1
        a[1] += d['gpu__time_duration'] / 1e6
        a[2] += d['dram__bytes_read'] + d['dram__bytes_write']
    for k, (n, ms, by) in sorted(oth.items(), key=lambda kv: -kv[1][1]):
        print(f"# {k:40s} n={n:3d} {ms:7.3f} ms  dram {by / 1e9:6.2f} GB  {by / ms / 1e9:6.2f} TB/s")


if __name__ == "__main__":
    main()
