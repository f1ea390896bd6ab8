#!/usr/bin/env python3
"""Per CUDA source line: share of executed warp-instructions, threads per instruction, stall samples, from an .ncu-rep
captured with --import-source on (read here, no GPU).  usage: tools/ncu_source_lines.py <report.ncu-rep> [min_share_pct]"""
import csv, io, subprocess, sys


def main():
    rep = sys.argv[1]
    floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, {}
    for r in csv.reader(io.StringIO(txt)):
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < 10 or r[2] != '-' or not r[0]:
            continue
        try:
            ex, th, smp = int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('Thread Instructions Executed')]), int(r[hdr.index('# Samples')])
        except ValueError:
            continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, 0, r[1]])
        a[0] += ex; a[1] += th; a[2] += smp
    tot = sum(a[0] for a in agg.values()) or 1
    smp_tot = sum(a[2] for a in agg.values()) or 1
    print(f"total warp-instructions {tot}, samples {smp_tot}")
    for k, a in sorted(agg.items()):
        if 100 * a[0] / tot >= floor:
            print(f"{k[0]:16s}:{k[1]:4d} instr {100 * a[0] / tot:5.1f}%  thr/instr {a[1] / max(a[0], 1):5.1f}  samples {100 * a[2] / smp_tot:5.1f}% | {a[3].strip()[:100]}")


if __name__ == '__main__':
    main()
