#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): key metrics + per-SASS-region share / threads per instruction.
usage: tools/ncu_summary.py <report.ncu-rep> [out.txt]"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__thread_inst_executed_pred_on_per_inst_executed.ratio',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'smsp__inst_executed_op_branch.sum', 'sm__cycles_elapsed.avg', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum.per_second', 'l1tex__t_bytes.sum.per_second', 'sm__inst_executed.avg.per_cycle_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'launch__shared_mem_config_size', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct']


def page(rep, name):
    return list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout)))


def main():
    rep = sys.argv[1]
    out = []
    raw = page(rep, 'raw')
    hdr, units = raw[0], raw[1]
    for vals in raw[2:]:
        out.append(f"== {vals[hdr.index('Kernel Name')]}")
        for i, h in enumerate(hdr):
            if h in KEYS or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) >= 0.15):
                out.append(f"{h:88s} {units[i]:16s} {vals[i]}")
    src = page(rep, 'source')
    # one block per kernel and view ("Kernel Name" row, header row, instruction rows); the first block of a kernel is the SASS view
    blocks, cur = [], None
    for row in src:
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = row
        elif cur is not None and len(row) == len(cur['hdr']):
            cur['rows'].append(row)
    seen = set()
    for b in blocks:
        if b['name'] in seen or not b['rows']:
            continue
        seen.add(b['name'])
        h, data = b['hdr'], b['rows']
        iex, ith, isrc, ismp = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Source'), h.index('# Samples')
        stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and not n.endswith('(Not Issued)')]
        tot = sum(int(r[iex]) for r in data) or 1
        tth = sum(int(r[ith]) for r in data)
        out.append(f"\n== {b['name']}: SASS regions (40 instructions each): share of executed warp-instructions, threads per instruction, stall samples, top stalls; overall {tth / tot:.2f} threads/instr")
        for s in range(0, len(data), 40):
            blk = data[s:s + 40]
            ex = sum(int(r[iex]) for r in blk); th = sum(int(r[ith]) for r in blk); smp = sum(int(r[ismp]) for r in blk)
            ops = {}
            for r in blk:
                t = r[isrc].split()
                op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).split('.')[0]
                ops[op] = ops.get(op, 0) + 1
            top = ' '.join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:6])
            st = sorted(((sum(int(r[i] or 0) for r in blk), n[6:]) for i, n in stall_cols), reverse=True)[:3]
            stx = ' '.join(f"{n}:{v}" for v, n in st if v)
            if ex / tot >= 0.005:
                out.append(f"  sass[{s:4d}..] share {100 * ex / tot:5.1f}%  thr/instr {th / max(ex, 1):5.1f}  samples {smp:7d}  {top} | {stx}")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + "\n")


if __name__ == '__main__':
    main()
