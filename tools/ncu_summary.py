"""Summarise an .ncu-rep (raw + source pages) into text.  Usage: python tools/ncu_summary.py rep [kernel-regex]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('====', d['Kernel Name'])
    for k in want:
        if k in d: print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
    st = [(k.replace('smsp__pcsamp_warps_issue_stalled_', ''), int(d[k])) for k in hdr if k.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in k and d[k].isdigit()]
    tot = sum(v for _, v in st)
    print('  stalls:', ', '.join(f"{k} {100*v/tot:.0f}%" for k, v in sorted(st, key=lambda kv: -kv[1])[:9]))
pat = sys.argv[2] if len(sys.argv) > 2 else None
if pat:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None; byop = collections.Counter(); tot = 0; samp = collections.Counter()
    for r in rows:
        if r and r[0] == 'Address': h = r; continue
        if h is None or len(r) < len(h): continue
        try: e = int(r[h.index('Instructions Executed')]); s = int(r[h.index('# Samples')])
        except ValueError: continue
        toks = r[h.index('Source')].split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        byop[op] += e; tot += e; samp[op] += s
    print('---- dynamic instruction mix for', pat, 'total warp instr', tot)
    print('  ' + ', '.join(f"{op} {100*c/tot:.1f}%" for op, c in byop.most_common(16)))
    print('  samples: ' + ', '.join(f"{op} {c}" for op, c in samp.most_common(10)))
