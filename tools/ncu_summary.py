"""Summarise an .ncu-rep (read here, no GPU): key launch metrics, stall samples, executed-instruction mix."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for d in data[:1]:
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print("%-72s %s %s" % (w, d[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h:
            try:
                st.append((float(d[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("stall samples:", ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:9]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if hi:
    h = rows[hi[0]]
    end = hi[1] - 1 if len(hi) > 1 else len(rows)
    ci, si = h.index('Instructions Executed'), h.index('Source')
    byop, tot = collections.Counter(), 0
    for r in rows[hi[0] + 1:end]:
        try:
            n = int(r[ci])
        except (ValueError, IndexError):
            continue
        t = r[si].split()
        op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '')).split('.')[0]
        byop[op] += n
        tot += n
    nwarp = float(sys.argv[2]) if len(sys.argv) > 2 else 32768.0
    print("warp-instructions %d  (%.1f per warp-step at %d warps)" % (tot, tot / nwarp, nwarp))
    print("  ".join("%s %.1f" % (op, n / nwarp) for op, n in byop.most_common(26)))
