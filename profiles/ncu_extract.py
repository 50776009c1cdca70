"""Print the metrics we track from an `ncu --page raw --csv` dump (one block per profiled launch)."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.per_cycle_active', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:]
for r in rows[2:]:
    print('----')
    for w in WANT + extra:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:75s} {r[i][:70]:>24s} {units[i]}")
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if v >= 5.0:
                print(f"  stall {h[len('smsp__average_warp'):][:60]:66s} {v:10.1f} %")
