"""Print the key metrics of an .ncu-rep (run here, no GPU needed): python tools/ncu_key.py file.ncu-rep"""
import csv, subprocess, sys
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_requests_srcunit_tex_op_red.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum', 'smsp__warp_issue_stalled_long_scoreboard',
        'lts__t_sectors_srcunit_tex_op_red.sum']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    print('==', f)
    for row in r[2:]:
        for h, v, un in zip(hdr, row, units):
            if h == 'Kernel Name':
                print(' --', v[:80])
            if any(h == w or h.startswith(w + '.') and h.count('.') == w.count('.') for w in want) or h in want:
                print('   %-86s %s %s' % (h, v, un))
