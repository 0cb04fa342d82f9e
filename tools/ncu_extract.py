"""Key metrics of every kernel launch in an .ncu-rep (`ncu --set full`), as JSON lines.
usage: python tools/ncu_extract.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    'gpu__time_duration.sum': 'us',
    'dram__bytes_read.sum': 'dram_read_MB',
    'dram__bytes_write.sum': 'dram_write_MB',
    'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg': 'tensor_active_cycles_per_tpc',
    'sm__cycles_elapsed.max': 'sm_cycles',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'smsp__inst_executed.sum': 'warp_instructions',
    'lts__t_sectors_srcunit_tex_op_read.sum': 'l2_read_sectors',
    'lts__t_sectors_srcunit_tex_op_write.sum': 'l2_write_sectors',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wavefronts',
    'launch__registers_per_thread': 'regs',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
}
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        d = {'report': rep.split('/')[-1], 'kernel': r[H.index('Kernel Name')][:70], 'grid': r[H.index('Grid Size')], 'block': r[H.index('Block Size')]}
        for h, u, v in zip(H, U, r):
            hk = h if h in KEYS else next((k for k in KEYS if h.endswith('.' + k)), None)
            if hk:
                try:
                    x = float(v.replace(',', ''))
                except ValueError:
                    continue
                if u == 'Mbyte' or u == 'us' or u == '%' or u == '':
                    pass
                elif u == 'Kbyte':
                    x /= 1e3
                elif u == 'Gbyte':
                    x *= 1e3
                elif u == 'byte':
                    x /= 1e6
                elif u == 'ms':
                    x *= 1e3
                elif u == 'ns':
                    x /= 1e3
                d[KEYS[hk]] = round(x, 3)
        if 'tensor_active_cycles_per_tpc' in d and 'sm_cycles' in d:
            d['tensor_pipe_active_frac'] = round(d['tensor_active_cycles_per_tpc'] / 2.0 / d['sm_cycles'], 3)
        print(json.dumps(d))
