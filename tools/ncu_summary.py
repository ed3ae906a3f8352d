"""Summarise an .ncu-rep (one line of key metrics per captured kernel) into JSON for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.json"""
import csv
import json
import subprocess
import sys

KEYS = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'dram__bytes_read.sum.pct_of_peak_sustained_elapsed': 'dram_read_pct',
    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active': 'dmma_pipe_pct_active',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed': 'tensor_pipe_pct_elapsed',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'launch__registers_per_thread': 'regs',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'launch__occupancy_limit_shared_mem': 'occ_limit_smem',
    'launch__occupancy_limit_registers': 'occ_limit_regs',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
    'sass__inst_executed_local_loads': 'local_loads',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'stall_math_pipe',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'stall_wait',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_sb',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_sb',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio': 'stall_no_inst',
    'sm__cycles_elapsed.max': 'cycles',
}


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        rec = {'kernel': vals[hdr.index('Kernel Name')][:80]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                try:
                    rec[KEYS[h]] = float(v.replace(',', ''))
                except ValueError:
                    rec[KEYS[h]] = v
                if u and KEYS[h] in ('duration', 'dram_read', 'dram_write'):
                    rec[KEYS[h] + '_unit'] = u
        res.append(rec)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
