"""Turns an .ncu-rep into the markdown summary committed under profiles/ (run where ncu is installed).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.md "title" """
import csv
import io
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.avg']


def main(rep, out, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keep = KEEP + [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    with open(out, 'w') as f:
        f.write(f"# {title}\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')]}\n\n| metric | value |\n|---|---|\n")
            for k in keep:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} {units[i]} |\n")
            f.write("\n")


if __name__ == '__main__':
    main(*sys.argv[1:4])
