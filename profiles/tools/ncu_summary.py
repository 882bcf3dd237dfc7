"""Key metrics of every launch in an `ncu --set full` report as CSV.
usage: python profiles/tools/ncu_summary.py <report.ncu-rep>"""
import csv
import subprocess
import sys

WANT = [('Kernel Name', 'kernel'), ('Grid Size', 'grid'), ('gpu__time_duration.sum', 'duration_us'),
        ('dram__bytes_read.sum', 'dram_read_MB'), ('dram__bytes_write.sum', 'dram_write_MB'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_active_pct'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_throughput_pct'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_throughput_pct'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('launch__registers_per_thread', 'regs'), ('launch__occupancy_limit_shared_mem', 'ctas_per_sm_smem_limit')]


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(h, n) for h, n in WANT if h in idx]
    w = csv.writer(sys.stdout)
    w.writerow([n for _, n in cols])
    for r in rows[2:]:
        w.writerow([r[idx[h]][:90] for h, _ in cols])


if __name__ == '__main__':
    main()
