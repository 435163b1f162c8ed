"""Key metrics of every launch in an .ncu-rep (from `ncu --set full`), as plain text for profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {path.split('/')[-1]}  launch {r[h.index('ID')]}  {r[h.index('Kernel Name')][:90]}")
        for k in WANT:
            if k in h:
                i = h.index(k)
                print(f"   {k:75s} {r[i]:>16s} {units[i]}")
