"""Summarise an `ncu --csv --page raw` file: key metrics per kernel."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "sm__ctas_launched.sum",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum"]
for f in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(f) if l.startswith('"')))
    if len(rows) < 3:
        print(f, "no data"); continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {f}: {d.get('Kernel Name','')[:40]} grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in hdr:
            if any(k == kk for kk in KEYS) or "tensor" in k and "pct" in k:
                print(f"   {k} = {d[k]} {units[hdr.index(k)]}")
