"""Summarise `ncu --csv --page raw` files: one compact line-block per kernel. usage: ncu_summ.py file..."""
import csv, sys
KEYS = [("gpu__time_duration.sum", "time_ns"), ("sm__cycles_elapsed.max", "cycles"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar2l1_rd"),
        ("lts__t_sectors_op_read.sum", "l2_rd_sect"), ("lts__t_sectors_op_write.sum", "l2_wr_sect"),
        ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "utchmma_bf16%"),
        ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "utchmma_fp16%"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "mem_tensor_active%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "pipe_tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "smem_dyn"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%")]
for f in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(f) if l.startswith('"')))
    if len(rows) < 3:
        print(f, "no data"); continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {f}: {d.get('Kernel Name','')[:50]}")
        print("   " + "  ".join(f"{n}={d[k]}" for k, n in KEYS if k in d))
