"""Extract the judged metrics of one `ncu --set full` capture into a compact markdown row set.

  ncu -i prof.ncu-rep --page raw --csv > prof_raw.csv ;  python tools/ncu_extract.py prof_raw.csv [label] >> profiles/xxx.md

Keeps: duration, tensor-pipe activity, DRAM bytes / throughput, L2 / L1 throughput, registers, shared memory, occupancy."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active (% of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (% of elapsed cycles)"),
    ("sm__inst_executed_pipe_tc.sum", "tensor-core instructions (UTCHMMA)"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-datapath instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 -> SM read sectors (32 B)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput (% of peak)"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts (% of peak)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory operand reads of the tensor core (% of peak)"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active, realtime (% of elapsed)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]


def main():
    path = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [ln for ln in open(path, newline="") if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    hdr, units = rows[0], rows[1]
    print(f"\n### {label}\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"kernel `{d.get('Kernel Name', '?')[:100]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k, name in KEYS:
            for h, u in zip(hdr, units):
                if (h == k or h.endswith("." + k)) and d[h] != "":
                    print(f"| {name} (`{k}`) | {d[h]} | {u} |")
                    break
        print()


if __name__ == "__main__":
    main()
