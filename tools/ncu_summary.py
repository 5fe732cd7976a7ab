#!/usr/bin/env python
"""Summaries of ncu captures for profiles/: `ncu -i X.ncu-rep --page raw --csv` -> the counters the north star names
(pipe utilisation, lanes per instruction, L1 / L2 hit rates, DRAM bytes, stall reasons), one column per launch; and
`--page source --csv` -> warp-instructions by opcode class with their average active lanes.

    python tools/ncu_summary.py raw   gpurun_out/r2_cfg2_raw.csv [more_raw.csv ...]
    python tools/ncu_summary.py source gpurun_out/r2_cfg2_source.csv
"""
import csv
import sys
from collections import defaultdict

RAW = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def raw(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        kn = hdr.index("Kernel Name")
        print(f"== {path}")
        names = []
        for d in data:
            n = d[kn]
            n = n[n.find("k_"):] if "k_" in n else n
            names.append(n.split("(")[0][:44])
        print(f"{'metric':92s}" + "".join(f"{n:>46s}" for n in names))
        for m in RAW:
            if m not in hdr:
                continue
            i = hdr.index(m)
            print(f"{m:92s}" + "".join(f"{d[i][:44]:>46s}" for d in data) + f"  {units[i]}")
        print()


def klass(op):
    op = op.split(".")[0]
    for k, names in (("FFMA2", ("FFMA2",)), ("FFMA/FMUL/FADD (fma pipe)", ("FFMA", "FMUL", "FADD")), ("FMNMX / FSETP / FSEL (alu)", ("FMNMX", "FMNMX3", "FSETP", "FSEL", "FCHK")),
                     ("int / logic / select (alu)", ("ISETP", "SEL", "LOP3", "PLOP3", "IADD3", "VIADD", "IADD", "SHF", "LEA", "PRMT", "IMNMX", "VIMNMX", "VIMNMX3", "POPC", "FLO", "BREV", "MOV", "CS2R", "R2P", "P2R", "I2F", "F2I", "I2FP", "F2FP")),
                     ("IMAD (fma pipe)", ("IMAD",)), ("FP64 + F2F", ("DFMA", "DADD", "DMUL", "DSETP", "F2F")), ("MUFU (xu)", ("MUFU",)),
                     ("LDS / STS", ("LDS", "STS", "LDSM")), ("LDG / STG / LDC / atomics", ("LDG", "STG", "LDC", "LDCU", "LD", "ST", "ATOM", "ATOMG", "RED", "ATOMS", "LDL", "STL")),
                     ("branch / sync / vote", ("BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "VOTE", "VOTEU", "SHFL", "MATCH", "CALL", "RET", "BRX", "NOP", "YIELD", "DEPBAR", "S2R", "S2UR", "ULEA", "UMOV", "UIADD3", "UISETP", "ULOP3", "USEL", "UIMAD", "ULDC", "BMOV", "ERRBAR", "MEMBAR", "CCTL", "ENDCOLLECTIVE", "ELECT"))):
        if op in names:
            return k
    return "other (" + op + ")"


def source(path, top=0):
    """Per kernel of the capture: warp-instructions by opcode class with their average active lanes (SASS view of the
    source page: rows whose Address is a code address)."""
    rows = list(csv.reader(open(path)))
    kernel, cols = None, None
    stats = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            kernel = r[1][r[1].find("k_"):].split("(")[0] if "k_" in r[1] else r[1]
            stats.setdefault(kernel, dict(tot_w=0, tot_t=0, by=defaultdict(lambda: [0, 0]), lines=[], seen=set()))
            cols = None
            continue
        if r[0] == "Address":
            h = [c.strip() for c in r]
            cols = (h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"))
            continue
        if kernel is None or cols is None or not r[0].startswith("0x"):
            continue
        if r[0] in stats[kernel]["seen"]:          # the page repeats the table per view
            continue
        stats[kernel]["seen"].add(r[0])
        si, wi, ti = cols
        try:
            w, t = int(float(r[wi])), int(float(r[ti]))
        except (ValueError, IndexError):
            continue
        toks = r[si].replace("@!", "@").split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        st = stats[kernel]
        k = klass(op)
        st["by"][k][0] += w; st["by"][k][1] += t
        st["tot_w"] += w; st["tot_t"] += t
        st["lines"].append((w, t, r[si].strip()))
    for kernel, st in stats.items():
        tw, tt = st["tot_w"], st["tot_t"]
        if not tw:
            continue
        print(f"== {path} :: {kernel}: {tw} warp-instructions, {tt / tw:.2f} active lanes per instruction")
        for k, (w, t) in sorted(st["by"].items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:36s} {w:14d} warp-instr  {100.0 * w / tw:5.1f} %   {t / max(w, 1):5.1f} lanes")
        if top:
            print(f"  -- the {top} most executed instructions")
            for w, t, src in sorted(st["lines"], key=lambda x: -x[0])[:top]:
                print(f"     {w:12d}  {t / max(w, 1):5.1f} lanes  {src}")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2:])
    else:
        for x in sys.argv[2:]:
            source(x)
