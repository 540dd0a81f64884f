#!/usr/bin/env python
"""Text summary of an ncu report (one kernel launch): the metrics the roofline discussion uses, the stall
reasons, the opcode mix and the hottest source lines.  Usage: tools/ncu_summary.py <file.ncu-rep> <units>
where <units> is the number of work units of the launch (particles), for the per-unit figures."""
import collections
import csv
import io
import re
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    h, u, v = rows[0], rows[1], rows[2]
    m = {n: (u[i], v[i]) for i, n in enumerate(h)}
    g = lambda k: float(m[k][1].replace(",", "")) if k in m and m[k][1] not in ("", "n/a") else float("nan")
    print(f"report: {rep}")
    print(f"kernel: {m.get('Kernel Name', ('', '?'))[1]}   grid {m.get('Grid Size', ('', '?'))[1]} block {m.get('Block Size', ('', '?'))[1]}")
    t = g("gpu__time_duration.sum")
    tu = m["gpu__time_duration.sum"][0]
    t_ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1.0)
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= sc.get(m["dram__bytes_read.sum"][0], 1)
    wr *= sc.get(m["dram__bytes_write.sum"][0], 1)
    print(f"duration (under ncu, cold caches, serialised): {t_ms:.3f} ms for {units:.4g} units")
    print(f"DRAM read {rd/1e9:.3f} GB + write {wr/1e9:.3f} GB = {(rd+wr)/units:.1f} B/unit; achieved {(rd+wr)/t_ms/1e6:.0f} GB/s")
    for k in ("launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
              "smsp__inst_executed_op_shared_atom.sum", "lts__t_sectors_srcunit_tex_op_atom.sum",
              "lts__t_sectors_srcunit_tex_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
              "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"):
        if k in m:
            print(f"  {k:75s} {m[k][1]:>18s} {m[k][0]}")
    ie = g("smsp__inst_executed.sum")
    print(f"  warp instructions per 32 units: {ie / (units / 32):.1f}")
    sa = g("smsp__inst_executed_op_shared_atom.sum")
    if sa == sa:
        print(f"  shared atomics: {sa/1e6:.2f} M warp-instr = {sa/t_ms/1e6:.2f} G/s; global atomic+red sectors: "
              f"{(g('lts__t_sectors_srcunit_tex_op_atom.sum')+g('lts__t_sectors_srcunit_tex_op_red.sum'))/t_ms/1e6:.2f} G/s")
    print("stall reasons (warps per issue):")
    st = [(float(v[i]), n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled") and
          n.endswith("_per_issue_active.ratio") and v[i] not in ("", "n/a")]
    for val, n in sorted(st, reverse=True)[:9]:
        print(f"  {n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {val:.3f}")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    if len(src) < 4:
        return
    hdr = src[2]
    iI, iS = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ops, per, cur = collections.Counter(), collections.OrderedDict(), None
    for r in src[3:]:
        if not r:
            continue
        if r[0].strip().isdigit():
            cur = (int(r[0]), r[1].strip())
            a = per.setdefault(cur, [0, 0])
            a[0] += int(r[iS]) if r[iS] not in ("-", "") else 0
            a[1] += int(r[iI]) if r[iI] not in ("-", "") else 0
        elif r[0] == "" and len(r) > iI and r[2] not in ("", "...", "-"):
            mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
            if mm and r[iI] not in ("-", ""):
                ops[mm.group(2).split(".")[0]] += int(r[iI])
    tot = sum(ops.values())
    rounds = units / 32
    print("opcode mix (warp instructions per 32 units):")
    print("  " + ", ".join(f"{k} {c/rounds:.1f}" for k, c in ops.most_common(16)))
    fp64 = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"  FP64 {fp64/rounds:.1f}  LDS {ops['LDS']/rounds:.1f}  ATOMS {ops['ATOMS']/rounds:.1f}  total {tot/rounds:.1f}")
    ts = sum(a[0] for a in per.values()) or 1
    print("hottest source lines (share of stall samples, instructions per 32 units):")
    for (ln, s_), a in sorted(per.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"  {ln:5d} {100*a[0]/ts:5.1f} % {a[1]/rounds:7.1f}  {s_[:96]}")


if __name__ == "__main__":
    main()
