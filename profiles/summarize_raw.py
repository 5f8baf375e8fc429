"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one `--set full` capture per kernel) into the
handful of numbers the roofline discussion needs: duration, DRAM traffic, occupancy, cache hit
rates, issue utilisation and the top warp-stall reasons.

    python profiles/summarize_raw.py gpurun_out/prof_raw.csv > profiles/kernels_rNN.md
"""
import csv
import io
import re
import sys

KERNELS = r"(BellSpmvKernel<[^>]*>|BellJacobiKernel<[^>]*>|emi_assemble_kernel|knp_assemble_kernel|EmiPrepassKernel|GradKernel|coarse_tail_kernel|p2p_halo_kernel)"


def main(path):
    text = open(path).read()
    rows = list(csv.reader(io.StringIO(text[text.find('"ID"'):])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name, default=float("nan")):
        if name not in idx:
            return default
        try:
            v = float(r[idx[name]].replace(",", ""))
        except ValueError:
            return default
        u = units[idx[name]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3,
                 "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
        return v * scale

    stalls = [h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h)]
    print("| kernel (launch) | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | regs | occ % | L1 hit % | L2 hit % | issue % | "
          "fp64 pipe % | top stalls (warps per issue) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for r in data:
        name = r[idx["Kernel Name"]]
        m = re.search(KERNELS, name)
        short = m.group(1) if m else name[:40]
        us = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        st = sorted(((val(r, s, 0.0), s) for s in stalls), reverse=True)[:3]
        sts = ", ".join(f"{re.sub('smsp__average_warps_issue_stalled_|_per_issue_active.ratio', '', s)} {v:.1f}" for v, s in st)
        print(f"| {short} (#{r[idx['ID']]}) | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / us / 1e3:.0f} | "
              f"{val(r, 'launch__registers_per_thread'):.0f} | {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{val(r, 'l1tex__t_sector_hit_rate.pct'):.0f} | {val(r, 'lts__t_sector_hit_rate.pct'):.0f} | "
              f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.0f} | {sts} |")


if __name__ == "__main__":
    main(sys.argv[1])
