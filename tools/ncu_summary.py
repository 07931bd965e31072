"""Turn an `ncu --set full` report into the committed evidence: a CSV of the columns the roofline argument uses and, for the
wide-GEMM capture, profiles/r2_traffic.json (DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum).

    python tools/ncu_summary.py gpurun_out/r2_prof_tc_v2.ncu-rep profiles/r2_ncu_full_tc_gemm_v2.csv [--traffic profiles/r2_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum.per_cycle_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
ALGO = {"fwd_L1": 86.37e6, "fwd_L2": 167.8e6, "dgrad_L2": 234.9e6, "wgrad_L2 (fused db)": 201.3e6, "wgrad_L1 (fused db)": 92.8e6}


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v) * scale


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    print("wrote", out, len(data), "launches")
    if "--traffic" in sys.argv:
        tj = sys.argv[sys.argv.index("--traffic") + 1]
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        names = list(ALGO)
        assert len(data) == len(names), (len(data), "launches; expected the five wide GEMMs")
        per = {n: to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for n, r in zip(names, data)}
        json.dump({"tc_gemm_kernel": {"per_launch_bytes": sum(per.values()) / len(per), "launches": per, "algorithmic_bytes": ALGO,
                                      "source": f"{out} (ncu --set full --clock-control none of tools/gemm_launches.py: the five wide GEMM launches of a C4 step on "
                                                "real operands, HEAD kernels; dram__bytes_read.sum + dram__bytes_write.sum, mean over the 5 launches)"}},
                  open(tj, "w"), indent=1)
        print("wrote", tj, {k: round(v / 1e6, 1) for k, v in per.items()})


if __name__ == "__main__":
    main()
