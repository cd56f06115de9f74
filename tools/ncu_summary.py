"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx_summary.csv ["header comment"]
"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|dram__throughput\.avg\.pct|"
    r"gpu__dram_throughput\.avg\.pct|dram__cycles_active\.avg\.pct|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct|"
    r"l1tex__m_xbar2l1tex_read_bytes\.sum$|l1tex__m_l1tex2xbar_write_bytes\.sum$|"
    r"sm__pipe_tensor_cycles_active\.avg\.pct|sm__inst_executed_pipe_tensor.*\.avg\.pct|sm__throughput\.avg\.pct|"
    r"sm__warps_active\.avg\.pct|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct|"
    r"launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic)$|"
    r"sm__cycles_elapsed\.avg(\.per_second)?$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|"
    r"smsp__average_warps_issue_stalled_(long_scoreboard|barrier|wait|short_scoreboard)_per_issue_active\.ratio)")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    comment = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        if comment:
            f.write(f"# {comment}\n")
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(launches))) + "\n")
        for i, name in enumerate(hdr):
            if KEEP.search(name):
                vals = [l[i].replace(",", "") if i < len(l) else "" for l in launches]
                f.write(f"{name},{units[i]}," + ",".join(v if "(" not in v else '"' + v + '"' for v in vals) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
