"""Turns `ncu -i X.ncu-rep --page raw --csv` (wide: one column per metric) into the tall `metric,unit,value` summary that
is committed under profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).

    python scripts/ncu_summary.py gpurun_out/r02_wcontract.ncu-rep profiles/r02_ncu_full_wcontract.csv
"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(dram__bytes|gpu__dram_throughput|gpu__time_duration|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|"
                  r"l1tex__data_pipe_lsu_wavefronts|l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|st|red)|launch__|"
                  r"lts__t_sector_hit_rate|lts__t_sectors_op_(read|write)\.sum|sm__inst_executed_pipe_(fp64|tensor|lsu|uniform)|"
                  r"sm__ops_path_tensor_src_fp64|sm__pipe_tensor|sm__throughput|sm__warps_active|sm__cycles_(active|elapsed)|"
                  r"smsp__average_warps?_issue_stalled|smsp__warp_issue_stalled|smsp__inst_executed\.sum|smsp__issue_active|"
                  r"sm__sass_inst_executed_op_(shared|global)|smsp__cycles_active\.avg)")

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write("metric,unit,value\n")
    for r in rows[2:]:
        f.write(f"# kernel,,{r[hdr.index('Kernel Name')]}\n")
        for i, h in enumerate(hdr):
            name = h.split(".", 2)[-1] if h.split(".")[0].isupper() or "Triage" in h else h
            if KEEP.match(h) or KEEP.match(name):
                f.write(f"{h},{units[i]},{r[i]}\n")
print(out, sum(1 for _ in open(out)))
