#!/bin/bash
# Round-2 measurement pass on ONE B200 (run under gpurun): FP64 peak record, ncu launch list of bench.py, ncu full captures
# of the three kernels (W contraction, its plain NT-GEMM mode, fused energy), bench lines of the other BASELINE shapes.
# Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
python scripts/fp64_peaks.py > $O/r02_fp64_peaks.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --job-units 296 --no-e2e --no-cpu-baseline --oracle-units 0 > $O/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:w_contract -s 2 -c 1 -f -o $O/r02_wcontract \
    python scripts/profile_target.py 63 297 8 2 dense > $O/r02_ncu_w.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:t_energy_fused -s 1 -c 1 -f -o $O/r02_tenergy \
    python scripts/profile_target.py 63 297 8 2 dense > $O/r02_ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:w_contract -s 2 -c 1 -f -o $O/r02_plain_gemm \
    python scripts/profile_target.py 63 297 2 2 df > $O/r02_ncu_g.log 2>&1
python scripts/profile_target.py 63 297 592 0 df > $O/r02_df_resident.log 2>&1
python scripts/profile_target.py 63 297 592 0 dfpanel > $O/r02_df_panel.log 2>&1
python bench.py --workload benzene-cc-pVDZ --steps 5 --warmup 3 > $O/r02_bench_benzene.log 2>&1
python bench.py --workload uracil-dimer-6-31Gs --steps 3 --warmup 2 > $O/r02_bench_dimer.log 2>&1
python bench.py --workload water10-cc-pVTZ --steps 2 --warmup 1 --job-units 148 --no-e2e --no-cpu-baseline --oracle-units 1 > $O/r02_bench_water10.log 2>&1
python bench.py --workload synthetic-o50-v500 --steps 2 --warmup 1 --job-units 148 --no-e2e --no-cpu-baseline --oracle-units 1 > $O/r02_bench_o50v500.log 2>&1
tail -c 600 $O/r02_fp64_peaks.log $O/r02_df_resident.log $O/r02_df_panel.log
for f in benzene dimer water10 o50v500; do tail -1 $O/r02_bench_$f.log | cut -c1-400; done
