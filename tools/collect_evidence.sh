#!/bin/bash
# Round evidence, one GPU: tests, bench (both arms), ncu full capture of the hot kernels, ncu launch list of the bench command,
# issue microbenchmark, per-kernel timings, parity report.  Everything lands in gpurun_out/ev/ (copy what is kept into profiles/).
cd "$(dirname "$0")/.."
O=gpurun_out/ev; mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.err
# full capture of the hot kernels (after the plain run above exited 0)
ncu --set full --import-source on --clock-control none -k regex:"chain_rows_kernel|decay2body_kernel|scan_kernel" --launch-skip 5 --launch-count 5 \
    -o $O/prof_r2_final -f python tools/ncu_target.py > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
# launch list of the bench command (device-resident steps + e2e; extra legs off so the list stays short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --legs none > $O/ncu_launch.log 2>&1; tail -c 200 $O/ncu_launch.log
bash tools/run_issue_microbench.sh > /dev/null 2>&1; cp gpurun_out/issue_microbench.txt $O/
python tools/parity_report.py > $O/parity.log 2>&1; cp gpurun_out/parity_r2.json $O/ 2>/dev/null; tail -3 $O/parity.log
python tools/profile_steps.py > $O/kernel_timings.json 2> $O/kernel_timings.err
{ for t in time_chain time_scan time_meson time_brem time_resonance time_misc time_chain_hist time_scatter time_volint time_general_decay time_c4; do
    echo "# tools/$t.py"; python tools/$t.py 2>&1 | tail -8; done; } > $O/kernel_timings_tools.txt
python tools/profile_scan.py 2>/dev/null | grep "^{" > $O/scan_profile_n1.json
python tools/profile_e2e_host.py > $O/e2e_host_profile.txt 2>&1
ls -la $O
