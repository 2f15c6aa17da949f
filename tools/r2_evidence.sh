#!/bin/bash
# profiler evidence of the code as shipped: ncu --set full of one default training step (eager, side streams off) and of one
# chest-sized resample call, and the ncu launch list of a short bench.py run
mkdir -p gpurun_out
T=r2e
python tools/prof_step.py 3 > gpurun_out/${T}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${T}_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'adam|attn|bn_|conv|dropout|gemm|gru|head|pool|transpose|wgrad' -s 54 -c 27 \
    -o gpurun_out/${T}_step -f python tools/prof_step.py 3 > gpurun_out/${T}_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -s 20 -c 20 -o gpurun_out/${T}_resample -f python tools/resample_probe.py > gpurun_out/${T}_ncu_resample.log 2>&1; echo "ncu resample rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-subrecords --no-cpu-baseline --no-library-baseline > gpurun_out/${T}_bench_short_plain.json 2>/dev/null; echo "plain short bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launch_list_bench.csv python bench.py --steps 2 --warmup 1 --no-subrecords --no-cpu-baseline --no-library-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python tools/ncu_summary.py gpurun_out/${T}_ncu_full_summary.json gpurun_out/${T}_step.ncu-rep gpurun_out/${T}_resample.ncu-rep; echo "summary rc=$?"
ncu -i gpurun_out/${T}_resample.ncu-rep --page raw --csv > gpurun_out/${T}_resample_raw.csv 2>/dev/null
rm -f gpurun_out/${T}_resample.ncu-rep          # gpurun_out/ travels back only below 64 MiB: the step report (47 MB) and the summaries do
ls -la gpurun_out/${T}_*
