set -x
python tools/time_adapters.py > gpurun_out/r2_time_adapters.jsonl 2> gpurun_out/r2_time_adapters.err; echo rc=$?
export ADAPTER_REPS=1 ADAPTER_SIZES=512x16
python tools/time_adapters.py > gpurun_out/r2_plain_adapters.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_adapters.csv python tools/time_adapters.py > gpurun_out/r2_ncu_adapters.log 2>&1
python tools/time_adapters.py > gpurun_out/r2_plain_adapters.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lk_dense -c 1 -o gpurun_out/r2_prof_lk python tools/time_adapters.py > gpurun_out/r2_ncu_lk.log 2>&1
cat gpurun_out/r2_time_adapters.jsonl; tail -n 2 gpurun_out/r2_ncu_adapters.log; tail -n 2 gpurun_out/r2_ncu_lk.log
