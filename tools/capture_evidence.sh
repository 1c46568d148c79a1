set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo bench rc=$?
BCMD="python bench.py --pairs-per-gpu 64 --steps 1 --warmup 3 --no-cpu --no-banded --no-configs --no-adapters --e2e-pairs 64 --e2e-steps 1"
$BCMD > gpurun_out/r2_plain_b64.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1292 -c 323 --csv --log-file gpurun_out/r2_launches_bench64.csv $BCMD > gpurun_out/r2_ncu_launches.log 2>&1
export SWEEP=one HS_VARIANT=24 LS_VARIANT=8 LS_FUSE=4 HS_FUSE=4
python tools/sweep_fuse.py > gpurun_out/r2_plain_one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hs_tma_kernel -s 360 -c 1 -o gpurun_out/r2_prof_hs_fast_fine python tools/sweep_fuse.py > gpurun_out/r2_ncu1.log 2>&1
python tools/sweep_fuse.py > gpurun_out/r2_plain_one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hs_tma_kernel -s 260 -c 1 -o gpurun_out/r2_prof_hs_fast_coarse python tools/sweep_fuse.py > gpurun_out/r2_ncu2.log 2>&1
python tools/sweep_fuse.py > gpurun_out/r2_plain_one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ls_tma_kernel -s 50 -c 1 -o gpurun_out/r2_prof_ls_fine python tools/sweep_fuse.py > gpurun_out/r2_ncu3.log 2>&1
export HS_PRECISE=2
python tools/sweep_fuse.py > gpurun_out/r2_plain_onep.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hs_tma_kernel -s 460 -c 1 -o gpurun_out/r2_prof_hs_precise_fine python tools/sweep_fuse.py > gpurun_out/r2_ncu4.log 2>&1
tail -2 gpurun_out/r2_ncu1.log gpurun_out/r2_ncu2.log gpurun_out/r2_ncu3.log gpurun_out/r2_ncu4.log gpurun_out/r2_ncu_launches.log
cat gpurun_out/r2_plain_one.log gpurun_out/r2_plain_onep.log | cut -c1-300
