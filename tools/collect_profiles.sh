# Collects the round's measurement artefacts on the GPU box into gpurun_out/ (run under gpurun):
#   bench line (N=1) and reference arm, per-config table, single-call latencies, sharp/riskiness times,
#   ncu launch list of the bench command, ncu --set full captures of the main kernels.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
python tools/configs.py > gpurun_out/configs.txt 2>&1
python tools/latency.py > gpurun_out/latency.txt 2>&1
python tools/sharp_bench.py > gpurun_out/sharp.txt 2>&1
python tools/sharp_bench.py 1920 1080 >> gpurun_out/sharp.txt 2>&1
python tools/bench_config5.py --reps 3 > gpurun_out/config5_n1.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'f1_fast|entropy_pack|stuff_kernel' -s 30 -c 3 \
    -o gpurun_out/pipeline_full -f python tools/run_f1.py 8 3 full > gpurun_out/pipeline_full.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'f1_fast|entropy_pack|stuff_kernel' -s 30 -c 3 --csv --log-file gpurun_out/pipeline_light.csv \
    python tools/run_f1.py 8 3 full > gpurun_out/pipeline_light.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'entropy_pack|stuff_kernel' -s 20 -c 2 \
    -o gpurun_out/entropy_warm -f python tools/run_f1.py 4 3 full > gpurun_out/entropy_warm.log 2>&1
ncu --set full --clock-control none -k regex:'sharp_|riskiness' -c 5 -o gpurun_out/sharp_full -f \
    python tools/sharp_bench.py 1920 1080 > gpurun_out/sharp_full.log 2>&1
ls -la gpurun_out
