# Final measurements of the round on the GPU box (ordered by importance; budget is short).
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
timeout 60 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'f1_fast|entropy_pack|stuff_kernel' -s 30 -c 3 --csv --log-file gpurun_out/pipeline_light.csv \
    python tools/run_f1.py 8 3 full > gpurun_out/pipeline_light.log 2>&1
{
export SJPEG_B200_LIB=$PWD/build/ab/lw8.so
for g in B A; do echo "== lw8 gen$g"; timeout 60 python tools/run_f1.py 16 10 full 3840 2160 1 0 $g 2>&1 | tail -2; done
export SJPEG_B200_LIB=$PWD/build/ab/g16.so SJB_GROUP_BUDGET_MB=400
echo "== g16 genB"; timeout 60 python tools/run_f1.py 16 10 full 3840 2160 1 0 B 2>&1 | tail -2
} > gpurun_out/ab_variants.txt 2>&1
cat gpurun_out/ab_variants.txt
