# Collects the round's measurement artefacts on the GPU box into gpurun_out/r02/ (run under gpurun,
# one GPU).  Summaries are made from them HERE afterwards (tools/ncu_summary.py) and copied to profiles/.
set -x
O=gpurun_out/r02
mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_n1.json 2>> $O/bench_n1.err
python tools/latency.py > $O/latency.txt 2>&1
python tools/quick_perf.py > $O/quick_perf.txt 2>&1
python tools/batch_methods.py B 16 > $O/batch_methods.txt 2>&1
python tools/batch_methods.py A 16 >> $O/batch_methods.txt 2>&1
python tools/planar_bench.py > $O/planar.txt 2>&1
python tools/sharp_bench.py > $O/sharp.txt 2>&1
python tools/sharp_batch.py > $O/sharp_batch.txt 2>&1
python tools/pageable_batch.py > $O/pageable_batch.txt 2>&1
# launch list of the bench command (headline part)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file $O/launches_bench.csv \
    python bench.py --quick --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
# --set full captures: F1 4:2:0 / 4:4:4 / planar, E + S (gen B and gen A), H1 + Q1 + S1 (m4), T1 (8K gen A m7)
ncu --set full --clock-control none --import-source on -k regex:'f1_fast' -s 20 -c 1 -o $O/f1_420_full -f \
    python tools/run_f1.py 16 3 f1 > $O/f1_420.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'f1_fast' -s 20 -c 1 -o $O/f1_444_full -f \
    python tools/run_f1.py 16 3 f1 3840 2160 3 0 B > $O/f1_444.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'entropy_pack|stuff_kernel' -s 8 -c 2 -o $O/es_genB_full -f \
    python tools/run_f1.py 16 3 full > $O/es_genB.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'entropy_pack|stuff_kernel' -s 8 -c 2 -o $O/es_genA_full -f \
    python tools/run_f1.py 16 3 full 3840 2160 1 0 A > $O/es_genA.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'histogram_kernel|analyse_|requantize|symbol_stats' -s 5 -c 5 -o $O/m4_full -f \
    python tools/run_f1.py 16 2 full 3840 2160 1 4 B > $O/m4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'trellis' -s 3 -c 3 -o $O/trellis_full -f \
    python tools/one_encode.py A 7680 4320 75 7 1 2 > $O/trellis.log 2>&1
ncu --set full --clock-control none -k regex:'f1_planar' -s 4 -c 1 -o $O/f1_planar_full -f \
    python tools/planar_bench.py > $O/f1_planar.log 2>&1
# compute-sanitizer on a small mixed workload (parity checked in the same run)
for t in memcheck racecheck synccheck initcheck; do
  echo "== $t" >> $O/sanitizer.txt
  timeout 600 compute-sanitizer --tool $t python tools/sanitize_run.py 2>&1 | grep -E "parity|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -5 >> $O/sanitizer.txt
done
# gpurun brings back at most 64 MiB: keep the raw-metric CSV of every report (what tools/ncu_summary.py needs)
# and drop reports, least needed first, until the directory fits
for r in $O/*_full.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null; done
for r in f1_planar f1_444 f1_420 m4 trellis; do
  [ "$(du -sm gpurun_out | cut -f1)" -lt 56 ] && break
  rm -f $O/${r}_full.ncu-rep
done
ls -la $O
