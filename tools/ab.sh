# A/B helper: timings of the device pipeline for the library in SJPEG_B200_LIB (default build if unset)
for g in B A; do timeout 120 python tools/run_f1.py 4 20 full 3840 2160 1 0 $g 2>&1 | tail -1; done
timeout 120 python tools/run_f1.py 16 10 full 2>&1 | tail -2 | head -1
timeout 120 python tools/run_f1.py 16 10 full 3840 2160 1 0 A 2>&1 | tail -2 | head -1
