for v in libsjpeg_b200 lib_B lib_C lib_D; do
  echo "== $v"
  export SJPEG_B200_LIB=$PWD/sjpeg_b200/$v.so
  for g in B A; do python tools/run_f1.py 4 20 full 3840 2160 1 0 $g 2>&1 | tail -1; done
  python tools/run_f1.py 16 10 full 2>&1 | tail -2 | head -1
  python tools/run_f1.py 16 10 full 3840 2160 1 0 A 2>&1 | tail -2 | head -1
done
unset SJPEG_B200_LIB
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
