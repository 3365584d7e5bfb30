# One-shot A/B on the GPU box: GPU test suite on the default build, then the 16-frame device pipeline
# (gen B and gen A) for each library under build/ab/ and for the default build.
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
{
for v in default $(ls build/ab/*.so 2>/dev/null); do
  if [ "$v" = default ]; then unset SJPEG_B200_LIB; else export SJPEG_B200_LIB=$PWD/$v; fi
  for g in B A; do
    echo "== $v gen$g"
    timeout 60 python tools/run_f1.py 16 10 full 3840 2160 1 0 $g 2>&1 | tail -2
  done
done
} > gpurun_out/ab_layout.txt 2>&1
cat gpurun_out/ab_layout.txt
