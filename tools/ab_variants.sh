# Kernel A/B on a GPU box.  Build the variants HERE first (no GPU needed):
#     bash tools/ab_variants.sh build name1:"-DFLAG=1" name2:"-DOTHER=2 -DMORE=3" ...
# writes build/ab/<name>.so (kernels.cu and engine.cu recompiled with the flags, the other objects
# reused), then run under gpurun:
#     gpurun --timeout 300 -- 'bash tools/ab_variants.sh run [test]'
# which times the 16-frame 4K device pipeline (gen B and gen A) for the default build and for every
# build/ab/*.so, and with `test` also runs the GPU test suite against each variant.
# Environment for a variant can be given in build/ab/<name>.env (sourced before its runs).
set -e
cd "$(dirname "$0")/.."
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wall --fmad=false --expt-relaxed-constexpr"
if [ "$1" = build ]; then
  shift
  make -C sjpeg_b200/csrc > /dev/null
  mkdir -p build/ab
  for spec in "$@"; do
    name=${spec%%:*}; defs=${spec#*:}
    $NVCC $FLAGS $defs -c sjpeg_b200/csrc/kernels.cu -o build/ab/$name.kernels.o
    $NVCC $FLAGS $defs -c sjpeg_b200/csrc/engine.cu -o build/ab/$name.engine.o
    (cd sjpeg_b200/csrc && $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/ab/$name.so \
        ../../build/ab/$name.kernels.o sharp.o ../../build/ab/$name.engine.o host_codec.o host_stager.o sjpeg_api.o -lcudart -lpthread -ldl)
    rm -f build/ab/$name.kernels.o build/ab/$name.engine.o
    echo "built build/ab/$name.so ($defs)"
  done
  exit 0
fi
mkdir -p gpurun_out
{
for v in default $(ls build/ab/*.so 2>/dev/null); do
  (
  if [ "$v" != default ]; then
    export SJPEG_B200_LIB=$PWD/$v
    [ -f "${v%.so}.env" ] && . "${v%.so}.env"
  fi
  for g in B A; do
    echo "== $v gen$g"
    timeout 60 python tools/run_f1.py 16 10 full 3840 2160 1 0 $g 2>&1 | tail -2
  done
  if [ "$2" = test ] && [ "$v" != default ]; then
    echo "== $v GPU tests"
    timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  fi
  )
done
} 2>&1 | tee gpurun_out/ab_variants.txt
