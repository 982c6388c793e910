mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nich or mixed or baseline_shapes or score_rows or far_from or mask" > gpurun_out/$1_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/$1_tests.log)"
bash scripts/exp.sh $1_C5 C5
bash scripts/exp.sh $1_C3 C3
