#!/bin/bash
# tables-only kernel check: the parity tests that sweep table states, then C2 with and without the persistent kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "sweep or bit_exact or baseline_shapes or tables or dd or bb" > gpurun_out/$1_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/$1_tests.log)"
bash scripts/exp.sh $1_C2 C2
bash scripts/exp.sh $1_C2_old C2 MSB_NO_PERSISTENT=1
