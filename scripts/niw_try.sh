#!/bin/bash
# NIW kernel check: the NIW parity tests, then the C4 bench record
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "niw or tensor or C4" > gpurun_out/$1_niwtests.log 2>&1; echo "niw tests rc=$? $(tail -1 gpurun_out/$1_niwtests.log)"
bash scripts/exp.sh $1_C4 C4
