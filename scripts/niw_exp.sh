#!/bin/bash
mkdir -p gpurun_out
bash scripts/exp.sh $1_noload C4 MSB_NIW_DIAG=1
bash scripts/exp.sh $1_noload_noA C4 MSB_NIW_DIAG=5
bash scripts/exp.sh $1_noload_1prod C4 MSB_NIW_DIAG=9
bash scripts/exp.sh $1_noload_noA_1prod C4 MSB_NIW_DIAG=13
