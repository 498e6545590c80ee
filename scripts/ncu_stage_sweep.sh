#!/bin/bash
# One ncu pass over every stage of scripts/stage_bench.py: per-kernel duration, DRAM bytes and the
# bytes that crossed the L1<->crossbar ports (a ratio above 1 against the DRAM bytes means half-filled
# sectors or re-fetched lines).  Usage (on the GPU box, after the plain command has exited 0):
#   scripts/ncu_stage_sweep.sh <scale> <out.csv>
set -e
SCALE=${1:-4}; OUT=${2:-gpurun_out/stage_sweep.csv}
python scripts/stage_bench.py --scale $SCALE --reps 2 > ${OUT%.csv}.plain.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_l1tex2xbar_write_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
    --clock-control none -k regex:"^k_|^void k_" -c 600 --csv --log-file $OUT \
    python scripts/stage_bench.py --scale $SCALE --reps 2 > ${OUT%.csv}.ncu.log 2>&1
