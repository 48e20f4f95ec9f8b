#!/bin/bash
# ncu on rank 0 of a 2-rank run (rank 1 plain): NVLink bytes + duration of the march (push) and of the slab optimiser kernel
# usage: tools/profile_mg.sh <workload> <out-prefix>
W=${1:-c2}; OUT=${2:-gpurun_out/r2_mg_$W}
export WORLD_SIZE=2 MASTER_ADDR=127.0.0.1 MASTER_PORT=29533
RANK=1 LOCAL_RANK=1 python tools/mg_steps.py $W 10 push > ${OUT}_rank1.log 2>&1 &
P1=$!
RANK=0 LOCAL_RANK=0 timeout 240 ncu --clock-control none --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    -k regex:'k_render_train|k_adam_slab|k_peer_barrier' -s 12 -c 6 --csv --log-file ${OUT}_ncu.csv python tools/mg_steps.py $W 10 push > ${OUT}_rank0.log 2>&1
wait $P1
tail -2 ${OUT}_rank0.log ${OUT}_rank1.log
