#!/bin/bash
# development aid (run under gpurun): time + parity (tools/ab.py) and DRAM bytes of one list-kernel launch per library variant
#   tools/ab_traffic.sh base build/variants/libpolargpu_X.so ...
for lib in "$@"; do
  python tools/ab.py $lib
  if [ "$lib" != base ]; then export POLARGPU_LIB=$PWD/$lib; else unset POLARGPU_LIB; fi
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio \
      --clock-control none -k regex:^list_decode_kernel -s 2 -c 1 python tools/occ_probe.py 2>&1 | grep -E "dram__|lts__|gpu__time|inst_executed|long_scoreboard|Mframes"
done
