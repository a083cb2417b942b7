# one full-set capture of the CA-SCL list kernel (source level), driven by tools/ab.py's timing loop (development aid)
set -x
python tools/ab.py base > gpurun_out/ab_base.log 2>&1; echo ab rc=$?
ncu --set full --clock-control none --import-source on -k regex:^list_decode_kernel -s 6 -c 1 -f -o gpurun_out/prof_list python tools/occ_probe.py > gpurun_out/ncu_list.log 2>&1; echo ncu rc=$?
