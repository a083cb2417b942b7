# Round-end evidence on one GPU: bench line, ncu launch list of the same command, one full-set capture per kernel.
# (run under gpurun; the .ncu-rep files land in gpurun_out/, their summaries are copied into profiles/ by tools/summarize_profiles.sh)
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo bench rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_b.log 2>&1; echo launchlist rc=$?
for k in list_decode_kernel:cascl bp_decode_kernel:bp bp_decode_h2_kernel:bph2; do
  ncu --set full --clock-control none --import-source on -k regex:^${k%%:*}\$ -s 3 -c 1 -f -o gpurun_out/prof_${k##*:}_r1_final2 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_${k##*:}_b.log 2>&1; echo ${k##*:} rc=$?
done
