# Round-end evidence on one GPU (run under gpurun): the bench line, the ncu launch list of the same command, one full-set capture per
# kernel (source level).  The .ncu-rep files land in gpurun_out/; `python tools/summarize_profiles.py r2` copies their summaries into profiles/.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo bench rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_b.log 2>&1; echo launchlist rc=$?
cap() {  # name, kernel regex (demangled name), legs
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 3 -c 1 -f -o gpurun_out/prof_$1_r2 python bench.py --steps 2 --warmup 3 --no-cpu --legs $3 > gpurun_out/ncu_full_$1.log 2>&1; echo $1 rc=$?
}
cap cascl 'list_decode_kernel<float' cascl
cap cascl64 'list_decode_kernel<double' cascl64
cap bp 'bp_decode_kernel<float' bp
cap bp64 'bp_decode_kernel<double' bp64
ls -la gpurun_out/*.ncu-rep
