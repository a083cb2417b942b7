# Round-2 multi-GPU evidence (run under `gpurun --gpus 8`): green multi-GPU pytest, H2D ceiling at 1/2/4/8 concurrent ranks,
# the 1e9-frame CA-SCL floor run on 8 GPUs through the drop-in program (pipelined pg_simulate, one all-reduce), 8-GPU bench line.
set -x
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/multi_pytest.log 2>&1; echo multi rc=$?; tail -3 gpurun_out/multi_pytest.log
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 tools/h2d_scaling.py 2>/dev/null | grep '^{' >> gpurun_out/h2d_scaling.jsonl
done
cat gpurun_out/h2d_scaling.jsonl
( time polardecoding_b200/host/bin/CASCL_1024_L8 --seed 1242 --gpus 8 --ebn0 3.0 --max-frames 1000000000 --verbose ) > gpurun_out/floor_1e9_8gpu.txt 2>&1; cat gpurun_out/floor_1e9_8gpu.txt
( time polardecoding_b200/host/bin/CASCL_1024_L8 --seed 1242 --gpus 8 --ebn0 2.5 --ble 20000 --verbose ) > gpurun_out/exact_stop_8gpu.txt 2>&1; cat gpurun_out/exact_stop_8gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err; echo bench8 rc=$?
tail -c 600 gpurun_out/bench_r2_8gpu.json
