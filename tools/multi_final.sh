# Round-2 final multi-GPU record (run under `gpurun --gpus 8`): the 1e9-frame floor run and an exact-stop run through the drop-in
# program on 8 GPUs (pipelined pg_simulate), with the same program on 1 GPU beside them for the per-GPU sustained rate.
set -x
polardecoding_b200/host/bin/CASCL_1024_L8 --seed 1242 --gpus 1 --ebn0 3.0 --max-frames 150000000 --verbose 2>&1 | tail -2 > gpurun_out/floor_1gpu.txt; cat gpurun_out/floor_1gpu.txt
polardecoding_b200/host/bin/CASCL_1024_L8 --seed 1242 --gpus 8 --ebn0 3.0 --max-frames 1000000000 --verbose > gpurun_out/floor_1e9_8gpu.txt 2>&1; cat gpurun_out/floor_1e9_8gpu.txt
polardecoding_b200/host/bin/CASCL_1024_L8 --seed 1242 --gpus 8 --ebn0 2.5 --ble 40000 --verbose > gpurun_out/exact_stop_8gpu.txt 2>&1; cat gpurun_out/exact_stop_8gpu.txt
polardecoding_b200/host/bin/BP_1024 --seed 555 --gpus 8 --ebn0 2.5 --max-frames 8000000 --verbose > gpurun_out/bp_8gpu.txt 2>&1; cat gpurun_out/bp_8gpu.txt
