set -x
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 300 python bench.py --config c4 --steps 3 --warmup 3 > gpurun_out/r2h_c4_n1.json 2> gpurun_out/r2h_c4_n1.err
timeout -s KILL 400 $T --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --config c4 --steps 3 --warmup 3 > gpurun_out/r2h_c4_n2.json 2> gpurun_out/r2h_c4_n2.err
timeout -s KILL 400 $T --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --config c4 --steps 3 --warmup 3 > gpurun_out/r2h_c4_n4.json 2> gpurun_out/r2h_c4_n4.err
timeout -s KILL 400 $T --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --config c4 --steps 3 --warmup 3 > gpurun_out/r2h_c4_n8.json 2> gpurun_out/r2h_c4_n8.err
timeout -s KILL 600 $T --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --config c5 --steps 2 --warmup 3 --no-extras --no-roofline --no-cpu-baseline > gpurun_out/r2h_c5_n8.json 2> gpurun_out/r2h_c5_n8.err
echo done
