# two ranks: the NCCL gather of the decoded frames, NUMA binding, the D2H ceiling
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 2 > gpurun_out/r2h_n2.json 2> gpurun_out/r2h_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 5 > gpurun_out/r2h_n2_ref.json 2> gpurun_out/r2h_n2_ref.err
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1; nproc >> gpurun_out/r2h_topo.txt
