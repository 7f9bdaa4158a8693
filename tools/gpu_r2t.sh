# two ranks under torchrun at the end of the round: our arm (NCCL gather in the timed region) and the reference arm
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-also > gpurun_out/r2t_n2.json 2> gpurun_out/r2t_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 5 > gpurun_out/r2t_n2_ref.json 2> gpurun_out/r2t_n2_ref.err
