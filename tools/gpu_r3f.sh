# two ranks under torchrun: our arm (NCCL gather in the timed region, streaming end-to-end loop) and the reference arm
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-also > gpurun_out/r3f_n2.json 2> gpurun_out/r3f_n2.err
tail -c 600 gpurun_out/r3f_n2.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r3f_n2.json").read().strip().splitlines()[-1])
print("N=2: value %.0f e2e %.0f ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
PY
