for envs in 4000 32768; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --only-ppo --ppo-envs $envs 2>gpurun_out/r2_ppo_n8_$envs.err | tee gpurun_out/r2_ppo_n8_$envs.json | cut -c1-330
done
timeout 300 python bench.py --only-ppo 2>/dev/null | cut -c1-200
