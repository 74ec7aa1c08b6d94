timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_worker.py 2>&1 | tail -5
for side in 1 0; do
RL_PPO_SIDE_COMM=$side timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --only-ppo 2>gpurun_out/r2_ppo_n2_side$side.err | cut -c1-200
done
