set -x
nvidia-smi -L
python -m pytest tests/test_multigpu.py -q -x 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --only-ppo 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --only-ppo --ppo-envs 32768 2>&1 | tail -3
python bench.py --only-ppo --ppo-envs 32768 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
