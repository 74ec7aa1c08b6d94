set -x
python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -12
python profiles/prof_rollout.py > gpurun_out/r2_rollout2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_rollout2_launches_raw.csv python profiles/prof_rollout.py > gpurun_out/r2_rollout2_ncu.log 2>&1
tail -1 gpurun_out/r2_rollout2_plain.log
PPO_B=24000 PPO_ITERS=3 python profiles/prof_minibatch.py && \
PPO_B=24000 PPO_ITERS=3 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|wgrad_persistent" -s 7 -c 7 -o gpurun_out/r2_chain_24000 -f python profiles/prof_minibatch.py > gpurun_out/r2_ncu_chain.log 2>&1
ls -la gpurun_out | tail -5
