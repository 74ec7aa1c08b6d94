set -x
python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_q.json 2>gpurun_out/r2_q.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_env_launches_raw.csv python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step_rows -s 70 -c 1 -o gpurun_out/r2_env_rows_32k_stagger -f python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_ncu2.log 2>&1
PPO_B=24000 PPO_ITERS=3 python profiles/prof_minibatch.py && \
PPO_B=24000 PPO_ITERS=3 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|wgrad_persistent" -s 7 -c 7 -o gpurun_out/r2_chain_24000_final -f python profiles/prof_minibatch.py > gpurun_out/r2_ncu3.log 2>&1
PPO_B=24000 PPO_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_mb24000_launches_raw.csv python profiles/prof_minibatch.py > gpurun_out/r2_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
