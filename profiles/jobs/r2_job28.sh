for nw in 2 4; do
RL_CHAIN_WORKERS=$nw B=18944 python profiles/prof_teacher.py
RL_CHAIN_WORKERS=$nw B=18944 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 2 -c 1 -o gpurun_out/r2_teacher_w$nw -f python profiles/prof_teacher.py > gpurun_out/r2_teacher_w${nw}_ncu.log 2>&1
tail -2 gpurun_out/r2_teacher_w${nw}_ncu.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
