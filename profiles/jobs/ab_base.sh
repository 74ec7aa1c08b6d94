# same-box A/B of the working tree's library against rapid_locomotion_rl_b200/librl_b200_base.so (env step sizes)
cp rapid_locomotion_rl_b200/librl_b200.so /tmp/cur.so
python profiles/jobs/env_ab.py NEW 2>&1 | grep envs
cp rapid_locomotion_rl_b200/librl_b200_base.so rapid_locomotion_rl_b200/librl_b200.so
python profiles/jobs/env_ab.py BASE 2>&1 | grep envs
cp /tmp/cur.so rapid_locomotion_rl_b200/librl_b200.so
