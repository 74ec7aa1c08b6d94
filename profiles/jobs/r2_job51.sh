set -x
cp rapid_locomotion_rl_b200/librl_b200.so /tmp/cur.so
python profiles/jobs/env_ab.py HEAD | head -3
cp rapid_locomotion_rl_b200/librl_b200_gate.so rapid_locomotion_rl_b200/librl_b200.so
python profiles/jobs/env_ab.py GATE | head -3
cp /tmp/cur.so rapid_locomotion_rl_b200/librl_b200.so
python profiles/jobs/env_ab.py HEAD2 | head -3
