set -x
cp rapid_locomotion_rl_b200/librl_b200.so /tmp/cur.so
python profiles/jobs/env_ab.py CUR | head -3
for v in PF SP; do
cp rapid_locomotion_rl_b200/librl_b200_$v.so rapid_locomotion_rl_b200/librl_b200.so
python profiles/jobs/env_ab.py $v | head -3
done
cp /tmp/cur.so rapid_locomotion_rl_b200/librl_b200.so
python profiles/jobs/env_ab.py CUR2 | head -3
