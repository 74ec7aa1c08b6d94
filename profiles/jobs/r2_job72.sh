set -x
python profiles/jobs/rough_ncu.py && \
ncu --set full --clock-control none --import-source on -k regex:"env_heights_prepass|env_step_quad" -s 6 -c 2 -o gpurun_out/r2_rough_32k -f python profiles/jobs/rough_ncu.py > gpurun_out/r2_ncu72.log 2>&1
ncu -i gpurun_out/r2_rough_32k.ncu-rep --page raw --csv > gpurun_out/r2_rough_32k_raw.csv 2>/dev/null
ls -la gpurun_out/r2_rough_32k*
