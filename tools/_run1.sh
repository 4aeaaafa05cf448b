set -x
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_final_render python tools/prof_b.py > gpurun_out/r02v_ncu_render.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_final_bake python tools/prof_bake.py > gpurun_out/r02v_ncu_bake.log 2>&1; echo "rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02v_ncu_list.log 2>&1; echo "rc=$?"
ls -la gpurun_out/prof_r02_final_* gpurun_out/launches_r02_final.csv
