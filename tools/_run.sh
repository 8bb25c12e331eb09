set -x
python tools/gpu_prof_eval.py 170 8 > gpurun_out/r02_prof_eval_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_eval.csv python tools/gpu_prof_eval.py 170 8 > gpurun_out/r02_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pack_to_slot|pack_copy" -c 3 -o gpurun_out/r02_pack python tools/gpu_prof_eval.py 170 8 > gpurun_out/r02_ncu2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:zgemm_kernel -s 60 -c 4 -o gpurun_out/r02_zgemm python tools/gpu_prof_eval.py 170 8 > gpurun_out/r02_ncu3.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"jacobi_rot_kernel|jacobi_blocks_kernel" -s 10 -c 4 -o gpurun_out/r02_svd python tools/gpu_prof_eval.py 170 8 > gpurun_out/r02_ncu4.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"merge_gate_kernel|build_factors_kernel|truncate_kernel" -s 12 -c 3 -o gpurun_out/r02_small python tools/gpu_prof_eval.py 170 8 > gpurun_out/r02_ncu5.log 2>&1
cat gpurun_out/r02_prof_eval_plain.log | tail -2
ls -la gpurun_out/*.ncu-rep
