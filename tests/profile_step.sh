KF='regex:k_(begin_step|risk_resolve|play_travel|move_commit|move_retry|claim_punish|repro_commit|repro_retry|finalize)'
python bench.py --steps 3 --warmup 3 --no-also --no-cpu-baseline --e2e-steps 1 > gpurun_out/s3_prof_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --launch-skip 747 --launch-count 27 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 3 --warmup 3 --no-also --no-cpu-baseline --e2e-steps 1 > gpurun_out/s3_ncu_l.log 2>&1
ncu --set full --import-source on --clock-control none -k "$KF" --launch-skip 756 --launch-count 9 -o gpurun_out/s3_full python bench.py --steps 3 --warmup 3 --no-also --no-cpu-baseline --e2e-steps 1 > gpurun_out/s3_ncu_f.log 2>&1
tail -2 gpurun_out/s3_ncu_f.log | cut -c1-200
