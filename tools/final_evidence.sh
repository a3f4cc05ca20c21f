set -x
(time python -m pytest tests -m gpu -q) > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/f_bench.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/f_ref.log 2>&1; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/f_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_tmem_kernel -c 1 -f -o gpurun_out/f_tmem python tools/profile_tc.py > gpurun_out/f_ncu_tmem.log 2>&1; echo "ncu tmem rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_deep_kernel -c 1 -f -o gpurun_out/f_deep python tools/deep_bench.py --K 67584 --H 6 --plans 1 > gpurun_out/f_ncu_deep.log 2>&1; echo "ncu deep rc=$?"
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/f_tests.log
