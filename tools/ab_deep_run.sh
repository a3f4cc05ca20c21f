for r in 1 2; do for v in B D E F; do
  a=$(BCMPC_LIB=bc_mpc_b200/_ab/lib_$v.so timeout 100 python tools/deep_bench.py --plans 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['kernel_ms_median'],2))")
  b=$(BCMPC_LIB=bc_mpc_b200/_ab/lib_$v.so timeout 100 python tools/deep_bench.py --K 8448 --H 60 --plans 7 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['kernel_ms_median']*1000/60,2))")
  echo "round $r variant $v: full plan $a ms, 33 units $b us/step"
done; done
