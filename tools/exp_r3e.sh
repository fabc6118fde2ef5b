#!/usr/bin/env bash
# round 2, session 2, call 5: permutation look-ahead; the whole GPU suite; the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3e_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r3e_pytest.log
tail -5 gpurun_out/r3e_pytest.log
timeout 400 python bench.py > gpurun_out/r3e_bench_n1.json 2> gpurun_out/r3e_bench_n1.err
echo "bench rc=$?"
B200SGD_NO_LOOKAHEAD=1 timeout 200 python bench.py --no-cpu --no-sweep --no-e2e > gpurun_out/r3e_bench_n1_nola.json 2> gpurun_out/r3e_bench_n1_nola.err
python - <<PY
import json
for f in ("r3e_bench_n1", "r3e_bench_n1_nola"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), d["quality"]["permutation_golden_check"], (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "parse failed", e)
PY
