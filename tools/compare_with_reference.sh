#!/usr/bin/env bash
# The reference's own experiment settings (report/sese_report_tvas.tex:532-538) run with the
# UNMODIFIED reference binary (both code paths) and with this repository's CLI on the same GPU.
#   features: 10k x {200..1000}, B = 1000, lr 0.1, 20 epochs;  samples: {50k,100k} x 4, B = 1000, lr 0.01, 10 epochs
# Usage (GPU box): bash tools/compare_with_reference.sh <workdir>
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
W="${1:-/tmp/cmp}"; mkdir -p "$W/data" "$W/ref0" "$W/ref1" "$W/ours"
python - "$W" <<'PY'
import sys, numpy as np
w = sys.argv[1]
def gen(path, R, C, seed=42):
    rng = np.random.default_rng(seed); X = rng.standard_normal((R, C)).astype(np.float32)
    wt = (100.0 * rng.random(C)).astype(np.float32); y = (X @ wt).astype(np.float32)
    with open(path, "w") as f:
        for row, lab in zip(X, y): f.write(",".join("%.9g" % v for v in row) + ",%.9g\n" % lab)
for C in (200, 600, 1000): gen(f"{w}/data/f-{C}.csv", 10000, C)
for R in (50000, 100000): gen(f"{w}/data/s-{R}.csv", R, 4)
PY
run() { # dir binary cudaRun lr epochs csv R C B
  ( cd "$1" && "$2" "$4" "$5" "$6" "$7" "$8" "$9" "$3" > out.txt 2>&1; \
    f=$(basename "$6" .csv); s=$([ "$3" = 0 ] && echo cublas || echo cuda); \
    python -c "import json;d=json.load(open('$f-$s.json'));print('%-8s %-22s cudaRun=%s gpu_time=%10.3f ms  error=%.6g' % ('$(basename $1)', '$f', '$3', d['gpu_time'], d['error']))" )
}
for C in 200 600 1000; do
  run "$W/ref0" "$HERE/oracle/_ref/main" 0 0.1 20 "$W/data/f-$C.csv" 10000 $C 1000
  run "$W/ref1" "$HERE/oracle/_ref/main" 1 0.1 20 "$W/data/f-$C.csv" 10000 $C 1000
  run "$W/ours" "$HERE/cuda_sgd_sese_project_b200/main" 0 0.1 20 "$W/data/f-$C.csv" 10000 $C 1000
done
for R in 50000 100000; do
  run "$W/ref0" "$HERE/oracle/_ref/main" 0 0.01 10 "$W/data/s-$R.csv" $R 4 1000
  run "$W/ref1" "$HERE/oracle/_ref/main" 1 0.01 10 "$W/data/s-$R.csv" $R 4 1000
  run "$W/ours" "$HERE/cuda_sgd_sese_project_b200/main" 0 0.01 10 "$W/data/s-$R.csv" $R 4 1000
done
