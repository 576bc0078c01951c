#!/bin/bash
# final verification on one B200: host overhead, GPU test-suite, default bench, smoke
timeout 120 python tools/host_overhead_profile.py 2>&1 | head -1
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_pytest_gpu_final.log
timeout 800 python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_final_n1.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["gemm"]["tflops"], d["wan_fp8_attention"]["ms_per_step"], d["wan_sparse"]["ms_per_step"])
for w in ("flux", "sd3", "qwen"):
    print(w, d[w]["ms_per_step"], d[w]["e2e_ms"])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
