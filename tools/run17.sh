#!/bin/bash
# final-ish scaling run: N = 8, 4, 2 on one box (cfg4), plus the direct-overlap A/B at N = 8 and the no-present e2e diagnostic
mkdir -p gpurun_out
run() {  # name, n, extra env
  name=$1; n=$2; shift 2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/r2_s17_$name.err | grep "^{" > gpurun_out/r2_s17_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_s17_$name.json")); print("$name", round(d["value"],1), "fps  e2e", round(d["e2e"]["value"],1), "blocking", round(d["e2e"]["blocking_readback_value"],1), "checksum", d["e2e"]["checksum"], {k: round(v,3) for k,v in d["per_pass_ms"].items()})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2_s17_$name.err").read()[-1500:])
PY
}
run n8 8 MV_X=0
run n8_nooverlap 8 MV_DIRECT_OVERLAP=0
run n8_nopresent 8 MV_BENCH_E2E=nopresent
run n4 4 MV_X=0
run n2 2 MV_X=0
