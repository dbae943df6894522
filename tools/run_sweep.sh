#!/bin/bash
# synthetic get_emb_eri sweep of BASELINE.json configs[4] on one GPU -> gpurun_out/sweep.jsonl
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for w in c1_hchain c2_graphene c3_nio_uhf small sweep_222_300_1500_200 sweep_224_200_1000_100 sweep_333_100_500_150 sweep_442_200_1000_150 sweep_444_100_500_100 sweep_444_300_1500_200; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 2 --no-e2e --no-cpu --no-dmet 2>/dev/null | tail -1 >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    l = l.strip()
    if not l.startswith("{"): continue
    d = json.loads(l); c = d["config"]
    print("%-28s kmesh %-9s nao %3d naux %4d neo %3d : %8.1f ms  %6.2f TFLOP/s  stage1 %.2f of peak" % (
        c["workload"].split(":")[0], "x".join(map(str, c["kmesh"])), c["nao"], c["naux"], c["neo"], d["ms_per_step"], d["value"], d["roofline"]["frac"] or 0))
PY
