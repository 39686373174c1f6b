#!/usr/bin/env bash
# The 1 -> 8 GPU scaling evidence of a round, one command on one multi-GPU box: bench.py at N = 1, 2, 4, 8 (the driver's
# own launch lines), the single-process device-group drivers for configs 3 and 4, the reference arm under torchrun, and
# the box's concurrent D2H ceiling.  Usage: profiles/scripts/scale_run.sh <tag> [steps]   (outputs gpurun_out/<tag>_*)
set -uo pipefail
TAG="${1:-scale}"
STEPS="${2:-50}"
cd "$(dirname "${BASH_SOURCE[0]}")/../.."
O="gpurun_out/${TAG}"
NG=$(nvidia-smi -L | wc -l)
python bench.py --steps "${STEPS}" --warmup 5 > "${O}_bench_n1.json" 2> "${O}_bench_n1.err"
for n in 2 4 8; do
  [ "${n}" -le "${NG}" ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "${n}" --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus "${n}" --steps "${STEPS}" --warmup 5 2> "${O}_bench_n${n}.err" | tail -1 > "${O}_bench_n${n}.json"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node "${NG}" --master-addr 127.0.0.1 --master-port 29599 \
    bench.py --impl reference --gpus "${NG}" --steps 2 --warmup 1 2> "${O}_ref_n${NG}.err" | tail -1 > "${O}_ref_n${NG}.json"
B=wavelet-noise-in-ray-tracing_b200/cpp/bin/sharded_b200
: > "${O}_group.jsonl"
for n in 1 2 4 8; do
  [ "${n}" -le "${NG}" ] || continue
  "${B}" volume --gpus "${n}" --reps 30 --sharding cyclic >> "${O}_group.jsonl" 2>&1
  "${B}" volume --gpus "${n}" --reps 30 --sharding slab >> "${O}_group.jsonl" 2>&1
  "${B}" plane --gpus "${n}" --reps 3 >> "${O}_group.jsonl" 2>&1
done
python profiles/scripts/d2h_ceiling.py > "${O}_d2h.json" 2> "${O}_d2h.err"
for f in "${O}"_bench_n*.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "N", d["n_gpus"], "value %.1f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["value"],
      "per-rank", ["%.4f" % x for x in d["per_rank_ms_per_step"]])
PY
done
cat "${O}_group.jsonl" | cut -c1-330
cat "${O}_d2h.json" "${O}_ref_n${NG}.json" | cut -c1-600
