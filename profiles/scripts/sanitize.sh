#!/usr/bin/env bash
# compute-sanitizer pass over the hot path: ONE tool per invocation (B200_PROFILING.md: never
# several tools in one gpurun call).  Usage: profiles/scripts/sanitize.sh memcheck|racecheck|synccheck|initcheck [log]
# Runs (1) the smoke case of __graft_entry__ (seeded tile build incl. the MT19937 jump kernels, folded FAST lattice with
# the replica kernel and its TMA ring, exact lattice) and (2) back-to-back device calls with tile rebuilds (side-stream
# chain, deferred frees, programmatic dependent launch).
set -uo pipefail
TOOL="${1:-memcheck}"
LOG="${2:-gpurun_out/sanitize_${TOOL}.log}"
cd "$(dirname "${BASH_SOURCE[0]}")/../.."
mkdir -p "$(dirname "${LOG}")"
{
  echo "== compute-sanitizer --tool ${TOOL}: smoke"
  compute-sanitizer --tool "${TOOL}" --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
  echo "exit code $?"
  echo "== compute-sanitizer --tool ${TOOL}: back-to-back device calls"
  compute-sanitizer --tool "${TOOL}" --error-exitcode 9 python -m pytest tests -q -m gpu -x \
      -k "back_to_back or replica_kernel_variants or device_group"
  echo "exit code $?"
} > "${LOG}" 2>&1
grep -E "ERROR SUMMARY|exit code|passed|failed|smoke ok" "${LOG}"
