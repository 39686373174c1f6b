#!/usr/bin/env bash
# Builds libwn_b200.so (sm_100a only) next to this script's parent package directory.
#   exact kernels  : -fmad=false  (reference operation order, un-fused)
#   fast kernels   : FMA allowed
# Host code is compiled without -march / with -ffp-contract=off so the libstdc++ polar method keeps
# the reference's accept/reject sequence.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/.."
BUILD="${HERE}/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -ccbin /usr/bin/g++ ${ARCH}"
mkdir -p "${BUILD}"
PTXAS_V="${WN_PTXAS_V:+-Xptxas -v}"
# MT19937 jump-ahead tables of wn_rng.cu: generated (and self-checked) by tools/gen_mt_jump_tables.py, ~7 s, cached
GEN="${HERE}/../../tools/gen_mt_jump_tables.py"
if [ ! -s "${BUILD}/wn_mt_jump_tables.inc" ] || [ "${GEN}" -nt "${BUILD}/wn_mt_jump_tables.inc" ]; then
    "${PYTHON:-python3}" "${GEN}" --check > "${BUILD}/wn_mt_jump_tables.inc.tmp"
    mv "${BUILD}/wn_mt_jump_tables.inc.tmp" "${BUILD}/wn_mt_jump_tables.inc"
fi
${NVCC} ${COMMON} ${PTXAS_V} -fmad=false -c "${HERE}/wn_tilegen.cu"        -o "${BUILD}/wn_tilegen.o"
${NVCC} ${COMMON} ${PTXAS_V} -fmad=false -c "${HERE}/wn_eval_exact.cu"     -o "${BUILD}/wn_eval_exact.o"
${NVCC} ${COMMON} ${PTXAS_V}             -c "${HERE}/wn_multiband_fast.cu" -o "${BUILD}/wn_multiband_fast.o"
${NVCC} ${COMMON} ${PTXAS_V} -fmad=false -c "${HERE}/wn_rng.cu"            -o "${BUILD}/wn_rng.o"
${NVCC} ${COMMON}               -fmad=false -c "${HERE}/wn_capi.cu"         -o "${BUILD}/wn_capi.o"
${NVCC} ${COMMON}               -fmad=false -c "${HERE}/wn_group.cu"        -o "${BUILD}/wn_group.o"
${NVCC} ${ARCH} -shared -ccbin /usr/bin/g++ -o "${OUT}/libwn_b200.so" \
    "${BUILD}/wn_tilegen.o" "${BUILD}/wn_eval_exact.o" "${BUILD}/wn_multiband_fast.o" "${BUILD}/wn_rng.o" "${BUILD}/wn_capi.o" "${BUILD}/wn_group.o" \
    -cudart static -ldl
echo "built ${OUT}/libwn_b200.so"
