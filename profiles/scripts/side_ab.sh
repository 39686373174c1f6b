for cfg in "" "WN_SIDE_MAX_LOG2=30" "WN_SIDE_MAX_LOG2=30 WN_PDL_SIDE_MAIN=1" "WN_SIDE_MAX_LOG2=30 WN_PDL_SIDE_CHAIN=1" "WN_SIDE_MAX_LOG2=30 WN_PDL_SIDE_MAIN=1 WN_PDL_SIDE_CHAIN=1"; do
  echo "== $cfg"
  env $cfg TUNE_ONLY="default plan, 8" timeout 120 python profiles/scripts/tune_rep.py 1024 30 2>&1 | grep "default plan"
done
