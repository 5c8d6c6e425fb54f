for v in "" "WC_ATTN_SMALL=0" "WC_IGEMM_TMA_STORE=0" "WC_IGEMM_TMA_RES=0" "WC_ROW3=0" "WC_ATTN_TP=0"; do
  echo "=== $v im64 B=32"; env $v python tools/batch_invariance.py 64 64 128 32 2>&1 | tail -3
  echo "=== $v im128 B=3"; env $v python tools/batch_invariance.py 128 128 256 3 2>&1 | tail -3
done
