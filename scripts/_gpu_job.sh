for sw in NONE TC_ATTN FUSED_FFN DWCONV_MMA DWCONV_TILED DWCONV3_TMA; do
  echo "== disable $sw"; env FVLA_DISABLE_$sw=1 python scripts/diag_ln.py 2>&1 | tail -6
done
