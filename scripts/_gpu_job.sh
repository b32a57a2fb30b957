timeout 60 python scripts/one_ffn_wide.py 300 20000 131072 2>&1 | grep -v "^$" | tail -6
for d in 4 3 7; do echo "dbg $d"; FVLA_FFN_WIDE_DEBUG=$d timeout 60 python scripts/one_ffn_wide.py 131072 2>&1 | grep fused; done
