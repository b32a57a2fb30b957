for d in 0 1 2 4 3 5 6 7; do echo "debug=$d"; FVLA_DW7_DEBUG=$d python scripts/bench_ops.py --what dwconv --batch 32 2>&1 | sed -n 3p; done
